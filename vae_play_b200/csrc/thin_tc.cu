// "Thin" layers on tcgen05: convolutions with one or two channels on one side (the first layer of the encoder, the
// last layer of the decoder and their gradients).  Their contraction index is (tap, thin channel) -- 25 values for a
// 5x5 grayscale layer -- so the channels-last tensor cannot feed the tensor core directly.  Each kernel here builds
// the operand it needs in shared memory:
//
//   thin_k_kernel   D[pix][n]   = sum_k Im[pix][k] . W[n][k]          first-layer forward, last-layer data gradient
//                   Im[pix][8i+j] = S[pix*stride + (i, j) + offset] is assembled by 128 builder threads as a K-major
//                   SWIZZLE_128B tile (row = pixel); the weight tile is built once per CTA from the fp32 master.
//   thin_w_kernel   dW[k][ch]  += sum_pix Im[pix][k] . Wide[pix][ch]   both thin weight gradients
//                   the same Im tile read as an MN-major operand (row = reduction index), the wide tensor arrives by
//                   TMA as an MN-major box; one accumulator per CTA lives in TMEM for the whole launch.
//   thin_n_kernel   P[hpix][t*CT+c] = sum_k A[hpix][k] . W[t,c][k];  y[pix][c] = act(b + sum_t P[pix (+) t][t*CT+c])
//                   last-layer forward (64 -> 1 channels): ONE GEMM over the (16+4)x(8+4) halo of a 16x8 brick gives,
//                   for every halo pixel, its contribution to each of the 25 outputs it touches; the epilogue adds the
//                   25 shifted planes through shared memory.  25x fewer MMA instructions than a tap loop with N=16.
//
// All three read the fp32 master weight / write the fp32 torch-layout gradient directly: no packing passes.
#include <cstring>

#include "tc_common.cuh"

namespace vp {
namespace {

using namespace tc;

constexpr int kMaxThinTaps = 32;

// How the 128 builder threads (warps 0-3) turn a 16 x 8 brick of grid pixels into a [128 rows][64 K] bf16 operand tile.
// The thin tensor S has ONE channel.  K index k = 8*i + j  <->  source pixel (gy*stride + oymin + i, gx*stride + oxmin + j),
// i < nrows <= 8 filter rows, j < 8 columns of which the first ncols are real taps: the padding columns read neighbouring
// (finite) pixels and meet zero weights.  With this layout chunk i of a tile row (16 bytes = 8 K values) is 8 CONSECUTIVE
// source pixels, so a row costs nrows 16-byte moves instead of 25 two-byte gathers + packing.
struct ThinGather {
    const bf16* S;          // [n][hs][ws]
    int hs, ws;
    int stride;
    int oymin, oxmin;       // offset of K row 0 / column 0
    int nrows;
    int Hh, Wc;             // source halo of one brick: Hh = 15*stride + nrows rows of Wc = even(7*stride + 8) pixels
    int8_t kmap[64];        // K index -> tap index of the weight (ky*kw + kx), or -1 for padding
};

constexpr int kHaloElems = 1664;      // >= Hh*Wc (stride <= 3)

bool finish_gather(ThinGather& g, int ncols) {
    g.Hh = 15 * g.stride + g.nrows;
    g.Wc = (7 * g.stride + 8 + 1) & ~1;
    return g.nrows <= 8 && ncols <= 8 && g.Hh * g.Wc <= kHaloElems;
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ uint32_t lds_b32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, unsigned short v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// Software pipeline across tiles, so that no global-memory latency is exposed:
//   prefetch(i+1): the source halo of the NEXT brick is loaded into registers (coalesced 2-byte loads, zero outside the image)
//   commit(i):     the registers of THIS brick go to shared memory -- twice: copy A at element e, copy B at e-1, so that a row
//                  segment starting at an odd element is 4-byte aligned in copy B; 128-thread named barrier
//   gather(i):     every thread moves nrows 16-byte segments from the halo into its swizzled tile row.
// Two halo buffers alternate per tile, which makes the single barrier sufficient.
constexpr int kHaloBuf = 2 * kHaloElems * 2 + 16;   // bytes of one halo buffer (copies A and B)

// NV = halo values per builder thread (3 at stride 1, 7 at stride 2, 13 at stride 3 for kernels up to 8x8)
template <int NV>
struct BuilderState {
    uint32_t hp[NV];           // (halo row << 16) | halo column of the thread's i-th value, ~0u = none
    int rel[NV];               // hy * ws + hx: source offset relative to the halo origin
    unsigned short hv[NV];     // prefetched values (raw; `ok` says which are inside the image)
    uint32_t ok;
};

// tile index -> (image m, brick row th, brick column tw), advanced by gridDim.x per step without divisions
struct TileWalk {
    int tw, th, m;
    int dw, dh, dm;
    __device__ __forceinline__ void init(int q, int step, int tiles_w, int tiles_h) {
        tw = q % tiles_w; q /= tiles_w; th = q % tiles_h; m = q / tiles_h;
        dw = step % tiles_w; step /= tiles_w; dh = step % tiles_h; dm = step / tiles_h;
    }
    __device__ __forceinline__ void next(int tiles_w, int tiles_h) {
        tw += dw; th += dh; m += dm;
        if (tw >= tiles_w) { tw -= tiles_w; ++th; }
        if (th >= tiles_h) { th -= tiles_h; ++m; }
    }
};

template <int NV>
__device__ __forceinline__ void halo_plan(const ThinGather& g, int r, BuilderState<NV>& b) {
    const int total = g.Hh * g.Wc;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int e = r + 128 * i;
        const int hy = e / g.Wc, hx = e - hy * g.Wc;
        b.hp[i] = e < total ? ((uint32_t)hy << 16) | (uint32_t)hx : ~0u;
        b.rel[i] = hy * g.ws + hx;
        b.hv[i] = 0;
    }
    b.ok = 0;
}
// issue the loads of one brick's halo; nothing here waits for them (the values are first touched by halo_commit)
template <int NV>
__device__ __forceinline__ void halo_prefetch(const ThinGather& g, BuilderState<NV>& b, int img, int gy0, int gx0) {
    const int sy_base = gy0 * g.stride + g.oymin, sx_base = gx0 * g.stride + g.oxmin;
    const unsigned short* __restrict__ base = reinterpret_cast<const unsigned short*>(g.S) + ((int64_t)img * g.hs + sy_base) * g.ws + sx_base;
    uint32_t ok = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int sy = sy_base + (int)(b.hp[i] >> 16), sx = sx_base + (int)(b.hp[i] & 0xffffu);
        const bool in = b.hp[i] != ~0u && (unsigned)sy < (unsigned)g.hs && (unsigned)sx < (unsigned)g.ws;
        if (in) b.hv[i] = __ldg(base + b.rel[i]);
        ok |= in ? (1u << i) : 0u;
    }
    b.ok = ok;
}
template <int NV>
__device__ __forceinline__ void halo_commit(uint32_t halo_saddr, int r, const BuilderState<NV>& b) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
        if (b.hp[i] != ~0u) {
            const uint32_t e = (uint32_t)(r + 128 * i);
            const unsigned short v = ((b.ok >> i) & 1u) ? b.hv[i] : (unsigned short)0;
            sts_u16(halo_saddr + e * 2u, v);                                         // copy A
            if (e > 0) sts_u16(halo_saddr + (kHaloElems + e - 1) * 2u, v);           // copy B (shifted by one element)
        }
}
// src_off: byte offset (inside a halo buffer) of the thread's first segment, already pointing into the aligned copy
__device__ __forceinline__ void halo_gather(uint32_t tile_saddr, uint32_t halo_saddr, int r, uint32_t src_off, int row_pitch_bytes, int nrows,
                                            bool valid) {
    const uint32_t row = tile_saddr + (uint32_t)r * 128u;
    const uint32_t src = halo_saddr + src_off;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
        if (valid && i < nrows) {
            const uint32_t a = src + (uint32_t)(i * row_pitch_bytes);
            w0 = lds_b32(a); w1 = lds_b32(a + 4); w2 = lds_b32(a + 8); w3 = lds_b32(a + 12);
        }
        sts_v4(row + (uint32_t)((i ^ (r & 7)) << 4), w0, w1, w2, w3);
    }
}

// The builder loop shared by thin_k_kernel and thin_w_kernel (warps 0-3): tiles q = blockIdx.x, +gridDim.x, ...;
// tile stage s of the ring lives at smem + s * stage_bytes and is handed to the MMA warp through full[s] (one arrive per warp).
template <int STAGES, int NV>
__device__ __forceinline__ void builder_loop_nv(const ThinGather& g, uint8_t* smem, int stage_bytes, uint8_t* halo, uint64_t* full, uint64_t* empty,
                                                int total_tiles, int tiles_w, int tiles_h, int gh, int gw) {
    const int r = threadIdx.x, lane = threadIdx.x & 31;
    const int by = r >> 3, bx = r & 7;
    BuilderState<NV> b;
    halo_plan<NV>(g, r, b);
    const int o = by * g.stride * g.Wc + bx * g.stride;               // element offset of the thread's first segment
    const uint32_t src_off = (o & 1) ? (uint32_t)(kHaloElems + o - 1) * 2u : (uint32_t)o * 2u;
    const int pitch = g.Wc * 2;
    int q = blockIdx.x;
    TileWalk tw;
    tw.init(q, gridDim.x, tiles_w, tiles_h);
    if (q < total_tiles) halo_prefetch<NV>(g, b, tw.m, tw.th * 16, tw.tw * 8);
    for (uint32_t it = 0; q < total_tiles; ++it) {
        const uint32_t halo_saddr = smem_u32(halo) + (it & 1) * kHaloBuf;
        halo_commit<NV>(halo_saddr, r, b);
        const bool valid = tw.th * 16 + by < gh && tw.tw * 8 + bx < gw;
        // next tile: its loads stay in flight during this tile's gather
        q += gridDim.x;
        tw.next(tiles_w, tiles_h);
        if (q < total_tiles) halo_prefetch<NV>(g, b, tw.m, tw.th * 16, tw.tw * 8);
        named_bar_sync(1, 128);
        const int s = it % STAGES;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        halo_gather(smem_u32(smem + s * stage_bytes), halo_saddr, r, src_off, pitch, g.nrows, valid);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
    }
}
template <int STAGES>
__device__ __forceinline__ void builder_loop(const ThinGather& g, uint8_t* smem, int stage_bytes, uint8_t* halo, uint64_t* full, uint64_t* empty,
                                             int total_tiles, int tiles_w, int tiles_h, int gh, int gw) {
    const int nv = (g.Hh * g.Wc + 127) >> 7;
    if (nv <= 3) builder_loop_nv<STAGES, 3>(g, smem, stage_bytes, halo, full, empty, total_tiles, tiles_w, tiles_h, gh, gw);
    else if (nv <= 7) builder_loop_nv<STAGES, 7>(g, smem, stage_bytes, halo, full, empty, total_tiles, tiles_w, tiles_h, gh, gw);
    else builder_loop_nv<STAGES, 13>(g, smem, stage_bytes, halo, full, empty, total_tiles, tiles_w, tiles_h, gh, gw);
}

// ===================================================================================================================
// thin_k_kernel
// ===================================================================================================================
struct ThinKParams {
    ThinGather g;
    const float* W;                   // fp32 master weight
    int64_t w_sn, w_st;               // element strides of (output channel n, tap index)
    void* D;
    const float* bias;
    int n, gh, gw, N;
    int act;
    float slope;
    int out_f32;
    int tiles_w, tiles_h, total_tiles;
    int tma_store;                    // bf16 output with N % 64 == 0: epilogue through shared memory + TMA store
    int tma_store32;                  // bf16 output with N % 32 == 0 (and not % 64): the same with 32-column SWIZZLE_64B tiles
    float* stat_parts;                // [gridDim.x][2][N] per-CTA BatchNorm partial sums of the stored output, or null
    // fused BatchNorm-backward reduction (N == 64 data gradient of the output layer): D is dL/da of the producer block, mapY its
    // pre-norm output; stat_parts then receives (sum d, sum d*(y - mean)) with d = D * (y*scale + shift > 0)
    const float *bn_scale, *bn_shift, *bn_mean;
};

constexpr int kKStages = 3;
constexpr int kKThreads = 288;        // warps 0-3 builders, warp 4 MMA, warps 5-8 epilogue

__global__ void __launch_bounds__(kKThreads, 2) thin_k_kernel(const __grid_constant__ CUtensorMap mapD, const __grid_constant__ CUtensorMap mapY,
                                                              const __grid_constant__ ThinKParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_b = smem + kKStages * 16384;                  // weight tile: N rows x 128 B (<= 16 KB)
    uint8_t* stage_out = smem_b + 16384;                        // per epilogue warp: 2 x [32 rows][128 B] staging tiles
    uint8_t* halo = stage_out + 4 * 2 * 4096;                   // 2 source-halo buffers
    float* s_stat = (float*)(halo + 2 * kHaloBuf);              // sum[128], sumsq[128]
    uint64_t* full = (uint64_t*)(s_stat + 256);
    uint64_t* empty = full + kKStages;
    uint64_t* acc_full = empty + kKStages;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* y_full = acc_empty + 2;                           // [4]: one y tile in flight per epilogue warp (fused BatchNorm backward)
    uint32_t* tmem_slot = (uint32_t*)(y_full + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kTmemCols = 256;                              // 2 accumulators x 128 columns
    if (threadIdx.x < 256) s_stat[threadIdx.x] = 0.f;

    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < kKStages; ++s) { mbar_init(&full[s], 4); mbar_init(&empty[s], 1); }   // one arrive per builder warp
            mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
            mbar_init(&acc_empty[0], 4); mbar_init(&acc_empty[1], 4);
            for (int i = 0; i < 4; ++i) mbar_init(&y_full[i], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // weight tile B[n][k] = W[n, kmap[k]] (zero for padding K), K-major SWIZZLE_128B.  Read BEFORE pdl_sync(): the fp32 master
    // weights are written by the optimiser kernels only, and those never let a dependent grid start early (common.cuh)
    for (int i = threadIdx.x; i < p.N * 64; i += kKThreads) {
        const int n = i >> 6, k = i & 63;
        const int t = p.g.kmap[k];
        const float v = t >= 0 ? p.W[n * p.w_sn + t * p.w_st] : 0.f;
        *reinterpret_cast<bf16*>(smem_b + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(v);
    }
    pdl_sync();     // everything above overlaps the previous kernel's tail; activations / gradients only from here on
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===== builders: one tile row (= output pixel) per thread =====
        builder_loop<kKStages>(p.g, smem, 16384, halo, full, empty, p.total_tiles, p.tiles_w, p.tiles_h, p.gh, p.gw);
    } else if (warp == 4) {
        // ===== MMA issuer =====
        if (elect_one()) {
            const uint32_t idesc = idesc_bf16_f32(128, p.N);
            const uint64_t bdesc = smem_desc_k_sw128(smem_u32(smem_b));
            uint32_t g = 0;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++g) {
                const int s = g % kKStages;
                const uint32_t buf = g & 1, use = g >> 1;
                mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
                mbar_wait(&full[s], (g / kKStages) & 1);
                tc_fence_after();
                const uint64_t adesc = smem_desc_k_sw128(smem_u32(smem + s * 16384));
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base + buf * 128, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, k != 0);
                tc_commit(&empty[s]);
                tc_commit(&acc_full[buf]);
            }
        }
    } else {
        // ===== epilogue: warps 5..8 -> TMEM lane quarter (warp & 3) =====
        const int lane_base = (warp & 3) * 32;
        const int r = lane_base + lane;
        const int by = r >> 3, bx = r & 7;
        uint32_t g = 0, sg = 0;
        const bool bn = p.bn_scale != nullptr;          // N == 64: one 64-column group per tile; staging slot 1 holds the y tile
        uint8_t* ybuf = stage_out + (warp & 3) * 8192 + 4096;
        uint64_t* ybar = &y_full[warp & 3];
        auto issue_y = [&](int q2) {                    // lane 0: the y tile of this warp's 32 rows of tile q2
            int m2 = q2;
            const int tw2 = m2 % p.tiles_w; m2 /= p.tiles_w;
            const int th2 = m2 % p.tiles_h; m2 /= p.tiles_h;
            mbar_expect_tx(ybar, 4096);
            tma_load_4d(ybuf, &mapY, ybar, 0, tw2 * 8, th2 * 16 + (warp & 3) * 4, m2);
        };
        if (bn && lane == 0 && (int)blockIdx.x < p.total_tiles) issue_y(blockIdx.x);
        for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++g) {
            int m = q;
            const int tw = m % p.tiles_w; m /= p.tiles_w;
            const int th = m % p.tiles_h; m /= p.tiles_h;
            const int gy = th * 16 + by, gx = tw * 8 + bx;
            const bool row_ok = gy < p.gh && gx < p.gw;
            const int64_t row_off = (((int64_t)m * p.gh + gy) * p.gw + gx) * p.N;
            const uint32_t buf = g & 1, use = g >> 1;
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
            if (p.tma_store) {
                uint8_t* my_stage = stage_out + (warp & 3) * 8192;
                const uint32_t row_mask = __ballot_sync(0xffffffffu, row_ok);
#pragma unroll 1
                for (int c = 0; c < p.N; c += 64, ++sg) {
                    uint8_t* st = my_stage + (bn ? 0 : (sg & 1) * 4096);
                    if (lane == 0) { if (bn) tma_store_wait_read<0>(); else tma_store_wait_read<1>(); }   // the store that last read this buffer has drained it
                    __syncwarp();
#pragma unroll
                    for (int cc = 0; cc < 64; cc += 32) {
                        uint32_t v[32];
                        tmem_ld32(tmem_base + buf * 128 + ((uint32_t)lane_base << 16) + (uint32_t)(c + cc), v);
                        tmem_ld_wait();
                        stage_chunk32(st, lane, c + cc, v, p.bias, p.act, p.slope);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (bn) {
                        mbar_wait(ybar, g & 1);
                        bnred_group64_sw128(st, ybuf, lane, s_stat + c, s_stat + 128 + c, row_mask, p.bn_scale + c, p.bn_shift + c, p.bn_mean + c);
                        __syncwarp();                                  // every lane is done with the y tile: fetch the next tile's
                        if (lane == 0 && q + (int)gridDim.x < p.total_tiles) issue_y(q + gridDim.x);
                    } else if (p.stat_parts) {
                        stats_group64_sw128(st, lane, s_stat + c, s_stat + 128 + c);
                    }
                    if (lane == 0) {
                        tma_store_4d(&mapD, st, c, tw * 8, th * 16 + (warp & 3) * 4, m);
                        tma_store_commit();
                    }
                }
            } else if (p.tma_store32) {
                // N % 32 == 0 (the VAE-GAN discriminator's 1 -> 32 first layer): 32-column [32 rows][64 B] SWIZZLE_64B staging tiles
                uint8_t* my_stage = stage_out + (warp & 3) * 8192;
#pragma unroll 1
                for (int c = 0; c < p.N; c += 32, ++sg) {
                    uint8_t* st = my_stage + (sg & 3) * 2048;
                    if (lane == 0) tma_store_wait_read<3>();
                    __syncwarp();
                    uint32_t v[32];
                    tmem_ld32(tmem_base + buf * 128 + ((uint32_t)lane_base << 16) + (uint32_t)c, v);
                    tmem_ld_wait();
                    stage_chunk32_sw64(st, lane, c, v, p.bias, p.act, p.slope);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(&mapD, st, c, tw * 8, th * 16 + (warp & 3) * 4, m);
                        tma_store_commit();
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < p.N; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + buf * 128 + ((uint32_t)lane_base << 16) + (uint32_t)c, v);
                    tmem_ld_wait();
                    if (row_ok) store_chunk<32>(v, p.D, row_off, c, p.N, p.bias, p.act, p.slope, p.out_f32 != 0);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if ((p.tma_store || p.tma_store32) && lane == 0) tma_store_wait_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
    if (p.stat_parts) {
        float* out = p.stat_parts + (size_t)blockIdx.x * 2 * p.N;
        for (int i = threadIdx.x; i < 2 * p.N; i += kKThreads) out[i] = s_stat[(i < p.N) ? i : (128 + i - p.N)];
    }
}

constexpr int kKSmem = kKStages * 16384 + 16384 + 4 * 2 * 4096 + 2 * kHaloBuf + 256 * 4 + (2 * kKStages + 4 + 4) * 8 + 16 + 1024;

int encode_box(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, int bw, int bh, int bc = 64);

int launch_thin_k(ThinKParams& p, cudaStream_t s, int stat_capacity = 0, int* stat_nparts = nullptr, const void* bn_y = nullptr) {
    CUtensorMap mD, mY;
    memset(&mD, 0, sizeof(mD));
    p.tma_store = (!p.out_f32 && p.N % 64 == 0) ? 1 : 0;
    p.tma_store32 = (!p.out_f32 && !p.tma_store && p.N % 32 == 0 && !p.stat_parts) ? 1 : 0;
    if (bn_y && (p.N != 64 || !p.tma_store || !p.stat_parts)) return VP_EUNSUPPORTED;
    if (p.stat_parts) {
        const int g = p.total_tiles < 2 * num_sms() ? p.total_tiles : 2 * num_sms();
        if (!p.tma_store || p.bias || p.act != VP_ACT_NONE) { set_error("thin_k: epilogue statistics need a bf16 output, N %% 64 == 0, no bias/activation"); return VP_EUNSUPPORTED; }
        if (g > stat_capacity) { set_error("thin_k: statistics buffer holds %d parts, %d needed", stat_capacity, g); return VP_EINVAL; }
        if (stat_nparts) *stat_nparts = g;
    }
    if (p.tma_store && encode_box(&mD, p.D, p.N, p.gw, p.gh, p.n, 8, 4)) { set_error("thin_k: cuTensorMapEncodeTiled(D) failed"); return VP_EUNSUPPORTED; }
    if (p.tma_store32 && encode_box(&mD, p.D, p.N, p.gw, p.gh, p.n, 8, 4, 32)) { set_error("thin_k: cuTensorMapEncodeTiled(D, 32) failed"); return VP_EUNSUPPORTED; }
    mY = mD;
    if (bn_y && encode_box(&mY, bn_y, p.N, p.gw, p.gh, p.n, 8, 4)) { set_error("thin_k: cuTensorMapEncodeTiled(y) failed"); return VP_EUNSUPPORTED; }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(thin_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKSmem);
        if (e != cudaSuccess) { set_error("thin_k: cannot set %d bytes of dynamic smem: %s", kKSmem, cudaGetErrorString(e)); return VP_ECUDA; }
        attr_set = true;
    }
    const int slots = 2 * num_sms();
    const int grid = p.total_tiles < slots ? p.total_tiles : slots;
    launch_k(thin_k_kernel, dim3(grid), dim3(kKThreads), kKSmem, s, mD, mY, p);
    VP_CHECK_LAUNCH("thin_k");
    return VP_OK;
}

// ===================================================================================================================
// thin_w_kernel
// ===================================================================================================================
struct ThinWParams {
    ThinGather g;
    float* dW;
    int64_t o_ch, o_t;                // element strides of (wide channel, tap index) in dW
    int n, gh, gw;                    // iteration grid = pixels of the wide tensor
    int tiles_w, tiles_h, total_tiles;
};

constexpr int kWStages = 2;
constexpr int kWThreads = 192;        // warps 0-3 builders (+ final epilogue), warp 4 MMA, warp 5 TMA
constexpr int kWStage = 32768;        // Im tile 16 KB + wide tile 16 KB

__host__ __device__ constexpr uint32_t idesc_mn(int m, int n) { return idesc_bf16_f32(m, n) | (1u << 15) | (1u << 16); }

__global__ void __launch_bounds__(kWThreads, 2) thin_w_kernel(const __grid_constant__ CUtensorMap mapWide, const __grid_constant__ ThinWParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* zero = smem + kWStages * kWStage;                  // 16 KB of zeros: rows 64..127 of the M = 128 operand
    uint8_t* halo = zero + 16384;                               // 2 source-halo buffers
    uint64_t* full = (uint64_t*)(halo + 2 * kHaloBuf);
    uint64_t* empty = full + kWStages;
    uint64_t* acc_full = empty + kWStages;
    uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cb = blockIdx.y;                                   // 64-channel block of the wide tensor

    if (warp == 5 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapWide) : "memory");
    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < kWStages; ++s) { mbar_init(&full[s], 5); mbar_init(&empty[s], 1); }   // 4 builder warps + TMA
            mbar_init(acc_full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 16384 / 16; i += kWThreads) reinterpret_cast<uint4*>(zero)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();
    const int my_tiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp < 4) {
        builder_loop<kWStages>(p.g, smem, kWStage, halo, full, empty, p.total_tiles, p.tiles_w, p.tiles_h, p.gh, p.gw);
        // final epilogue: rows m = K index of the accumulator live in TMEM lanes 0..63 (warps 0 and 1)
        if (my_tiles > 0 && warp < 2) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const int t = p.g.kmap[warp * 32 + lane];
            float* out = p.dW + (t >= 0 ? t * p.o_t : 0) + (int64_t)cb * 64 * p.o_ch;
#pragma unroll 1
            for (int c = 0; c < 64; c += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
                tmem_ld_wait();
                if (t >= 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(out + (c + j) * p.o_ch, __uint_as_float(v[j]));
                }
            }
        }
    } else if (warp == 5) {
        if (elect_one()) {
            uint32_t g = 0;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++g) {
                const int s = g % kWStages;
                int m = q;
                const int tw = m % p.tiles_w; m /= p.tiles_w;
                const int th = m % p.tiles_h; m /= p.tiles_h;
                mbar_wait(&empty[s], ((g / kWStages) & 1) ^ 1);
                mbar_expect_tx(&full[s], 16384);
                tma_load_4d(smem + s * kWStage + 16384, &mapWide, &full[s], cb * 64, tw * 8, th * 16, m);
            }
        }
    } else {
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_mn(128, 64);
            uint32_t g = 0;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++g) {
                const int s = g % kWStages;
                mbar_wait(&full[s], (g / kWStages) & 1);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * kWStage);
                // MN-major SW128: LBO = distance to the second group of 64 M rows (the zero region), SBO = 1024 B per 8 pixels
                const uint64_t lbo = (uint64_t)((smem_u32(zero) - sa) >> 4);
                const uint64_t adesc = (uint64_t)((sa & 0x3FFFF) >> 4) | (lbo << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
                const uint64_t bdesc = (uint64_t)(((sa + 16384) & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
                                       ((uint64_t)2 << 61);
#pragma unroll
                for (int k = 0; k < 8; ++k)      // 16 pixels per MMA = 2048 B
                    tc_mma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (g | k) != 0);
                tc_commit(&empty[s]);
            }
            if (my_tiles > 0) tc_commit(acc_full);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
    }
}

constexpr int kWSmem = kWStages * kWStage + 16384 + 2 * kHaloBuf + (2 * kWStages + 1) * 8 + 16 + 1024;

int launch_thin_w(const CUtensorMap& mw, const ThinWParams& p, int cblocks, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(thin_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem);
        if (e != cudaSuccess) { set_error("thin_w: cannot set %d bytes of dynamic smem: %s", kWSmem, cudaGetErrorString(e)); return VP_ECUDA; }
        attr_set = true;
    }
    int slots = 2 * num_sms() / cblocks;
    if (slots < 1) slots = 1;
    dim3 grid((unsigned)(p.total_tiles < slots ? p.total_tiles : slots), (unsigned)cblocks);
    launch_k(thin_w_kernel, dim3(grid), dim3(kWThreads), kWSmem, s, mw, p);
    VP_CHECK_LAUNCH("thin_w");
    return VP_OK;
}

// ===================================================================================================================
// thin_n_kernel
// ===================================================================================================================
struct ThinNParams {
    const float* W;
    int64_t w_sc, w_sk, w_st;         // element strides of (thin OUTPUT channel c, input channel k, tap index)
    const float* bias;
    void* D;
    int n, gh, gw, CT;
    int ntaps;
    int8_t ty[kMaxThinTaps], tx[kMaxThinTaps], widx[kMaxThinTaps];
    int tymin, txmin, Hh, Wh;
    int kblocks;
    int act;
    float slope;
    int out_f32;
    int tiles_w, tiles_h, total_tiles;
};

constexpr int kNStages = 2;
constexpr int kNStage = 32768;        // 256 halo rows x 128 B (240 used)
constexpr int kNThreads = 320;        // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int kNMaxKb = 4;
constexpr int kPStride = 36;          // floats per halo pixel in the partial-sum buffer (16-byte rows, conflict-free float4 stores)

__global__ void __launch_bounds__(kNThreads, 2) thin_n_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ ThinNParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_b = smem + kNStages * kNStage;               // weight tiles: kblocks x 4 KB
    float* psum = (float*)(smem_b + p.kblocks * 4096);         // [256][kPStride]
    uint64_t* full = (uint64_t*)(psum + 256 * kPStride);
    uint64_t* empty = full + kNStages;
    uint64_t* acc_full = empty + kNStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kTmemCols = 128;                              // 2 buffers x 2 halves x 32 columns
    const int nvals = p.ntaps * p.CT;

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kNStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
            mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
            mbar_init(&acc_empty[0], 8); mbar_init(&acc_empty[1], 8);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // weight tiles: per 64-channel k-block, 32 rows j = t*CT + c of 64 input channels (K-major SW128); read before pdl_sync()
    // like thin_k_kernel's (only the optimiser kernels write the fp32 master weights, and they hold their dependents back)
    for (int i = threadIdx.x; i < p.kblocks * 32 * 64; i += kNThreads) {
        const int k = i & 63, j = (i >> 6) & 31, kb = i >> 11;
        float v = 0.f;
        if (j < nvals) v = p.W[(j % p.CT) * p.w_sc + (int64_t)(kb * 64 + k) * p.w_sk + p.widx[j / p.CT] * p.w_st];
        *reinterpret_cast<bf16*>(smem_b + kb * 4096 + j * 128 + (((k >> 3) ^ (j & 7)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(v);
    }
    pdl_sync();     // everything above overlaps the previous kernel's tail; activations only from here on
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            uint32_t g = 0;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x) {
                int m = q;
                const int tw = m % p.tiles_w; m /= p.tiles_w;
                const int th = m % p.tiles_h; m /= p.tiles_h;
                for (int kb = 0; kb < p.kblocks; ++kb, ++g) {
                    const int s = g % kNStages;
                    mbar_wait(&empty[s], ((g / kNStages) & 1) ^ 1);
                    mbar_expect_tx(&full[s], p.Hh * p.Wh * 128);
                    tma_load_4d(smem + s * kNStage, &mapA, &full[s], kb * 64, tw * 8 + p.txmin, th * 16 + p.tymin, m);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_bf16_f32(128, 32);
            uint32_t g = 0, i = 0;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++i) {
                const uint32_t buf = i & 1, use = i >> 1;
                mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < p.kblocks; ++kb, ++g) {
                    const int s = g % kNStages;
                    mbar_wait(&full[s], (g / kNStages) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * kNStage);
                    const uint64_t bdesc = smem_desc_k_sw128(smem_u32(smem_b + kb * 4096));
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint64_t adesc = smem_desc_k_sw128(sa + half * 16384);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_bf16(tmem_base + (buf * 2 + half) * 32, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    }
                    tc_commit(&empty[s]);
                }
                tc_commit(&acc_full[buf]);
            }
        }
    } else {
        // ===== epilogue: 256 threads.  TMEM read: thread <-> halo pixel h.  Tap sum: two threads per brick pixel (even / odd
        // taps), all partial-sum loads of a thread issued together from offsets that do not depend on the tile. =====
        const int et = threadIdx.x - 64;                          // 0..255
        const int half = et >> 7;                                 // warps 2-5: rows 0..127, warps 6-9: rows 128..255
        const int lane_base = (warp & 3) * 32;
        const int h = half * 128 + lane_base + lane;
        const int pr = et >> 1, part = et & 1;                    // brick pixel, tap parity
        const int by = pr >> 3, bx = pr & 7;
        // shared-space byte addresses of this thread's 16 partial sums (taps part, part+2, ...); a tap index past the filter
        // points at the padding words of row 0 (columns 32..35 are never written by the tiles: cleared once, below)
        const uint32_t ps_u32 = smem_u32(psum);
        uint32_t poff[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int t = 2 * u + part;
            poff[u] = t < p.ntaps ? ps_u32 + 4u * (uint32_t)(((by + p.ty[t] - p.tymin) * p.Wh + bx + p.tx[t] - p.txmin) * kPStride + t * p.CT)
                                  : ps_u32 + 4u * 32u;
        }
        if (et < 4) psum[32 + et] = 0.f;
        const uint32_t my_row = ps_u32 + 4u * (uint32_t)(h * kPStride);
        uint32_t vmask = 0;                                       // bit u: tap 2u+part exists (its address moves with the channel c)
#pragma unroll
        for (int u = 0; u < 16; ++u) vmask |= (2 * u + part < p.ntaps ? 1u : 0u) << u;
        TileWalk tl;
        tl.init(blockIdx.x, gridDim.x, p.tiles_w, p.tiles_h);
        uint32_t i = 0;
        for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++i, tl.next(p.tiles_w, p.tiles_h)) {
            const uint32_t buf = i & 1, use = i >> 1;
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld32(tmem_base + (buf * 2 + half) * 32 + ((uint32_t)lane_base << 16), v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            named_bar_sync(2, 256);                               // everyone has finished reading the previous tile's sums
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + 4u * j), "r"(v[j]), "r"(v[j + 1]), "r"(v[j + 2]), "r"(v[j + 3])
                             : "memory");
            named_bar_sync(1, 256);
            const int gy = tl.th * 16 + by, gx = tl.tw * 8 + bx;
            const bool ok = part == 0 && gy < p.gh && gx < p.gw;
            const int64_t off0 = (((int64_t)tl.m * p.gh + gy) * p.gw + gx) * p.CT;
            for (int c = 0; c < p.CT; ++c) {
                float part_sum[16];
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(part_sum[u]) : "r"(poff[u] + ((vmask >> u) & 1u ? 4u * c : 0u)) : "memory");
                float a0 = part_sum[0] + part_sum[1], a1 = part_sum[2] + part_sum[3], a2 = part_sum[4] + part_sum[5], a3 = part_sum[6] + part_sum[7];
                a0 += part_sum[8] + part_sum[9]; a1 += part_sum[10] + part_sum[11]; a2 += part_sum[12] + part_sum[13]; a3 += part_sum[14] + part_sum[15];
                float acc = (a0 + a1) + (a2 + a3);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                if (ok) {
                    acc = act_fwd(acc + (p.bias ? p.bias[c] : 0.f), p.act, p.slope);
                    if (p.out_f32) ((float*)p.D)[off0 + c] = acc;
                    else ((bf16*)p.D)[off0 + c] = __float2bfloat16_rn(acc);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

constexpr int kNSmemBase = kNStages * kNStage + 256 * kPStride * 4 + (2 * kNStages + 4) * 8 + 16 + 1024;   // + kblocks * 4096

int encode_box(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, int bw, int bh, int bc) {
    EncodeTiledFn encode = get_encode();
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        bc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

bool thin_geom_ok(const VpConvGeom* g, const char* who) {
    if (!g) { set_error("%s: null geometry", who); return false; }
    if (g->n <= 0 || g->hi <= 0 || g->wi <= 0 || g->ci <= 0 || g->ho <= 0 || g->wo <= 0 || g->co <= 0 || g->kh <= 0 || g->kw <= 0 ||
        g->stride <= 0 || g->pad < 0) { set_error("%s: non-positive dimension in geometry", who); return false; }
    return true;
}

}  // namespace
}  // namespace vp

using namespace vp;

namespace {
// gather description for a single-channel thin tensor: K row i / column j  <->  source offset (oy0 + i*dy, ox0 + j*dx) would need
// signed steps; instead offsets always INCREASE with (i, j) and kmap says which weight tap sits there.
// forward-type gather  (x[p*s + k - pad]):  oymin = -pad, tap(i, j) = i*kw + j
// flipped gather       (dy[p + pad - k]):   oymin = pad - (kh-1), tap(i, j) = (kh-1-i)*kw + (kw-1-j)
void fill_gather(ThinGather& tg, const VpConvGeom* g, bool flipped) {
    tg.nrows = g->kh;
    tg.oymin = flipped ? g->pad - (g->kh - 1) : -g->pad;
    tg.oxmin = flipped ? g->pad - (g->kw - 1) : -g->pad;
    for (int k = 0; k < 64; ++k) {
        const int i = k >> 3, j = k & 7;
        tg.kmap[k] = (int8_t)((i < g->kh && j < g->kw) ? (flipped ? (g->kh - 1 - i) * g->kw + (g->kw - 1 - j) : i * g->kw + j) : -1);
    }
}
}  // namespace

// y = act(conv(x, w) + bias) for an nn.Conv2d with ONE input channel (thin K) or <= 2 output channels at stride 1 (thin N).
static int thin_conv_fwd_impl(const VpConvGeom* g, const void* x, const float* w, const float* bias, void* y, int out_dtype, int act, float slope,
                              float* stat_parts, int stat_capacity, int* stat_nparts, void* stream);

extern "C" int vp_thin_conv_fwd(const VpConvGeom* g, const void* x, const float* w, const float* bias, void* y, int out_dtype, int act,
                                float slope, void* stream) {
    return thin_conv_fwd_impl(g, x, w, bias, y, out_dtype, act, slope, nullptr, 0, nullptr, stream);
}

// same (single input channel, bf16 output, no bias / activation) + per-CTA BatchNorm partial sums of y: stat_parts[*nparts][2][co]
extern "C" int vp_thin_conv_fwd_stats(const VpConvGeom* g, const void* x, const float* w, void* y, float* stat_parts, int stat_capacity,
                                      int* nparts, void* stream) {
    VP_CHECK_ARG(stat_parts && nparts && stat_capacity > 0, "vp_thin_conv_fwd_stats: bad statistics buffer");
    if (g && g->ci != 1) { set_error("vp_thin_conv_fwd_stats: single-channel input only"); return VP_EUNSUPPORTED; }
    return thin_conv_fwd_impl(g, x, w, nullptr, y, VP_BF16, VP_ACT_NONE, 0.f, stat_parts, stat_capacity, nparts, stream);
}

static int launch_thin_n(const ThinNParams& p, const void* A, int C, int W, int H, int N, cudaStream_t s);
static int thin_conv_fwd_impl(const VpConvGeom* g, const void* x, const float* w, const float* bias, void* y, int out_dtype, int act, float slope,
                              float* stat_parts, int stat_capacity, int* stat_nparts, void* stream) {
    if (!thin_geom_ok(g, "vp_thin_conv_fwd")) return VP_EINVAL;
    VP_CHECK_ARG(x && w && y, "vp_thin_conv_fwd: null pointer");
    if (!tc_available() || g->transposed) { set_error("vp_thin_conv_fwd: needs sm_100 and a plain Conv2d"); return VP_EUNSUPPORTED; }
    const int T = g->kh * g->kw;
    cudaStream_t s = (cudaStream_t)stream;
    if (g->ci == 1 && g->kh <= 8 && g->kw <= 8 && g->stride <= 3 && g->co % 32 == 0 && g->co <= 128 && ((uintptr_t)y & 15) == 0) {
        ThinKParams p;
        memset(&p, 0, sizeof(p));
        p.g.S = (const bf16*)x; p.g.hs = g->hi; p.g.ws = g->wi; p.g.stride = g->stride;
        fill_gather(p.g, g, false);
        if (!finish_gather(p.g, g->kw)) { set_error("vp_thin_conv_fwd: halo too large"); return VP_EUNSUPPORTED; }
        p.W = w; p.w_sn = T; p.w_st = 1;
        p.D = y; p.bias = bias; p.n = g->n; p.gh = g->ho; p.gw = g->wo; p.N = g->co;
        p.act = act; p.slope = slope; p.out_f32 = out_dtype == VP_F32;
        p.tiles_w = (p.gw + 7) / 8; p.tiles_h = (p.gh + 15) / 16;
        const int64_t total = (int64_t)p.tiles_w * p.tiles_h * p.n;
        if (total > 0x7fffffff) { set_error("vp_thin_conv_fwd: too many tiles"); return VP_EUNSUPPORTED; }
        p.total_tiles = (int)total;
        p.stat_parts = stat_parts;
        return launch_thin_k(p, s, stat_capacity, stat_nparts);
    }
    if (stat_parts) { set_error("vp_thin_conv_fwd_stats: shape not of the thin-input form"); return VP_EUNSUPPORTED; }
    if (g->stride == 1 && g->co * T <= 32 && T <= kMaxThinTaps && g->ci % 64 == 0 && g->ci <= 64 * kNMaxKb && g->kh <= 5 && g->kw <= 5 &&
        ((uintptr_t)x & 15) == 0) {
        ThinNParams p;
        memset(&p, 0, sizeof(p));
        p.W = w; p.w_sc = (int64_t)g->ci * T; p.w_sk = T; p.w_st = 1;
        p.bias = bias; p.D = y; p.n = g->n; p.gh = g->ho; p.gw = g->wo; p.CT = g->co; p.ntaps = T;
        for (int ky = 0; ky < g->kh; ++ky)
            for (int kx = 0; kx < g->kw; ++kx) {
                const int t = ky * g->kw + kx;
                p.ty[t] = (int8_t)(ky - g->pad); p.tx[t] = (int8_t)(kx - g->pad); p.widx[t] = (int8_t)t;
            }
        p.tymin = -g->pad; p.txmin = -g->pad; p.Hh = 16 + g->kh - 1; p.Wh = 8 + g->kw - 1;
        p.kblocks = g->ci / 64; p.act = act; p.slope = slope; p.out_f32 = out_dtype == VP_F32;
        p.tiles_w = (p.gw + 7) / 8; p.tiles_h = (p.gh + 15) / 16;
        const int64_t total = (int64_t)p.tiles_w * p.tiles_h * p.n;
        if (total > 0x7fffffff) { set_error("vp_thin_conv_fwd: too many tiles"); return VP_EUNSUPPORTED; }
        p.total_tiles = (int)total;
        return launch_thin_n(p, x, g->ci, g->wi, g->hi, g->n, s);
    }
    set_error("vp_thin_conv_fwd: shape not of a thin form (ci=%d co=%d k=%dx%d s=%d)", g->ci, g->co, g->kh, g->kw, g->stride);
    return VP_EUNSUPPORTED;
}

// the wide tensor A [N][H][W][C] (C = 64 * kblocks) of a thin_n problem -> tensor map, shared-memory attribute, launch
static int launch_thin_n(const ThinNParams& p, const void* A, int C, int W, int H, int N, cudaStream_t s) {
    CUtensorMap mA;
    if (encode_box(&mA, A, C, W, H, N, p.Wh, p.Hh)) { set_error("thin_n: cuTensorMapEncodeTiled failed"); return VP_EUNSUPPORTED; }
    const int smem_bytes = kNSmemBase + p.kblocks * 4096;
    static int attr_set = 0;
    if (attr_set < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(thin_n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) { set_error("thin_n: cannot set %d bytes of dynamic smem: %s", smem_bytes, cudaGetErrorString(e)); return VP_ECUDA; }
        attr_set = smem_bytes;
    }
    const int slots = num_sms() * (2 * smem_bytes <= 227 * 1024 ? 2 : 1);
    const int grid = p.total_tiles < slots ? p.total_tiles : slots;
    launch_k(thin_n_kernel, dim3(grid), dim3(kNThreads), smem_bytes, s, mA, p);
    VP_CHECK_LAUNCH("thin_n");
    return VP_OK;
}

/* dx [n][hi][wi][1] = dL/dx of a stride-1 nn.Conv2d with ONE INPUT channel (the VAE-GAN discriminator's first layer, whose input
 * x_tilde needs a gradient: models/networks.py:159):  dx[p] = sum_t sum_co dy[p + pad - t][co] . w[co][0][t]  -- the thin-output
 * forward kernel on dy with flipped taps.  co % 64 == 0 (the caller pads 32 -> 64), k <= 5, dx bf16 or fp32. */
extern "C" int vp_thin_conv_dgrad_in1(const VpConvGeom* g, const void* dy, const float* w, void* dx, int out_dtype, void* stream) {
    if (!thin_geom_ok(g, "vp_thin_conv_dgrad_in1")) return VP_EINVAL;
    VP_CHECK_ARG(dy && w && dx, "vp_thin_conv_dgrad_in1: null pointer");
    const int T = g->kh * g->kw;
    if (!tc_available() || g->transposed || g->stride != 1 || g->ci != 1 || g->co % 64 != 0 || g->co > 64 * kNMaxKb || g->kh > 5 || g->kw > 5 ||
        T > kMaxThinTaps || ((uintptr_t)dy & 15) != 0) {
        set_error("vp_thin_conv_dgrad_in1: shape not served (ci=%d co=%d k=%dx%d s=%d)", g->ci, g->co, g->kh, g->kw, g->stride);
        return VP_EUNSUPPORTED;
    }
    ThinNParams p;
    memset(&p, 0, sizeof(p));
    p.W = w; p.w_sc = 0; p.w_sk = T; p.w_st = 1;
    p.bias = nullptr; p.D = dx; p.n = g->n; p.gh = g->hi; p.gw = g->wi; p.CT = 1; p.ntaps = T;
    for (int ky = 0; ky < g->kh; ++ky)
        for (int kx = 0; kx < g->kw; ++kx) {
            const int t = ky * g->kw + kx;
            p.ty[t] = (int8_t)(g->pad - ky); p.tx[t] = (int8_t)(g->pad - kx); p.widx[t] = (int8_t)t;
        }
    p.tymin = g->pad - (g->kh - 1); p.txmin = g->pad - (g->kw - 1); p.Hh = 16 + g->kh - 1; p.Wh = 8 + g->kw - 1;
    p.kblocks = g->co / 64; p.act = VP_ACT_NONE; p.slope = 0.f; p.out_f32 = out_dtype == VP_F32;
    p.tiles_w = (p.gw + 7) / 8; p.tiles_h = (p.gh + 15) / 16;
    const int64_t total = (int64_t)p.tiles_w * p.tiles_h * p.n;
    if (total > 0x7fffffff) { set_error("vp_thin_conv_dgrad_in1: too many tiles"); return VP_EUNSUPPORTED; }
    p.total_tiles = (int)total;
    return launch_thin_n(p, dy, g->co, g->wo, g->ho, g->n, (cudaStream_t)stream);
}

// dx = dL/dx of a stride-1 nn.Conv2d with ONE output channel:  dx[p][ci] = sum_t dy[p + pad - t] . w[0][ci][t]
extern "C" int vp_thin_conv_dgrad(const VpConvGeom* g, const void* dy, const float* w, void* dx, int out_dtype, void* stream) {
    if (!thin_geom_ok(g, "vp_thin_conv_dgrad")) return VP_EINVAL;
    VP_CHECK_ARG(dy && w && dx, "vp_thin_conv_dgrad: null pointer");
    const int T = g->kh * g->kw;
    if (!tc_available() || g->transposed || g->stride != 1 || g->co != 1 || g->kh > 8 || g->kw > 8 || g->ci % 32 != 0 || g->ci > 128 ||
        ((uintptr_t)dx & 15) != 0) {
        set_error("vp_thin_conv_dgrad: shape not of the thin form (ci=%d co=%d k=%dx%d s=%d)", g->ci, g->co, g->kh, g->kw, g->stride);
        return VP_EUNSUPPORTED;
    }
    ThinKParams p;
    memset(&p, 0, sizeof(p));
    p.g.S = (const bf16*)dy; p.g.hs = g->ho; p.g.ws = g->wo; p.g.stride = 1;
    fill_gather(p.g, g, true);
    if (!finish_gather(p.g, g->kw)) { set_error("vp_thin_conv_dgrad: halo too large"); return VP_EUNSUPPORTED; }
    p.W = w; p.w_sn = T; p.w_st = 1;                                   // n = input channel ci
    p.D = dx; p.bias = nullptr; p.n = g->n; p.gh = g->hi; p.gw = g->wi; p.N = g->ci;
    p.act = VP_ACT_NONE; p.slope = 0.f; p.out_f32 = out_dtype == VP_F32;
    p.tiles_w = (p.gw + 7) / 8; p.tiles_h = (p.gh + 15) / 16;
    const int64_t total = (int64_t)p.tiles_w * p.tiles_h * p.n;
    if (total > 0x7fffffff) { set_error("vp_thin_conv_dgrad: too many tiles"); return VP_EUNSUPPORTED; }
    p.total_tiles = (int)total;
    return launch_thin_k(p, (cudaStream_t)stream);
}

/* vp_thin_conv_dgrad (bf16 dx, ci == 64) fused with the first pass of the PRODUCER block's BatchNorm backward, as
 * vp_conv_dgrad_cl_bnred: parts[*nparts][2][ci] = per-CTA (sum d, sum d*(y - mean)), d = dx * (y_prev*scale + shift > 0). */
extern "C" int vp_thin_conv_dgrad_bnred(const VpConvGeom* g, const void* dy, const float* w, void* dx, const void* y_prev, const float* scale,
                                        const float* shift, const float* mean, float* parts, int capacity, int* nparts, void* stream) {
    if (!thin_geom_ok(g, "vp_thin_conv_dgrad_bnred")) return VP_EINVAL;
    VP_CHECK_ARG(dy && w && dx && y_prev && scale && shift && mean && parts && nparts && capacity > 0, "vp_thin_conv_dgrad_bnred: null pointer");
    const int T = g->kh * g->kw;
    if (!tc_available() || g->transposed || g->stride != 1 || g->co != 1 || g->kh > 8 || g->kw > 8 || g->ci != 64 || ((uintptr_t)dx & 15) != 0 ||
        ((uintptr_t)y_prev & 15) != 0)
        return VP_EUNSUPPORTED;
    ThinKParams p;
    memset(&p, 0, sizeof(p));
    p.g.S = (const bf16*)dy; p.g.hs = g->ho; p.g.ws = g->wo; p.g.stride = 1;
    fill_gather(p.g, g, true);
    if (!finish_gather(p.g, g->kw)) return VP_EUNSUPPORTED;
    p.W = w; p.w_sn = T; p.w_st = 1;
    p.D = dx; p.bias = nullptr; p.n = g->n; p.gh = g->hi; p.gw = g->wi; p.N = g->ci;
    p.act = VP_ACT_NONE; p.slope = 0.f; p.out_f32 = 0;
    p.tiles_w = (p.gw + 7) / 8; p.tiles_h = (p.gh + 15) / 16;
    const int64_t total = (int64_t)p.tiles_w * p.tiles_h * p.n;
    if (total > 0x7fffffff) return VP_EUNSUPPORTED;
    p.total_tiles = (int)total;
    p.stat_parts = parts; p.bn_scale = scale; p.bn_shift = shift; p.bn_mean = mean;
    return launch_thin_k(p, (cudaStream_t)stream, capacity, nparts, y_prev);
}

// dw (fp32, torch layout [co][ci][kh][kw], zeroed by the call) = dL/dw of an nn.Conv2d with ONE input channel, or with ONE
// output channel at stride 1.
extern "C" int vp_thin_conv_wgrad(const VpConvGeom* g, const void* x, const void* dy, float* dw, int accumulate, void* stream) {
    if (!thin_geom_ok(g, "vp_thin_conv_wgrad")) return VP_EINVAL;
    VP_CHECK_ARG(x && dy && dw, "vp_thin_conv_wgrad: null pointer");
    const int T = g->kh * g->kw;
    cudaStream_t s = (cudaStream_t)stream;
    if (!tc_available() || g->transposed || g->kh > 8 || g->kw > 8) { set_error("vp_thin_conv_wgrad: needs sm_100 and a plain Conv2d up to 8x8"); return VP_EUNSUPPORTED; }
    ThinWParams p;
    memset(&p, 0, sizeof(p));
    const void* wide = nullptr;
    int wc = 0;
    if (g->ci == 1 && g->co % 64 == 0 && g->stride <= 3) {
        // thin input: wide = dy on the output grid, Im[p][k] = x[p*stride + k - pad]
        wide = dy; wc = g->co;
        p.g.S = (const bf16*)x; p.g.hs = g->hi; p.g.ws = g->wi; p.g.stride = g->stride;
        p.n = g->n; p.gh = g->ho; p.gw = g->wo;
        fill_gather(p.g, g, false);
        p.o_ch = T; p.o_t = 1;
    } else if (g->co == 1 && g->ci % 64 == 0 && g->stride == 1) {
        // thin output: wide = x on the input grid, Im[p][k] = dy[p + pad - k]
        wide = x; wc = g->ci;
        p.g.S = (const bf16*)dy; p.g.hs = g->ho; p.g.ws = g->wo; p.g.stride = 1;
        p.n = g->n; p.gh = g->hi; p.gw = g->wi;
        fill_gather(p.g, g, true);
        p.o_ch = T; p.o_t = 1;
    } else {
        set_error("vp_thin_conv_wgrad: shape not of a thin form (ci=%d co=%d k=%dx%d s=%d)", g->ci, g->co, g->kh, g->kw, g->stride);
        return VP_EUNSUPPORTED;
    }
    if (((uintptr_t)wide & 15) != 0) { set_error("vp_thin_conv_wgrad: 16-byte alignment"); return VP_EUNSUPPORTED; }
    if (!finish_gather(p.g, g->kw)) { set_error("vp_thin_conv_wgrad: halo too large"); return VP_EUNSUPPORTED; }
    p.dW = dw;
    p.tiles_w = (p.gw + 7) / 8; p.tiles_h = (p.gh + 15) / 16;
    const int64_t total = (int64_t)p.tiles_w * p.tiles_h * p.n;
    if (total > 0x7fffffff) { set_error("vp_thin_conv_wgrad: too many tiles"); return VP_EUNSUPPORTED; }
    p.total_tiles = (int)total;
    CUtensorMap mw;
    if (encode_box(&mw, wide, wc, p.gw, p.gh, p.n, 8, 16)) { set_error("vp_thin_conv_wgrad: cuTensorMapEncodeTiled failed"); return VP_EUNSUPPORTED; }
    if (!accumulate) zero_async(dw, sizeof(float) * (size_t)g->co * g->ci * T, s);
    return launch_thin_w(mw, p, wc / 64, s);
}
