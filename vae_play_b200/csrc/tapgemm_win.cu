// Windowed tap GEMM (tcgen05): stride-1 tap sets (every output-parity phase of a transposed conv / strided-conv
// dgrad, and plain stride-1 convs) with the A operand loaded ONCE per k-block as a halo tile and re-used by all taps.
//
// tapgemm_tc_kernel reloads the 128 x 64 A tile for every tap: a 3x3-tap phase moves 9x the activation bytes
// through L2 -> SMEM, and ncu shows that kernel bound by exactly that traffic (tensor pipe 11..25 % busy).  Here an
// M-tile is a pair of bricks of 16 rows x 8 columns of output pixels of one image.  Per 64-channel k-block each
// brick's input halo ((16+dy) x (8+dx) pixels, 128 B per pixel) arrives by ONE 4-D TMA box; the tap (ty,tx) operand
// is the *window* of that halo starting at row (ty-tymin)*Wh + (tx-txmin): 16 groups of 8 consecutive 128-byte rows
// with group stride Wh*128 B, which is a legal K-major SWIZZLE_128B UMMA operand because the hardware swizzle is a
// function of the absolute shared-memory address (tools/probe_umma.py: any 128-byte-aligned start and any group
// stride read back exactly).  The weight tile of a (tap, k-block) is streamed once and feeds both bricks, so
// L2->SMEM bytes per MMA drop ~3.5x for 64-channel outputs.
//
// Warp roles (224 threads, 1 CTA/SM): warp 0 = halo (A) TMA producer, warp 6 = weight (B) TMA producer,
// warp 1 = MMA issuer (accumulators double-buffered in TMEM: 2 buffers x 2 bricks x BN columns),
// warps 2..5 = epilogue.  All phases of a layer are tiles of one persistent launch.
#include <cstring>

#include "tc_common.cuh"

namespace vp {
namespace {

using namespace tc;

constexpr int kWThreads = 224;
constexpr int kBrickH = 16, kBrickW = 8;
constexpr int kAStages = 2;
constexpr int kMaxPh = 4;

struct WinPhase {
    int gh, gw, doy, dox;
    int tiles_w, tiles_h;      // bricks along x / brick PAIRS along y ... see decode()
    int tile_begin;
    TapList taps;
};

struct WinParams {
    void* D;
    const float* bias;
    int n, hd, wd, N;
    int ds;
    int act;
    float slope;
    int out_f32;
    int kblocks;
    int tymin, txmin;          // halo origin relative to the brick origin (common to all phases)
    int Hh, Wh;                // halo extent in pixels
    int halo_bytes;            // Hh*Wh*128 rounded up to 1024
    int ntiles_n, total_tiles, nphases;
    int tma_store;             // bf16 output, N % 64 == 0: staging tiles + TMA stores (omaps)
    float* stat_parts;         // [gridDim.x][2][N] per-CTA BatchNorm partial sums of the stored output, or null
    WinPhase ph[kMaxPh];
};

struct WTile { int phase, n, gy0, gx0, col0; };   // a tile = two bricks: (gy0, gx0) and (gy0, gx0 + 8)

__device__ __forceinline__ WTile wdecode(const WinParams& p, int q, int BN) {
    WTile c;
    const int nt = q % p.ntiles_n;
    int mt = q / p.ntiles_n;
    int ph = 0;
#pragma unroll
    for (int i = 1; i < kMaxPh; ++i)
        if (i < p.nphases && mt >= p.ph[i].tile_begin) ph = i;
    mt -= p.ph[ph].tile_begin;
    const int tw = mt % p.ph[ph].tiles_w; mt /= p.ph[ph].tiles_w;
    const int th = mt % p.ph[ph].tiles_h; mt /= p.ph[ph].tiles_h;
    c.phase = ph; c.n = mt; c.gy0 = th * kBrickH; c.gx0 = tw * (2 * kBrickW); c.col0 = nt * BN;
    return c;
}

// MN-major SWIZZLE_128B weight tile (the module's channels-last weight read in place): 64 (n) x 64 (k) boxes, LBO = one box
__device__ __forceinline__ uint64_t wdesc_mn_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int BN, int BSTAGES, bool BMN>
__global__ void __launch_bounds__(kWThreads, 1) tapgemm_win_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                   const __grid_constant__ CUtensorMap mapB,
                                                                   const __grid_constant__ OutMaps omaps,
                                                                   const __grid_constant__ WinParams p) {
    constexpr int kBBytes = BN * 128;
    constexpr int kAccCols = BN < 32 ? 32 : BN;        // TMEM columns per brick accumulator
    constexpr int kTmemCols = 4 * kAccCols;            // 2 buffers x 2 bricks
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_stage_bytes = 2 * p.halo_bytes;
    uint8_t* smem_b = smem + kAStages * a_stage_bytes;
    uint8_t* smem_out = smem_b + BSTAGES * kBBytes;            // 4 warps x 2 x [32 rows][64 B] staging tiles
    float* s_stat = (float*)(smem_out + 4 * 2 * 2048);          // sum[kStatMaxN], sumsq[kStatMaxN]
    uint64_t* a_full = (uint64_t*)(s_stat + 2 * kStatMaxN);
    if (p.stat_parts)
        for (int i = threadIdx.x; i < 2 * kStatMaxN; i += kWThreads) s_stat[i] = 0.f;
    uint64_t* a_empty = a_full + kAStages;
    uint64_t* b_full = a_empty + kAStages;
    uint64_t* b_empty = b_full + BSTAGES;
    uint64_t* acc_full = b_empty + BSTAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
            for (int s = 0; s < BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
            mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
            mbar_init(&acc_empty[0], 4); mbar_init(&acc_empty[1], 4);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();     // everything above overlapped the previous kernel's tail; global memory is touched only from here on

    if (warp == 0) {
        // ===== halo producer: per (tile, k-block) two 4-D boxes {64 ch, Wh, Hh, 1 image} =====
        if (elect_one()) {
            uint32_t ga = 0;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x) {
                const WTile t = wdecode(p, q, BN);
                for (int kb = 0; kb < p.kblocks; ++kb, ++ga) {
                    const int s = ga % kAStages;
                    mbar_wait(&a_empty[s], ((ga / kAStages) & 1) ^ 1);
                    uint8_t* sa = smem + s * a_stage_bytes;
                    mbar_expect_tx(&a_full[s], 2 * p.Hh * p.Wh * 128);
                    tma_load_4d(sa, &mapA, &a_full[s], kb * 64, t.gx0 + p.txmin, t.gy0 + p.tymin, t.n);
                    tma_load_4d(sa + p.halo_bytes, &mapA, &a_full[s], kb * 64, t.gx0 + kBrickW + p.txmin, t.gy0 + p.tymin, t.n);
                }
            }
        }
    } else if (warp == 6) {
        // ===== weight producer: one {64, BN, 1} box per (k-block, tap) =====
        if (elect_one()) {
            uint32_t gb = 0;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x) {
                const WTile t = wdecode(p, q, BN);
                const TapList& taps = p.ph[t.phase].taps;
                for (int kb = 0; kb < p.kblocks; ++kb)
                    for (int tp = 0; tp < taps.ntaps; ++tp, ++gb) {
                        const int s = gb % BSTAGES;
                        mbar_wait(&b_empty[s], ((gb / BSTAGES) & 1) ^ 1);
                        mbar_expect_tx(&b_full[s], kBBytes);
                        if constexpr (BMN) {
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j)
                                tma_load_3d(smem_b + s * kBBytes + j * 8192, &mapB, &b_full[s], t.col0 + 64 * j, kb * 64, taps.widx[tp]);
                        } else {
                            tma_load_3d(smem_b + s * kBBytes, &mapB, &b_full[s], kb * 64, t.col0, taps.widx[tp]);
                        }
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = idesc_bf16_f32(128, BN < 16 ? 16 : BN) | (BMN ? (1u << 16) : 0u);
        if (elect_one()) {
            uint32_t ga = 0, gb = 0, i = 0;
            const uint64_t sbo_field = (uint64_t)((p.Wh * 128) >> 4) << 32;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++i) {
                const WTile t = wdecode(p, q, BN);
                const TapList& taps = p.ph[t.phase].taps;
                const uint32_t buf = i & 1, use = i >> 1;
                mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * 2 * kAccCols;
                for (int kb = 0; kb < p.kblocks; ++kb, ++ga) {
                    const int sa_i = ga % kAStages;
                    mbar_wait(&a_full[sa_i], (ga / kAStages) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + sa_i * a_stage_bytes);
                    for (int tp = 0; tp < taps.ntaps; ++tp, ++gb) {
                        const int sb_i = gb % BSTAGES;
                        mbar_wait(&b_full[sb_i], (gb / BSTAGES) & 1);
                        tc_fence_after();
                        const uint64_t bdesc = BMN ? wdesc_mn_sw128(smem_u32(smem_b + sb_i * kBBytes)) : smem_desc_k_sw128(smem_u32(smem_b + sb_i * kBBytes));
                        // window of the halo: first row (ty-tymin)*Wh + (tx-txmin); 16 groups of 8 rows, group stride Wh rows
                        const uint32_t woff = (uint32_t)((taps.ty[tp] - p.tymin) * p.Wh + (taps.tx[tp] - p.txmin)) * 128u;
#pragma unroll
                        for (int br = 0; br < 2; ++br) {
                            const uint32_t a_addr = sa + br * p.halo_bytes + woff;
                            uint64_t adesc = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | sbo_field | ((uint64_t)1 << 46) |
                                             ((uint64_t)2 << 61);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_mma_bf16(tmem_d + br * kAccCols, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * (BMN ? 128 : 2)), idesc,
                                            (kb | tp | k) != 0);
                        }
                        tc_commit(&b_empty[sb_i]);
                    }
                    tc_commit(&a_empty[sa_i]);
                }
                tc_commit(&acc_full[buf]);
            }
        }
    } else if (warp >= 2 && warp <= 5) {
        // ===== epilogue: TMEM lane r = pixel (r/8, r%8) of the brick =====
        const int lane_base = (warp & 3) * 32;
        const int r = lane_base + lane;
        const int by = r >> 3, bx = r & 7;
        uint32_t i = 0, sg = 0;
        for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++i) {
            const WTile t = wdecode(p, q, BN);
            const WinPhase& ph = p.ph[t.phase];
            const uint32_t buf = i & 1, use = i >> 1;
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
#pragma unroll 1
            for (int br = 0; br < 2; ++br) {
                const int gy = t.gy0 + by, gx = t.gx0 + br * kBrickW + bx;
                const int oy = gy * p.ds + ph.doy, ox = gx * p.ds + ph.dox;
                const bool row_ok = gy < ph.gh && gx < ph.gw && oy < p.hd && ox < p.wd;
                const int64_t row_off = (((int64_t)t.n * p.hd + oy) * p.wd + ox) * p.N;
                constexpr int CH = BN >= 32 ? 32 : 16;
                if (BN >= 32 && p.tma_store) {
                    uint8_t* my_stage = smem_out + (warp & 3) * 4096;
                    const uint32_t row_mask = __ballot_sync(0xffffffffu, row_ok);
#pragma unroll 1
                    for (int c = 0; c < BN; c += 32, ++sg) {
                        if (t.col0 + c >= p.N) break;
                        uint8_t* st = my_stage + (sg & 1) * 2048;
                        uint32_t v[32];
                        tmem_ld32(tmem_base + (buf * 2 + br) * kAccCols + ((uint32_t)lane_base << 16) + (uint32_t)c, v);
                        if (lane == 0) tma_store_wait_read<1>();       // the store that last read this buffer has drained it
                        tmem_ld_wait();
                        __syncwarp();
                        stage_chunk32_sw64(st, lane, t.col0 + c, v, p.bias, p.act, p.slope);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (p.stat_parts) stats_chunk32_sw64(st, lane, s_stat + t.col0 + c, s_stat + kStatMaxN + t.col0 + c, row_mask);
                        if (lane == 0) {
                            tma_store_4d(&omaps.m[t.phase], st, t.col0 + c, t.gx0 + br * kBrickW, t.gy0 + (warp & 3) * 4, t.n);
                            tma_store_commit();
                        }
                    }
                } else {
#pragma unroll 1
                for (int c = 0; c < BN; c += CH) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + (buf * 2 + br) * kAccCols + ((uint32_t)lane_base << 16) + (uint32_t)c;
                    if (CH == 32) tmem_ld32(taddr, v);
                    else tmem_ld16(taddr, v);
                    tmem_ld_wait();
                    if (row_ok) store_chunk<CH>(v, p.D, row_off, t.col0 + c, p.N, p.bias, p.act, p.slope, p.out_f32 != 0);
                }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[buf])) : "memory");
        }
        if (p.tma_store && lane == 0) tma_store_wait_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
    if (p.stat_parts) {
        float* out = p.stat_parts + (size_t)blockIdx.x * 2 * p.N;
        for (int i = threadIdx.x; i < 2 * p.N; i += kWThreads) out[i] = s_stat[(i < p.N) ? i : (kStatMaxN + i - p.N)];
    }
}

template <int BN, int BSTAGES, bool BMN = false>
int launch_win(const CUtensorMap& mA, const CUtensorMap& mB, const OutMaps& om, const WinParams& wp, cudaStream_t s) {
    const int smem_bytes = smem_for_occupancy(kAStages * 2 * wp.halo_bytes + BSTAGES * BN * 128 + 4 * 2 * 2048 + 2 * kStatMaxN * 4 +
                                              (2 * kAStages + 2 * BSTAGES + 4) * 8 + 16 + 1024, 1);
    if (smem_bytes > 227 * 1024) { set_error("windowed tap GEMM: %d bytes of shared memory", smem_bytes); return VP_EUNSUPPORTED; }
    static int attr_set = 0;
    if (attr_set < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(tapgemm_win_kernel<BN, BSTAGES, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) { set_error("tapgemm_win: cannot set %d bytes of dynamic smem: %s", smem_bytes, cudaGetErrorString(e)); return VP_ECUDA; }
        attr_set = smem_bytes;
    }
    const int grid = wp.total_tiles < num_sms() ? wp.total_tiles : num_sms();
    launch_k(tapgemm_win_kernel<BN, BSTAGES, BMN>, dim3(grid), dim3(kWThreads), smem_bytes, s, mA, mB, om, wp);
    VP_CHECK_LAUNCH("tapgemm_win");
    return VP_OK;
}

}  // namespace

// Returns VP_EUNSUPPORTED when the problem is not of the windowed form (caller falls back to tapgemm_tc).
int launch_tapgemm_win(const TapGemm* phases, int nphases, cudaStream_t s) {
    const TapGemm& p = phases[0];
    if (!tc_available() || nphases < 1 || nphases > kMaxPh) return VP_EUNSUPPORTED;
    if (p.as != 1 || p.K % 64 != 0 || p.N < 1 || p.n <= 0) return VP_EUNSUPPORTED;
    if (((uintptr_t)p.A & 15) || ((uintptr_t)p.Wp & 15) || ((uintptr_t)p.D & 15)) return VP_EUNSUPPORTED;
    int tymin = 1 << 20, tymax = -(1 << 20), txmin = 1 << 20, txmax = -(1 << 20), gh = 0, gw = 0;
    for (int i = 0; i < nphases; ++i) {
        const TapList& t = phases[i].taps;
        if (t.ntaps < 1) return VP_EUNSUPPORTED;
        for (int j = 0; j < t.ntaps; ++j) {
            tymin = t.ty[j] < tymin ? t.ty[j] : tymin; tymax = t.ty[j] > tymax ? t.ty[j] : tymax;
            txmin = t.tx[j] < txmin ? t.tx[j] : txmin; txmax = t.tx[j] > txmax ? t.tx[j] : txmax;
        }
        gh = phases[i].gh > gh ? phases[i].gh : gh;
        gw = phases[i].gw > gw ? phases[i].gw : gw;
    }
    if (tymax - tymin > 4 || txmax - txmin > 4) return VP_EUNSUPPORTED;       // halo up to 20 x 12 pixels
    if (gh < 12 || gw < 12) return VP_EUNSUPPORTED;                           // bricks of 16 x 8 would be mostly padding
    WinParams wp;
    memset(&wp, 0, sizeof(wp));
    wp.tymin = tymin; wp.txmin = txmin;
    wp.Hh = kBrickH + (tymax - tymin); wp.Wh = kBrickW + (txmax - txmin);
    wp.halo_bytes = (wp.Hh * wp.Wh * 128 + 1023) & ~1023;
    const int BN = (p.N % 128 == 0) ? 128 : (p.N >= 64 ? 64 : (p.N > 16 ? 32 : 16));
    wp.ntiles_n = (p.N + BN - 1) / BN;
    int64_t mtiles = 0;
    for (int i = 0; i < nphases; ++i) {
        WinPhase& ph = wp.ph[i];
        ph.gh = phases[i].gh; ph.gw = phases[i].gw; ph.doy = phases[i].doy; ph.dox = phases[i].dox; ph.taps = phases[i].taps;
        ph.tiles_w = (ph.gw + 2 * kBrickW - 1) / (2 * kBrickW);
        ph.tiles_h = (ph.gh + kBrickH - 1) / kBrickH;
        ph.tile_begin = (int)mtiles;
        mtiles += (int64_t)ph.tiles_w * ph.tiles_h * p.n;
    }
    if (mtiles * wp.ntiles_n > 0x7fffffff) return VP_EUNSUPPORTED;
    wp.total_tiles = (int)(mtiles * wp.ntiles_n);
    wp.nphases = nphases;
    wp.D = p.D; wp.bias = p.bias; wp.n = p.n; wp.hd = p.hd; wp.wd = p.wd; wp.N = p.N; wp.ds = p.ds;
    wp.act = p.act; wp.slope = p.slope; wp.out_f32 = (p.out_dtype == VP_F32); wp.kblocks = p.K / 64;

    EncodeTiledFn encode = get_encode();
    CUtensorMap mA, mB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)p.K, (cuuint64_t)p.wa, (cuuint64_t)p.ha, (cuuint64_t)p.n};
        cuuint64_t strides[3] = {(cuuint64_t)p.K * 2, (cuuint64_t)p.wa * p.K * 2, (cuuint64_t)p.ha * p.wa * p.K * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)wp.Wh, (cuuint32_t)wp.Hh, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (encode(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.A), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return VP_EUNSUPPORTED;
    }
    const bool bmn = p.w_sn == 1 && p.w_sk != 1;
    if ((!bmn && p.w_sk != 1) || (bmn && (p.N % 64 != 0 || BN < 64))) return VP_EUNSUPPORTED;
    if (encode_weight_map(&mB, p, bmn, BN)) return VP_EUNSUPPORTED;
    OutMaps om;
    memset(&om, 0, sizeof(om));
    wp.tma_store = 0; wp.stat_parts = nullptr;
    if (!wp.out_f32 && p.N % 64 == 0 && BN >= 32) {
        bool ok = true;
        for (int i = 0; i < nphases && ok; ++i)
            ok = encode_out_map(&om.m[i], p.D, p.N, p.hd, p.wd, p.n, p.ds, phases[i].doy, phases[i].dox, kBrickW, 4, 1) == 0;
        if (ok) {
            wp.tma_store = 1;
            if (p.stat_parts && p.N <= kStatMaxN && p.bias == nullptr && p.act == VP_ACT_NONE) wp.stat_parts = p.stat_parts;
        }
    }
    if (p.stat_parts && !wp.stat_parts) return VP_EUNSUPPORTED;      // the caller's other engine reports the error
    if (wp.stat_parts) {
        const int g = wp.total_tiles < num_sms() ? wp.total_tiles : num_sms();
        if (g > p.stat_capacity) { set_error("windowed tap GEMM: statistics buffer holds %d parts, %d needed", p.stat_capacity, g); return VP_EINVAL; }
        if (p.stat_nparts) *p.stat_nparts = g;
    }
    if (bmn) return BN == 128 ? launch_win<128, 4, true>(mA, mB, om, wp, s) : launch_win<64, 6, true>(mA, mB, om, wp, s);
    switch (BN) {
        case 128: return launch_win<128, 4>(mA, mB, om, wp, s);
        case 64: return launch_win<64, 6>(mA, mB, om, wp, s);
        case 32: return launch_win<32, 8>(mA, mB, om, wp, s);
        default: return launch_win<16, 8>(mA, mB, om, wp, s);
    }
}

}  // namespace vp
