// Windowed tap GEMM for the GATHER form with stride 2 (stride-2 conv forward, stride-2 transposed-conv data gradient):
//   D[n, gy, gx, :] = sum_t A[n, 2*gy + ty_t, 2*gx + tx_t, :] . W[widx_t]
// tapgemm_tc_kernel streams a fresh 128 x 64 A tile per tap (25 per k-block): 125 B/clk/SM of L2->SMEM traffic, which is
// what bounds those layers.  Here the taps are grouped by stride parity (ty mod 2, tx mod 2): all taps of a class read the
// same sub-lattice A[2i + r][2j + c] shifted by whole pixels, so per (k-block, class) each 16x8-pixel brick loads ONE halo
// box of that sub-lattice (a tensor map whose strides skip every other pixel) and every tap is a window of it -- the same
// mechanism as tapgemm_win_kernel (any 128-byte-aligned start / any 8-row-group stride is a legal K-major SWIZZLE_128B
// operand).  A bytes per k-block and brick: (180 + 162 + 170 + 153) pixels x 128 B = 85 KB instead of 25 x 16 KB.
//
// Warp roles as tapgemm_win_kernel (224 threads, 1 CTA/SM): warp 0 halo producer, warp 6 weight producer, warp 1 MMA
// issuer (double-buffered TMEM: 2 buffers x 2 bricks x BN columns), warps 2..5 epilogue (staging tile -> BatchNorm partial
// sums -> TMA store).
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace vp {
namespace {

using namespace tc;

constexpr int kGThreads = 224;
constexpr int kBrickH = 16, kBrickW = 8;
constexpr int kAStages = 2;
constexpr int kMaxCls = 4;

struct InMaps { CUtensorMap m[kMaxCls]; };

struct GClass {
    int amin, bmin, Hh, Wh;    // halo origin (whole-pixel shift) and extent in sub-lattice pixels
    int t0, t1;                // taps [t0, t1) of the class-sorted tap list
};

struct GwinParams {
    const float* bias;
    int n, gh, gw, N;
    int act;
    float slope;
    int kblocks;
    int halo_bytes;            // per brick, largest class, rounded up to 1024
    int tiles_w, tiles_h, ntiles_n, total_tiles;
    float* stat_parts;
    const float *bn_scale, *bn_shift, *bn_mean;      // non-null: fused BatchNorm-backward reduction (mapY = the block's pre-norm output)
    int ncls;
    GClass cls[kMaxCls];
    int8_t sy[kMaxTaps], sx[kMaxTaps], widx[kMaxTaps];      // window offset inside the class halo, weight tap index
};

struct GTile { int n, gy0, gx0, col0; };

__device__ __forceinline__ GTile gdecode(const GwinParams& p, int q, int BN) {
    GTile c;
    const int nt = q % p.ntiles_n;
    int mt = q / p.ntiles_n;
    const int tw = mt % p.tiles_w; mt /= p.tiles_w;
    const int th = mt % p.tiles_h; mt /= p.tiles_h;
    c.n = mt; c.gy0 = th * kBrickH; c.gx0 = tw * (2 * kBrickW); c.col0 = nt * BN;
    return c;
}

template <int BN, int BSTAGES>
__global__ void __launch_bounds__(kGThreads, 1) tapgemm_gwin_kernel(const __grid_constant__ InMaps mapsA, const __grid_constant__ CUtensorMap mapB,
                                                                    const __grid_constant__ CUtensorMap mapD, const __grid_constant__ CUtensorMap mapY,
                                                                    const __grid_constant__ GwinParams p) {
    constexpr int kBBytes = BN * 128;
    constexpr int kTmemCols = 4 * BN;                    // 2 buffers x 2 bricks
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_stage_bytes = 2 * p.halo_bytes;
    uint8_t* smem_b = smem + kAStages * a_stage_bytes;
    uint8_t* smem_out = smem_b + BSTAGES * kBBytes;              // 4 warps x 2 x [32 rows][64 B] staging tiles
    float* s_stat = (float*)(smem_out + 4 * 2 * 2048);            // sum[kStatMaxN], sumsq[kStatMaxN]
    uint64_t* a_full = (uint64_t*)(s_stat + 2 * kStatMaxN);
    uint64_t* a_empty = a_full + kAStages;
    uint64_t* b_full = a_empty + kAStages;
    uint64_t* b_empty = b_full + BSTAGES;
    uint64_t* acc_full = b_empty + BSTAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* y_full = acc_empty + 2;                            // [4 warps][2]: y tiles of the fused BatchNorm-backward reduction
    uint32_t* tmem_slot = (uint32_t*)(y_full + 8);
    uint8_t* smem_y = (uint8_t*)(((uintptr_t)(tmem_slot + 4) + 1023) & ~(uintptr_t)1023);   // 4 warps x 2 x [32 rows][64 B]
    if (p.stat_parts)
        for (int i = threadIdx.x; i < 2 * kStatMaxN; i += kGThreads) s_stat[i] = 0.f;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int c = 0; c < p.ncls; ++c) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapsA.m[c]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
            for (int s = 0; s < BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
            mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
            mbar_init(&acc_empty[0], 4); mbar_init(&acc_empty[1], 4);
            for (int i = 0; i < 8; ++i) mbar_init(&y_full[i], 1);
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
        // ===== halo producer: per (tile, k-block, class) two boxes {64 ch, Wh_c, Hh_c, 1 image} of the class sub-lattice =====
        if (elect_one()) {
            uint32_t ga = 0;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x) {
                const GTile t = gdecode(p, q, BN);
                for (int kb = 0; kb < p.kblocks; ++kb)
                    for (int c = 0; c < p.ncls; ++c, ++ga) {
                        const GClass& cl = p.cls[c];
                        const int s = ga % kAStages;
                        mbar_wait(&a_empty[s], ((ga / kAStages) & 1) ^ 1);
                        uint8_t* sa = smem + s * a_stage_bytes;
                        mbar_expect_tx(&a_full[s], 2 * cl.Hh * cl.Wh * 128);
                        tma_load_4d(sa, &mapsA.m[c], &a_full[s], kb * 64, t.gx0 + cl.bmin, t.gy0 + cl.amin, t.n);
                        tma_load_4d(sa + p.halo_bytes, &mapsA.m[c], &a_full[s], kb * 64, t.gx0 + kBrickW + cl.bmin, t.gy0 + cl.amin, t.n);
                    }
            }
        }
    } else if (warp == 6) {
        // ===== weight producer: one {64 k, BN n, 1 tap} box per (k-block, class, tap) =====
        if (elect_one()) {
            uint32_t gb = 0;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x) {
                const GTile t = gdecode(p, q, BN);
                for (int kb = 0; kb < p.kblocks; ++kb)
                    for (int c = 0; c < p.ncls; ++c)
                        for (int tp = p.cls[c].t0; tp < p.cls[c].t1; ++tp, ++gb) {
                            const int s = gb % BSTAGES;
                            mbar_wait(&b_empty[s], ((gb / BSTAGES) & 1) ^ 1);
                            mbar_expect_tx(&b_full[s], kBBytes);
                            tma_load_3d(smem_b + s * kBBytes, &mapB, &b_full[s], kb * 64, t.col0, p.widx[tp]);
                        }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = idesc_bf16_f32(128, BN);
        if (elect_one()) {
            uint32_t ga = 0, gb = 0, i = 0;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++i) {
                const uint32_t buf = i & 1, use = i >> 1;
                mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * 2 * BN;
                for (int kb = 0; kb < p.kblocks; ++kb)
                    for (int c = 0; c < p.ncls; ++c, ++ga) {
                        const GClass& cl = p.cls[c];
                        const int sa_i = ga % kAStages;
                        mbar_wait(&a_full[sa_i], (ga / kAStages) & 1);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + sa_i * a_stage_bytes);
                        const uint64_t sbo_field = (uint64_t)((cl.Wh * 128) >> 4) << 32;
                        for (int tp = cl.t0; tp < cl.t1; ++tp, ++gb) {
                            const int sb_i = gb % BSTAGES;
                            mbar_wait(&b_full[sb_i], (gb / BSTAGES) & 1);
                            tc_fence_after();
                            const uint64_t bdesc = smem_desc_k_sw128(smem_u32(smem_b + sb_i * kBBytes));
                            const uint32_t woff = (uint32_t)(p.sy[tp] * cl.Wh + p.sx[tp]) * 128u;
#pragma unroll
                            for (int br = 0; br < 2; ++br) {
                                const uint32_t a_addr = sa + br * p.halo_bytes + woff;
                                const uint64_t adesc = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | sbo_field | ((uint64_t)1 << 46) |
                                                       ((uint64_t)2 << 61);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    tc_mma_bf16(tmem_d + br * BN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | c | (tp - cl.t0) | k) != 0);
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
        uint8_t* my_stage = smem_out + (warp & 3) * 4096;
        uint8_t* my_y = smem_y + (warp & 3) * 4096;
        uint64_t* my_ybar = y_full + (warp & 3) * 2;
        const bool bn = p.bn_scale != nullptr;
        constexpr int kChunksN = BN / 32, kChunks = 2 * kChunksN;       // chunks of a tile: (brick, 32-column group)
        uint32_t i = 0, sg = 0, yk = 0;                                  // yk: y tiles consumed so far (buffer = yk & 1, parity = (yk >> 1) & 1)
        for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++i) {
            const GTile t = gdecode(p, q, BN);
            const uint32_t buf = i & 1, use = i >> 1;
            // y tile of chunk k of this tile -> buffer (yk + k) & 1; two chunks are kept in flight
            auto issue_y = [&](int k, uint32_t slot) {
                const int br = k / kChunksN, c = (k % kChunksN) * 32;
                mbar_expect_tx(&my_ybar[slot], 2048);
                tma_load_4d(my_y + slot * 2048, &mapY, &my_ybar[slot], t.col0 + c, t.gx0 + br * kBrickW, t.gy0 + (warp & 3) * 4, t.n);
            };
            if (bn && lane == 0) { issue_y(0, yk & 1); if (kChunks > 1) issue_y(1, (yk + 1) & 1); }      // (N is a multiple of BN here)
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
            int k = 0;
#pragma unroll 1
            for (int br = 0; br < 2; ++br) {
                const int gy = t.gy0 + by, gx = t.gx0 + br * kBrickW + bx;
                const uint32_t row_mask = __ballot_sync(0xffffffffu, gy < p.gh && gx < p.gw);
#pragma unroll 1
                for (int c = 0; c < BN; c += 32, ++sg, ++k) {
                    if (t.col0 + c >= p.N) break;
                    uint8_t* st = my_stage + (sg & 1) * 2048;
                    uint32_t v[32];
                    tmem_ld32(tmem_base + (buf * 2 + br) * BN + ((uint32_t)lane_base << 16) + (uint32_t)c, v);
                    if (lane == 0) tma_store_wait_read<1>();
                    tmem_ld_wait();
                    __syncwarp();
                    stage_chunk32_sw64(st, lane, t.col0 + c, v, p.bias, p.act, p.slope);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (bn) {
                        const uint32_t slot = yk & 1;
                        mbar_wait(&my_ybar[slot], (yk >> 1) & 1);
                        bnred_chunk32_sw64(st, my_y + slot * 2048, lane, s_stat + t.col0 + c, s_stat + kStatMaxN + t.col0 + c, row_mask,
                                           p.bn_scale + t.col0 + c, p.bn_shift + t.col0 + c, p.bn_mean + t.col0 + c);
                        __syncwarp();                                  // every lane is done with this y buffer before it is refilled
                        if (lane == 0 && k + 2 < kChunks) issue_y(k + 2, slot);
                        ++yk;
                    } else if (p.stat_parts) {
                        stats_chunk32_sw64(st, lane, s_stat + t.col0 + c, s_stat + kStatMaxN + t.col0 + c, row_mask);
                    }
                    if (lane == 0) {
                        tma_store_4d(&mapD, st, t.col0 + c, t.gx0 + br * kBrickW, t.gy0 + (warp & 3) * 4, t.n);
                        tma_store_commit();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[buf])) : "memory");
        }
        if (lane == 0) tma_store_wait_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
    if (p.stat_parts) {
        float* out = p.stat_parts + (size_t)blockIdx.x * 2 * p.N;
        for (int i = threadIdx.x; i < 2 * p.N; i += kGThreads) out[i] = s_stat[(i < p.N) ? i : (kStatMaxN + i - p.N)];
    }
}

template <int BN, int BSTAGES>
int launch_gwin(const InMaps& mA, const CUtensorMap& mB, const CUtensorMap& mD, const CUtensorMap& mY, const GwinParams& gp, cudaStream_t s) {
    // + 8 y barriers, the TMEM slot, and (1 KB-aligned) 4 x 2 x 2 KB of y staging tiles for the fused BatchNorm-backward reduction
    const int smem_bytes = smem_for_occupancy(kAStages * 2 * gp.halo_bytes + BSTAGES * BN * 128 + 4 * 2 * 2048 + 2 * kStatMaxN * 4 +
                                              (2 * kAStages + 2 * BSTAGES + 4 + 8) * 8 + 16 + 1024 + 1024 + 4 * 2 * 2048, 1);
    if (smem_bytes > 227 * 1024) return VP_EUNSUPPORTED;
    static int attr_set = 0;
    if (attr_set < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(tapgemm_gwin_kernel<BN, BSTAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) { set_error("tapgemm_gwin: cannot set %d bytes of dynamic smem: %s", smem_bytes, cudaGetErrorString(e)); return VP_ECUDA; }
        attr_set = smem_bytes;
    }
    const int grid = gp.total_tiles < num_sms() ? gp.total_tiles : num_sms();
    launch_k(tapgemm_gwin_kernel<BN, BSTAGES>, dim3(grid), dim3(kGThreads), smem_bytes, s, mA, mB, mD, mY, gp);
    VP_CHECK_LAUNCH("tapgemm_gwin");
    return VP_OK;
}

int floor_div2(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

}  // namespace

// VP_EUNSUPPORTED unless: gather form with stride 2, bf16 output, K-major weights, K % 64 == 0, N in {64, 128}, taps within
// +-2, output grid of at least 12 x 12 (caller falls back to tapgemm_tc_kernel).
int launch_tapgemm_gwin(const TapGemm& p, cudaStream_t s) {
    if (!tc_available() || getenv("VP_NO_GWIN")) return VP_EUNSUPPORTED;
    if (p.as != 2 || p.ds != 1 || p.doy != 0 || p.dox != 0 || p.K % 64 != 0 || (p.N != 64 && p.N != 128) || p.n <= 0) return VP_EUNSUPPORTED;
    if (p.out_dtype != VP_BF16 || p.w_sk != 1 || p.gh < 12 || p.gw < 12 || p.taps.ntaps < 1 || p.taps.ntaps > kMaxTaps) return VP_EUNSUPPORTED;
    if (p.gh != p.hd || p.gw != p.wd) return VP_EUNSUPPORTED;
    if (((uintptr_t)p.A & 15) || ((uintptr_t)p.Wp & 15) || ((uintptr_t)p.D & 15)) return VP_EUNSUPPORTED;
    GwinParams gp;
    memset(&gp, 0, sizeof(gp));
    InMaps mA;
    memset(&mA, 0, sizeof(mA));
    EncodeTiledFn encode = get_encode();
    int nt = 0, halo_max = 0;
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) {
            int amin = 1 << 20, amax = -(1 << 20), bmin = 1 << 20, bmax = -(1 << 20), cnt = 0;
            for (int t = 0; t < p.taps.ntaps; ++t) {
                const int ty = p.taps.ty[t], tx = p.taps.tx[t];
                if (ty < -2 || ty > 2 || tx < -2 || tx > 2) return VP_EUNSUPPORTED;
                const int a = floor_div2(ty, 2), b = floor_div2(tx, 2);
                if (ty - 2 * a != r || tx - 2 * b != c) continue;
                amin = a < amin ? a : amin; amax = a > amax ? a : amax; bmin = b < bmin ? b : bmin; bmax = b > bmax ? b : bmax;
                ++cnt;
            }
            if (!cnt) continue;
            const int hs = (p.ha - r + 1) / 2, ws = (p.wa - c + 1) / 2;
            if (hs <= 0 || ws <= 0) return VP_EUNSUPPORTED;
            GClass& cl = gp.cls[gp.ncls];
            cl.amin = amin; cl.bmin = bmin; cl.Hh = kBrickH + amax - amin; cl.Wh = kBrickW + bmax - bmin; cl.t0 = nt;
            for (int t = 0; t < p.taps.ntaps; ++t) {
                const int ty = p.taps.ty[t], tx = p.taps.tx[t];
                const int a = floor_div2(ty, 2), b = floor_div2(tx, 2);
                if (ty - 2 * a != r || tx - 2 * b != c) continue;
                gp.sy[nt] = (int8_t)(a - amin); gp.sx[nt] = (int8_t)(b - bmin); gp.widx[nt] = p.taps.widx[t];
                ++nt;
            }
            cl.t1 = nt;
            const int hb = cl.Hh * cl.Wh * 128;
            halo_max = hb > halo_max ? hb : halo_max;
            // sub-lattice (r, c): pixels (2i + r, 2j + c)
            const uint8_t* base = (const uint8_t*)p.A + ((int64_t)r * p.wa + c) * p.K * 2;
            cuuint64_t dims[4] = {(cuuint64_t)p.K, (cuuint64_t)ws, (cuuint64_t)hs, (cuuint64_t)p.n};
            cuuint64_t strides[3] = {(cuuint64_t)2 * p.K * 2, (cuuint64_t)2 * p.wa * p.K * 2, (cuuint64_t)p.ha * p.wa * p.K * 2};
            cuuint32_t box[4] = {64, (cuuint32_t)cl.Wh, (cuuint32_t)cl.Hh, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            if (encode(&mA.m[gp.ncls], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<uint8_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return VP_EUNSUPPORTED;
            ++gp.ncls;
        }
    if (gp.ncls == 0) return VP_EUNSUPPORTED;
    gp.halo_bytes = (halo_max + 1023) & ~1023;
    const int BN = p.N % 128 == 0 ? 128 : 64;
    gp.ntiles_n = (p.N + BN - 1) / BN;
    gp.tiles_w = (p.gw + 2 * kBrickW - 1) / (2 * kBrickW);
    gp.tiles_h = (p.gh + kBrickH - 1) / kBrickH;
    const int64_t total = (int64_t)gp.tiles_w * gp.tiles_h * p.n * gp.ntiles_n;
    // one CTA per SM and ~6 us tiles: with fewer than 4 waves the tail wave costs more than the saved L2 traffic (measured
    // on the second EncoderBlock's forward: 256 tiles, 28.6 us here vs 27.2 us in tapgemm_tc_kernel)
    if (total > 0x7fffffff || total < 4 * (int64_t)num_sms()) return VP_EUNSUPPORTED;
    gp.total_tiles = (int)total;
    gp.bias = p.bias; gp.n = p.n; gp.gh = p.gh; gp.gw = p.gw; gp.N = p.N; gp.act = p.act; gp.slope = p.slope; gp.kblocks = p.K / 64;
    CUtensorMap mB, mD;
    if (encode_weight_map(&mB, p, false, BN)) return VP_EUNSUPPORTED;
    if (encode_out_map(&mD, p.D, p.N, p.hd, p.wd, p.n, 1, 0, 0, kBrickW, 4, 1)) return VP_EUNSUPPORTED;
    gp.stat_parts = nullptr;
    gp.bn_scale = gp.bn_shift = gp.bn_mean = nullptr;
    CUtensorMap mY = mD;
    if (p.bn_y) {
        // fused BatchNorm-backward reduction: the same {32 ch, 8, 4, 1} SWIZZLE_64B boxes are LOADED from the block's pre-norm output
        if (!p.stat_parts || !p.bn_scale || !p.bn_shift || !p.bn_mean || p.N % BN != 0) return VP_EUNSUPPORTED;
        if (encode_out_map(&mY, const_cast<void*>(p.bn_y), p.N, p.hd, p.wd, p.n, 1, 0, 0, kBrickW, 4, 1)) return VP_EUNSUPPORTED;
        gp.bn_scale = p.bn_scale; gp.bn_shift = p.bn_shift; gp.bn_mean = p.bn_mean;
    }
    if (p.stat_parts) {
        if (p.bias || p.act != VP_ACT_NONE || p.N > kStatMaxN) return VP_EUNSUPPORTED;
        const int g = gp.total_tiles < num_sms() ? gp.total_tiles : num_sms();
        if (g > p.stat_capacity) { set_error("gather-window tap GEMM: statistics buffer holds %d parts, %d needed", p.stat_capacity, g); return VP_EINVAL; }
        if (p.stat_nparts) *p.stat_nparts = g;
        gp.stat_parts = p.stat_parts;
    }
    return BN == 128 ? launch_gwin<128, 4>(mA, mB, mD, mY, gp, s) : launch_gwin<64, 6>(mA, mB, mD, mY, gp, s);
}

}  // namespace vp
