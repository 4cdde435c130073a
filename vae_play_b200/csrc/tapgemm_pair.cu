// Phase-pair tap GEMM (tcgen05): the stride-2 transposed conv forward / stride-2 conv data gradient with 64 output
// channels (the last DecoderBlock, the second EncoderBlock's dgrad), two output-parity phases per tile.
//
// An M128 x N64 x K16 SS-mode MMA reads 4 KB (A) + 2 KB (B) of shared memory for 32 issue cycles -- 192 B/clk against a
// 128 B/clk port -- so tapgemm_win_kernel<64> runs the tensor pipe at ~33 % (ncu: profiles/r01_ncu_prof_ct3_fwd_r1d).
// The two phases (py, 0) and (py, 1) of an output row parity read the SAME input windows for most of their taps (5x5, s2:
// column shifts {1, 0} are shared, {-1} belongs to px = 0 only), with different weights.  Here a tile computes both
// phases: a shared shift is ONE MMA with N = 128 whose B operand is the two phases' weight tiles side by side
// (MN-major, LBO = one 8 KB box) and whose accumulator is [phase 0 | phase 1]; A is read once for both.
// Per k-block and input row shift: 2 x N128 + 1 x N64 instead of 5 x N64 -> 137 B/clk.
//
// Warp roles as tapgemm_win_kernel: warp 0 halo producer, warp 6 weight producer, warp 1 MMA issuer (double-buffered
// TMEM: 2 buffers x 2 bricks x 2 phases x 64 columns = 512), warps 2..5 epilogue (staging tile -> BatchNorm partial sums
// -> TMA store through the phase's output map).
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace vp {
namespace {

using namespace tc;

constexpr int kPThreads = 224;
constexpr int kBrickH = 16, kBrickW = 8;
constexpr int kAStages = 2, kBStages = 4;
constexpr int kBStage = 16384;             // two 64 (n) x 64 (k) weight boxes
constexpr int kMaxEnt = 12;

struct PairInfo {
    int gh, gw[2];             // iteration grid of the two phases (same rows, columns may differ by one)
    int doy, dox[2];
    int tiles_w, tiles_h, tile_begin;
    int nent;
    int8_t ty[kMaxEnt], tx[kMaxEnt], mode[kMaxEnt];      // mode 0: both phases, 1: phase 0 only, 2: phase 1 only
    int8_t widx0[kMaxEnt], widx1[kMaxEnt];
};

struct PairParams {
    const float* bias;
    int n, hd, wd, N, ds;
    int act;
    float slope;
    int kblocks;
    int tymin, txmin, Hh, Wh, halo_bytes;
    int total_tiles;
    float* stat_parts;
    PairInfo pr[2];
};

struct PTile { int pair, n, gy0, gx0; };

__device__ __forceinline__ PTile pdecode(const PairParams& p, int q) {
    PTile c;
    c.pair = q >= p.pr[1].tile_begin ? 1 : 0;
    int mt = q - p.pr[c.pair].tile_begin;
    const int tw = mt % p.pr[c.pair].tiles_w; mt /= p.pr[c.pair].tiles_w;
    const int th = mt % p.pr[c.pair].tiles_h; mt /= p.pr[c.pair].tiles_h;
    c.n = mt; c.gy0 = th * kBrickH; c.gx0 = tw * (2 * kBrickW);
    return c;
}

__global__ void __launch_bounds__(kPThreads, 1) tapgemm_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                    const __grid_constant__ OutMaps omaps, const __grid_constant__ PairParams p) {
    constexpr int kTmemCols = 512;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_stage_bytes = 2 * p.halo_bytes;
    uint8_t* smem_b = smem + kAStages * a_stage_bytes;
    uint8_t* smem_out = smem_b + kBStages * kBStage;             // 4 warps x 2 x [32 rows][64 B] staging tiles
    float* s_stat = (float*)(smem_out + 4 * 2 * 2048);            // sum[64], sumsq[64]
    uint64_t* a_full = (uint64_t*)(s_stat + 128);
    uint64_t* a_empty = a_full + kAStages;
    uint64_t* b_full = a_empty + kAStages;
    uint64_t* b_empty = b_full + kBStages;
    uint64_t* acc_full = b_empty + kBStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
    if (threadIdx.x < 128) s_stat[threadIdx.x] = 0.f;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
            for (int s = 0; s < kBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
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
                const PTile t = pdecode(p, q);
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
        // ===== weight producer: per (k-block, entry) one or two {64 n, 64 k} boxes of the channels-last weight =====
        if (elect_one()) {
            uint32_t gb = 0;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x) {
                const PTile t = pdecode(p, q);
                const PairInfo& pi = p.pr[t.pair];
                for (int kb = 0; kb < p.kblocks; ++kb)
                    for (int e = 0; e < pi.nent; ++e, ++gb) {
                        const int s = gb % kBStages;
                        mbar_wait(&b_empty[s], ((gb / kBStages) & 1) ^ 1);
                        uint8_t* sb = smem_b + s * kBStage;
                        const int mode = pi.mode[e];
                        mbar_expect_tx(&b_full[s], mode == 0 ? 16384 : 8192);
                        if (mode == 0) {
                            tma_load_3d(sb, &mapB, &b_full[s], 0, kb * 64, pi.widx0[e]);
                            tma_load_3d(sb + 8192, &mapB, &b_full[s], 0, kb * 64, pi.widx1[e]);
                        } else {
                            tma_load_3d(sb, &mapB, &b_full[s], 0, kb * 64, mode == 1 ? pi.widx0[e] : pi.widx1[e]);
                        }
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc128 = idesc_bf16_f32(128, 128) | (1u << 16);
        constexpr uint32_t idesc64 = idesc_bf16_f32(128, 64) | (1u << 16);
        if (elect_one()) {
            uint32_t ga = 0, gb = 0, i = 0;
            const uint64_t sbo_field = (uint64_t)((p.Wh * 128) >> 4) << 32;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++i) {
                const PTile t = pdecode(p, q);
                const PairInfo& pi = p.pr[t.pair];
                const uint32_t buf = i & 1, use = i >> 1;
                mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * 256;
                for (int kb = 0; kb < p.kblocks; ++kb, ++ga) {
                    const int sa_i = ga % kAStages;
                    mbar_wait(&a_full[sa_i], (ga / kAStages) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + sa_i * a_stage_bytes);
                    for (int e = 0; e < pi.nent; ++e, ++gb) {
                        const int sb_i = gb % kBStages;
                        mbar_wait(&b_full[sb_i], (gb / kBStages) & 1);
                        tc_fence_after();
                        const uint32_t sb = smem_u32(smem_b + sb_i * kBStage);
                        // MN-major SW128 weight operand: 64-wide N groups one box (8192 B) apart, K atoms every 1024 B
                        const uint64_t bdesc = (uint64_t)((sb & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
                                               ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
                        const int mode = pi.mode[e];
                        const uint32_t idesc = mode == 0 ? idesc128 : idesc64;
                        const uint32_t col = mode == 2 ? 64u : 0u;
                        const uint32_t woff = (uint32_t)((pi.ty[e] - p.tymin) * p.Wh + (pi.tx[e] - p.txmin)) * 128u;
#pragma unroll
                        for (int br = 0; br < 2; ++br) {
                            const uint32_t a_addr = sa + br * p.halo_bytes + woff;
                            const uint64_t adesc = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | sbo_field | ((uint64_t)1 << 46) |
                                                   ((uint64_t)2 << 61);
#pragma unroll
                            for (int k = 0; k < 4; ++k)      // entries shared by both phases come first: entry 0 initialises both accumulators
                                tc_mma_bf16(tmem_d + br * 128 + col, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 128), idesc, (kb | e | k) != 0);
                        }
                        tc_commit(&b_empty[sb_i]);
                    }
                    tc_commit(&a_empty[sa_i]);
                }
                tc_commit(&acc_full[buf]);
            }
        }
    } else if (warp >= 2 && warp <= 5) {
        // ===== epilogue: TMEM lane r = pixel (r/8, r%8) of the brick; 2 bricks x 2 phases x 64 columns per tile =====
        const int lane_base = (warp & 3) * 32;
        const int r = lane_base + lane;
        const int by = r >> 3, bx = r & 7;
        uint8_t* my_stage = smem_out + (warp & 3) * 4096;
        uint32_t i = 0, sg = 0;
        for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++i) {
            const PTile t = pdecode(p, q);
            const PairInfo& pi = p.pr[t.pair];
            const uint32_t buf = i & 1, use = i >> 1;
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
#pragma unroll 1
            for (int bp = 0; bp < 4; ++bp) {
                const int br = bp >> 1, px = bp & 1;
                const int gy = t.gy0 + by, gx = t.gx0 + br * kBrickW + bx;
                const int oy = gy * p.ds + pi.doy, ox = gx * p.ds + pi.dox[px];
                const bool row_ok = gy < pi.gh && gx < pi.gw[px] && oy < p.hd && ox < p.wd;
                const uint32_t row_mask = __ballot_sync(0xffffffffu, row_ok);
#pragma unroll 1
                for (int c = 0; c < 64; c += 32, ++sg) {
                    uint8_t* st = my_stage + (sg & 1) * 2048;
                    uint32_t v[32];
                    tmem_ld32(tmem_base + buf * 256 + br * 128 + px * 64 + ((uint32_t)lane_base << 16) + (uint32_t)c, v);
                    if (lane == 0) tma_store_wait_read<1>();
                    tmem_ld_wait();
                    __syncwarp();
                    stage_chunk32_sw64(st, lane, c, v, p.bias, p.act, p.slope);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (p.stat_parts) stats_chunk32_sw64(st, lane, s_stat + c, s_stat + 64 + c, row_mask);
                    if (lane == 0) {
                        tma_store_4d(&omaps.m[t.pair * 2 + px], st, c, t.gx0 + br * kBrickW, t.gy0 + (warp & 3) * 4, t.n);
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
        float* out = p.stat_parts + (size_t)blockIdx.x * 128;
        for (int i = threadIdx.x; i < 128; i += kPThreads) out[i] = s_stat[i];
    }
}

}  // namespace

// VP_EUNSUPPORTED unless: 4 output-parity phases (stride 2) in (py, px) order, 64 output channels, bf16 output, MN-major
// (channels-last, in-place) weights, K a multiple of 64, grids of at least 12 x 12.
int launch_tapgemm_pair(const TapGemm* phases, int nphases, cudaStream_t s) {
    const TapGemm& p = phases[0];
    if (!tc_available() || nphases != 4 || getenv("VP_NO_PAIR")) { set_error("pair tap GEMM: not eligible (check %d)", 1); return VP_EUNSUPPORTED; }
    if (p.as != 1 || p.ds != 2 || p.N != 64 || p.K % 64 != 0 || p.n <= 0 || p.out_dtype != VP_BF16) { set_error("pair tap GEMM: not eligible (check %d)", 2); return VP_EUNSUPPORTED; }
    if (!(p.w_sn == 1 && p.w_sk != 1)) { set_error("pair tap GEMM: not eligible (check %d)", 3); return VP_EUNSUPPORTED; }
    if (((uintptr_t)p.A & 15) || ((uintptr_t)p.Wp & 15) || ((uintptr_t)p.D & 15)) { set_error("pair tap GEMM: not eligible (check %d)", 4); return VP_EUNSUPPORTED; }
    for (int i = 0; i < 4; ++i)
        if (phases[i].doy != (i >> 1) || phases[i].dox != (i & 1) || phases[i].taps.ntaps < 1) { set_error("pair tap GEMM: not eligible (check %d)", 5); return VP_EUNSUPPORTED; }
    int tymin = 1 << 20, tymax = -(1 << 20), txmin = 1 << 20, txmax = -(1 << 20), gh = 0, gw = 0;
    for (int i = 0; i < 4; ++i) {
        const TapList& t = phases[i].taps;
        for (int j = 0; j < t.ntaps; ++j) {
            tymin = t.ty[j] < tymin ? t.ty[j] : tymin; tymax = t.ty[j] > tymax ? t.ty[j] : tymax;
            txmin = t.tx[j] < txmin ? t.tx[j] : txmin; txmax = t.tx[j] > txmax ? t.tx[j] : txmax;
        }
        gh = phases[i].gh > gh ? phases[i].gh : gh;
        gw = phases[i].gw > gw ? phases[i].gw : gw;
    }
    if (tymax - tymin > 4 || txmax - txmin > 4 || gh < 12 || gw < 12) { set_error("pair tap GEMM: not eligible (check %d)", 6); return VP_EUNSUPPORTED; }
    PairParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.tymin = tymin; pp.txmin = txmin;
    pp.Hh = kBrickH + (tymax - tymin); pp.Wh = kBrickW + (txmax - txmin);
    pp.halo_bytes = (pp.Hh * pp.Wh * 128 + 1023) & ~1023;
    int64_t tiles = 0;
    for (int pr = 0; pr < 2; ++pr) {
        PairInfo& pi = pp.pr[pr];
        const TapGemm &a = phases[2 * pr], &b = phases[2 * pr + 1];
        if (a.gh != b.gh) { set_error("pair tap GEMM: not eligible (check %d)", 7); return VP_EUNSUPPORTED; }
        pi.gh = a.gh; pi.gw[0] = a.gw; pi.gw[1] = b.gw; pi.doy = a.doy; pi.dox[0] = a.dox; pi.dox[1] = b.dox;
        const int gwm = a.gw > b.gw ? a.gw : b.gw;
        pi.tiles_w = (gwm + 2 * kBrickW - 1) / (2 * kBrickW);
        pi.tiles_h = (pi.gh + kBrickH - 1) / kBrickH;
        pi.tile_begin = (int)tiles;
        tiles += (int64_t)pi.tiles_w * pi.tiles_h * p.n;
        // merged entries: shifts used by both phases first
        int ne = 0;
        for (int pass = 0; pass < 3; ++pass)
            for (int ja = 0; ja < (pass == 2 ? 0 : a.taps.ntaps); ++ja) {
                int jb = -1;
                for (int k = 0; k < b.taps.ntaps; ++k)
                    if (b.taps.ty[k] == a.taps.ty[ja] && b.taps.tx[k] == a.taps.tx[ja]) jb = k;
                if ((pass == 0) != (jb >= 0)) continue;
                if (ne >= kMaxEnt) { set_error("pair tap GEMM: not eligible (check %d)", 8); return VP_EUNSUPPORTED; }
                pi.ty[ne] = a.taps.ty[ja]; pi.tx[ne] = a.taps.tx[ja];
                pi.mode[ne] = pass == 0 ? 0 : 1;
                pi.widx0[ne] = a.taps.widx[ja]; pi.widx1[ne] = jb >= 0 ? b.taps.widx[jb] : 0;
                ++ne;
            }
        for (int kb_ = 0; kb_ < b.taps.ntaps; ++kb_) {       // taps of phase 1 that phase 0 does not have
            bool shared = false;
            for (int k = 0; k < a.taps.ntaps; ++k)
                if (a.taps.ty[k] == b.taps.ty[kb_] && a.taps.tx[k] == b.taps.tx[kb_]) shared = true;
            if (shared) continue;
            if (ne >= kMaxEnt) { set_error("pair tap GEMM: not eligible (check %d)", 9); return VP_EUNSUPPORTED; }
            pi.ty[ne] = b.taps.ty[kb_]; pi.tx[ne] = b.taps.tx[kb_]; pi.mode[ne] = 2; pi.widx0[ne] = 0; pi.widx1[ne] = b.taps.widx[kb_];
            ++ne;
        }
        pi.nent = ne;
        if (ne == 0 || pi.mode[0] != 0) { set_error("pair tap GEMM: not eligible (check %d)", 10); return VP_EUNSUPPORTED; }      // the first entry must initialise both accumulators
    }
    if (tiles > 0x7fffffff) { set_error("pair tap GEMM: not eligible (check %d)", 11); return VP_EUNSUPPORTED; }
    pp.total_tiles = (int)tiles;
    pp.bias = p.bias; pp.n = p.n; pp.hd = p.hd; pp.wd = p.wd; pp.N = p.N; pp.ds = p.ds; pp.act = p.act; pp.slope = p.slope;
    pp.kblocks = p.K / 64;

    EncodeTiledFn encode = get_encode();
    CUtensorMap mA, mB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)p.K, (cuuint64_t)p.wa, (cuuint64_t)p.ha, (cuuint64_t)p.n};
        cuuint64_t strides[3] = {(cuuint64_t)p.K * 2, (cuuint64_t)p.wa * p.K * 2, (cuuint64_t)p.ha * p.wa * p.K * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)pp.Wh, (cuuint32_t)pp.Hh, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (encode(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.A), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            { set_error("pair tap GEMM: not eligible (check %d)", 12); return VP_EUNSUPPORTED; }
    }
    if (encode_weight_map(&mB, p, true, 64)) { set_error("pair tap GEMM: not eligible (check %d)", 13); return VP_EUNSUPPORTED; }
    OutMaps om;
    memset(&om, 0, sizeof(om));
    for (int i = 0; i < 4; ++i)
        if (encode_out_map(&om.m[i], p.D, p.N, p.hd, p.wd, p.n, p.ds, phases[i].doy, phases[i].dox, kBrickW, 4, 1)) { set_error("pair tap GEMM: not eligible (check %d)", 14); return VP_EUNSUPPORTED; }
    const int grid = pp.total_tiles < num_sms() ? pp.total_tiles : num_sms();
    pp.stat_parts = nullptr;
    if (p.stat_parts) {
        if (p.bias || p.act != VP_ACT_NONE) { set_error("pair tap GEMM: not eligible (check %d)", 15); return VP_EUNSUPPORTED; }
        if (grid > p.stat_capacity) { set_error("pair tap GEMM: statistics buffer holds %d parts, %d needed", p.stat_capacity, grid); return VP_EINVAL; }
        if (p.stat_nparts) *p.stat_nparts = grid;
        pp.stat_parts = p.stat_parts;
    }
    const int smem_bytes = smem_for_occupancy(kAStages * 2 * pp.halo_bytes + kBStages * kBStage + 4 * 2 * 2048 + 128 * 4 + (2 * kAStages + 2 * kBStages + 4) * 8 + 16 + 1024, 1);
    if (smem_bytes > 227 * 1024) { set_error("pair tap GEMM: not eligible (check %d)", 16); return VP_EUNSUPPORTED; }
    static int attr_set = 0;
    if (attr_set < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(tapgemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) { set_error("tapgemm_pair: cannot set %d bytes of dynamic smem: %s", smem_bytes, cudaGetErrorString(e)); return VP_ECUDA; }
        attr_set = smem_bytes;
    }
    launch_k(tapgemm_pair_kernel, dim3(grid), dim3(kPThreads), smem_bytes, s, mA, mB, om, pp);
    VP_CHECK_LAUNCH("tapgemm_pair");
    return VP_OK;
}

}  // namespace vp
