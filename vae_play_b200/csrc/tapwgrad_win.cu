// Tap-row weight gradient on tcgen05:  dW[ky, kx][gc][ac] += sum_pix G[pix][gc] * A[pix*s + (ky, kx) - pad][ac]
//
// tapwgrad_tc_kernel gives every filter tap its own CTA, so the G tile and the (heavily overlapping) shifted A tiles are
// streamed from L2 once PER TAP: for the 5x5 layers that is ~2 GB of L2->SMEM traffic per launch and the kernel runs at
// the L2 bandwidth, not at the tensor-core rate (measured 10-14 TB/s, 30-40 % of the MMA peak).  Here a CTA owns one
// filter ROW (fixed ky, all kx) of a 128 (gc) x 64 (ac) weight tile, with one TMEM accumulator per tap (kw x 64 columns):
//   * G (the grid-side tensor) arrives once per 8x8-pixel brick and feeds all kw taps;
//   * the taps kx of one stride-parity class c = (kx - pad) mod s read the same sub-lattice of A, shifted by whole
//     pixels b = floor((kx - pad) / s): ONE halo box {64 ch, 8 + (bmax - bmin), 8 rows} per class is loaded and every tap
//     uses a window of it.  A window is a legal MN-major SWIZZLE_128B operand: an 8-pixel brick row is one 8-row K atom
//     (1024 contiguous bytes starting at any 128-byte multiple), the next brick row is SBO = Wh*128 bytes further, and the
//     hardware swizzle is a function of the absolute shared-memory address, which is also how TMA wrote the halo.
// L2->SMEM bytes per 5-tap row and brick: 16 KB (G) + 19 KB (two halos) instead of 5 x 24 KB.
//
// Warp roles (224 threads, 1 CTA/SM): warp 0 = G producer, warp 6 = halo producer, warp 1 = MMA issuer,
// warps 2..5 = final epilogue (fp32 red.global.add of the kw accumulators; pixel splits combine there).
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace vp {
namespace {

using namespace tc;

constexpr int kRThreads = 224;
constexpr int kGStages = 3;
constexpr int kHStages = 3;
constexpr int kGBytes = 2 * 8192;          // 128 gc x 64 pixels
constexpr int kMaxCls = 4;                 // gather stride <= 4
constexpr int kMaxKw = 8;

struct RowMaps { CUtensorMap m[kMaxCls]; };   // sub-lattice (r, c) of A for the CTA's row parity r, c = 0..s-1

struct RowParams {
    float* dW;
    int64_t o_st, o_sg;
    int GC, AC;
    int tiles_w, tiles_h, nbricks, bricks_per_split;
    int mtiles, ntiles, nrows;
    int kw, s;
    int8_t row_r[kMaxKw], row_a[kMaxKw];          // per filter row ky: sub-lattice parity and whole-pixel shift
    int ncls;
    int8_t cls_ntaps[kMaxCls], cls_bmin[kMaxCls], cls_wh[kMaxCls];
    int8_t cls_kx[kMaxCls][kMaxKw];               // filter column of the j-th tap of class c
    int8_t cls_acc[kMaxCls][kMaxKw];              // its accumulator index (= kx)
    int cls_off[kMaxCls];                         // byte offset of the class halo inside a halo stage
    int halo_stage_bytes;
    // MMA groups: up to 4 taps of one class are consecutive 128-byte shifts of the same halo, i.e. ONE MN-major operand with
    // N = 64*ntaps whose 64-wide N groups sit LBO = 128 bytes apart -> one tcgen05.mma per K step reads G once for all of them
    // (an N = 64 MMA would be bound by the shared-memory read of its 128 x 16 A operand: 6 KB per 32 cycles).
    int ngroups;
    int grp_off[kMaxKw], grp_col[kMaxKw], grp_wh[kMaxKw];
    uint32_t grp_idesc[kMaxKw];
    int8_t acc_kx[kMaxKw];                        // accumulator (64-column slot) -> filter column
};

__host__ __device__ constexpr uint32_t idesc_mn64(int m, int n) { return idesc_bf16_f32(m, n) | (1u << 15) | (1u << 16); }

__global__ void __launch_bounds__(kRThreads, 1) tapwgrad_row_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ RowMaps maps0,
                                                                   const __grid_constant__ RowMaps maps1, const __grid_constant__ RowMaps maps2,
                                                                   const __grid_constant__ RowMaps maps3, const __grid_constant__ RowParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_h = smem + kGStages * kGBytes;
    uint64_t* g_full = (uint64_t*)(smem_h + kHStages * p.halo_stage_bytes);
    uint64_t* g_empty = g_full + kGStages;
    uint64_t* h_full = g_empty + kGStages;
    uint64_t* h_empty = h_full + kHStages;
    uint64_t* acc_ready = h_empty + kHStages;
    uint32_t* tmem_slot = (uint32_t*)(acc_ready + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int id = blockIdx.x;
    const int nt = id % p.ntiles; id /= p.ntiles;
    const int mt = id % p.mtiles; id /= p.mtiles;
    const int ky = id;
    const int b0 = blockIdx.y * p.bricks_per_split;
    const int b1 = min(b0 + p.bricks_per_split, p.nbricks);
    const int iters = b1 - b0;
    if (iters <= 0) return;
    const int m0 = mt * 128, c0 = nt * 64;
    const int rr = p.row_r[ky], ra = p.row_a[ky];
    const RowMaps& maps = rr == 0 ? maps0 : (rr == 1 ? maps1 : (rr == 2 ? maps2 : maps3));
    constexpr int kTmemCols = 512;

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapG) : "memory");
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kGStages; ++s) { mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], 1); }
            for (int s = 0; s < kHStages; ++s) { mbar_init(&h_full[s], 1); mbar_init(&h_empty[s], 1); }
            mbar_init(acc_ready, 1);
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
        // ===== G producer: two {64 gc, 8, 8, 1} boxes per brick =====
        if (elect_one()) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % kGStages;
                mbar_wait(&g_empty[s], ((it / kGStages) & 1) ^ 1);
                int b = b0 + it;
                const int tw = b % p.tiles_w; b /= p.tiles_w;
                const int th = b % p.tiles_h; b /= p.tiles_h;
                uint8_t* sg = smem + s * kGBytes;
                mbar_expect_tx(&g_full[s], kGBytes);
                tma_load_4d(sg, &mapG, &g_full[s], m0, tw * 8, th * 8, b);
                tma_load_4d(sg + 8192, &mapG, &g_full[s], m0 + 64, tw * 8, th * 8, b);
            }
        }
    } else if (warp == 6) {
        // ===== halo producer: one {64 ac, Wh_c, 8, 1} box per parity class of the row =====
        if (elect_one()) {
            uint32_t bytes = 0;
            for (int c = 0; c < p.ncls; ++c) bytes += (uint32_t)p.cls_wh[c] * 8u * 128u;
            for (int it = 0; it < iters; ++it) {
                const int s = it % kHStages;
                mbar_wait(&h_empty[s], ((it / kHStages) & 1) ^ 1);
                int b = b0 + it;
                const int tw = b % p.tiles_w; b /= p.tiles_w;
                const int th = b % p.tiles_h; b /= p.tiles_h;
                uint8_t* sh = smem_h + s * p.halo_stage_bytes;
                mbar_expect_tx(&h_full[s], bytes);
                for (int c = 0; c < p.ncls; ++c)
                    tma_load_4d(sh + p.cls_off[c], &maps.m[c], &h_full[s], c0, tw * 8 + p.cls_bmin[c], th * 8 + ra, b);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            // per-group constants in registers: the issuing thread's instruction stream must stay far below the MMA time
            const int ngroups = p.ngroups;
            uint64_t g_desc[kMaxKw];
            uint32_t g_off[kMaxKw], g_col[kMaxKw], g_kstep[kMaxKw], g_idesc[kMaxKw];
#pragma unroll
            for (int g = 0; g < kMaxKw; ++g) {
                const uint32_t wh = (uint32_t)p.grp_wh[g];
                // MN-major SW128: LBO = 128 B (next tap = window shifted by one pixel), SBO = Wh*128 B (next brick row)
                g_desc[g] = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((wh * 128u) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
                g_off[g] = (uint32_t)p.grp_off[g]; g_col[g] = (uint32_t)p.grp_col[g]; g_kstep[g] = wh * 16u; g_idesc[g] = p.grp_idesc[g];
            }
            for (int it = 0; it < iters; ++it) {
                const int sg_i = it % kGStages, sh_i = it % kHStages;
                mbar_wait(&g_full[sg_i], (it / kGStages) & 1);
                mbar_wait(&h_full[sh_i], (it / kHStages) & 1);
                tc_fence_after();
                const uint32_t sg = smem_u32(smem + sg_i * kGBytes);
                // G: MN-major, M = 128 = two 64-gc boxes (LBO = 8192), K atoms of 8 pixels every 1024 B
                const uint64_t adesc = (uint64_t)((sg & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
                                       ((uint64_t)2 << 61);
                const uint32_t sh = smem_u32(smem_h + sh_i * p.halo_stage_bytes);
#pragma unroll
                for (int g = 0; g < kMaxKw; ++g) {
                    if (g < ngroups) {
                        const uint64_t bdesc = g_desc[g] + (uint64_t)(((sh + g_off[g]) & 0x3FFFF) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)          // 16 pixels = two brick rows: G +2048 B, window +2*Wh*128 B
                            tc_mma_bf16(tmem_base + g_col[g], adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * g_kstep[g]), g_idesc[g], (it | k) != 0);
                    }
                }
                tc_commit(&g_empty[sg_i]);
                tc_commit(&h_empty[sh_i]);
            }
            tc_commit(acc_ready);
        }
    } else if (warp >= 2 && warp <= 5) {
        const int lane_base = (warp & 3) * 32;
        const int gc = m0 + lane_base + lane;
        mbar_wait(acc_ready, 0);
        tc_fence_after();
#pragma unroll 1
        for (int slot = 0; slot < p.kw; ++slot) {
            const int kx = p.acc_kx[slot];
            float* out = p.dW + (int64_t)(ky * p.kw + kx) * p.o_st + (int64_t)gc * p.o_sg + c0;
#pragma unroll 1
            for (int c = 0; c < 64; c += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(slot * 64 + c), v);
                tmem_ld_wait();
                if (gc < p.GC) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c + j), "f"(__uint_as_float(v[j])),
                                     "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                     : "memory");
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

int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

}  // namespace

// VP_EUNSUPPORTED when the problem is not of the tap-row form (caller falls back to tapwgrad_tc_kernel).
// Requires the filter to be kh x kw with taps listed row-major (gather_taps order) and pad = (kw-1)/2 = (kh-1)/2.
int launch_tapwgrad_win(const TapWgrad& p, int kh, int kw, int pad, cudaStream_t s) {
    if (!tc_available() || getenv("VP_WGRAD_OLD")) return VP_EUNSUPPORTED;
    if (kh * kw != p.taps.ntaps || kw < 2 || kw > kMaxKw || kh > kMaxKw || pad != (kw - 1) / 2 || pad != (kh - 1) / 2) return VP_EUNSUPPORTED;
    if (p.GC % 64 != 0 || p.AC % 64 != 0 || p.as < 1 || p.as > kMaxCls || p.gh < 8 || p.gw < 8) return VP_EUNSUPPORTED;
    if (((uintptr_t)p.G & 15) || ((uintptr_t)p.A & 15)) return VP_EUNSUPPORTED;
    const int st = p.as;
    RowParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.dW = p.dWp; rp.o_st = p.o_st; rp.o_sg = p.o_sg; rp.GC = p.GC; rp.AC = p.AC;
    rp.kw = kw; rp.s = st; rp.nrows = kh;
    for (int ky = 0; ky < kh; ++ky) {
        const int ty = ky - pad;
        const int a = floor_div(ty, st);
        rp.row_a[ky] = (int8_t)a;
        rp.row_r[ky] = (int8_t)(ty - a * st);
    }
    // column classes
    int ncls = 0, off = 0;
    int cls_of_c[kMaxCls];
    for (int c = 0; c < st; ++c) {
        int nt_ = 0, bmin = 1 << 20, bmax = -(1 << 20);
        for (int kx = 0; kx < kw; ++kx) {
            const int tx = kx - pad;
            const int b = floor_div(tx, st);
            if (tx - b * st != c) continue;
            rp.cls_kx[ncls][nt_] = (int8_t)kx; rp.cls_acc[ncls][nt_] = (int8_t)kx; ++nt_;
            bmin = b < bmin ? b : bmin; bmax = b > bmax ? b : bmax;
        }
        cls_of_c[c] = -1;
        if (nt_ == 0) continue;
        cls_of_c[c] = ncls;
        rp.cls_ntaps[ncls] = (int8_t)nt_; rp.cls_bmin[ncls] = (int8_t)bmin; rp.cls_wh[ncls] = (int8_t)(8 + bmax - bmin);
        rp.cls_off[ncls] = off;
        off += ((8 + bmax - bmin) * 8 * 128 + 1023) & ~1023;
        ++ncls;
    }
    rp.ncls = ncls;
    rp.halo_stage_bytes = off;
    {
        int slot = 0, ng = 0;
        for (int ci = 0; ci < ncls; ++ci) {
            int left = rp.cls_ntaps[ci], j = 0;
            while (left > 0) {
                const int take = left > 4 ? (left + 1) / 2 > 4 ? 4 : (left + 1) / 2 : left;
                const int kx0 = rp.cls_kx[ci][j];
                const int b0 = floor_div(kx0 - pad, st) - rp.cls_bmin[ci];
                rp.grp_off[ng] = rp.cls_off[ci] + b0 * 128;
                rp.grp_col[ng] = slot * 64;
                rp.grp_wh[ng] = rp.cls_wh[ci];
                rp.grp_idesc[ng] = idesc_mn64(128, 64 * take);
                for (int q = 0; q < take; ++q) rp.acc_kx[slot + q] = rp.cls_kx[ci][j + q];
                slot += take; j += take; left -= take; ++ng;
            }
        }
        rp.ngroups = ng;
    }
    rp.tiles_w = (p.gw + 7) / 8; rp.tiles_h = (p.gh + 7) / 8;
    const int64_t nbricks = (int64_t)rp.tiles_w * rp.tiles_h * p.n;
    if (nbricks > 0x7fffffff) return VP_EUNSUPPORTED;
    rp.nbricks = (int)nbricks;
    rp.mtiles = (p.GC + 127) / 128; rp.ntiles = p.AC / 64;
    const int sets = kh * rp.mtiles * rp.ntiles;
    // pixel splits: minimise  waves x (bricks per CTA + epilogue)  -- one CTA per SM (the kw accumulators fill TMEM), so a
    // launch of `sets x nsplit` CTAs runs in ceil(./#SMs) waves; the wide layers (sets = 80 / 160 at 512 channels) would
    // otherwise leave half of the machine idle in the last wave
    const int max_split = (rp.nbricks + 7) / 8;
    int cand_max = num_sms() / sets > 8 ? num_sms() / sets : 8;
    if (cand_max > max_split) cand_max = max_split;
    if (cand_max < 1) cand_max = 1;
    int nsplit = 1;
    int64_t best = -1;
    for (int c = 1; c <= cand_max; ++c) {
        const int bps = (rp.nbricks + c - 1) / c;
        const int ctas = sets * ((rp.nbricks + bps - 1) / bps);
        const int64_t waves = (ctas + num_sms() - 1) / num_sms();
        const int64_t cost = waves * (bps + 8);          // an epilogue (kw x 128 x 64 fp32 reductions) costs about 8 bricks
        if (best < 0 || cost < best) { best = cost; nsplit = c; }
    }
    rp.bricks_per_split = (rp.nbricks + nsplit - 1) / nsplit;
    nsplit = (rp.nbricks + rp.bricks_per_split - 1) / rp.bricks_per_split;

    EncodeTiledFn encode = get_encode();
    CUtensorMap mG;
    {
        cuuint64_t dims[4] = {(cuuint64_t)p.GC, (cuuint64_t)p.gw, (cuuint64_t)p.gh, (cuuint64_t)p.n};
        cuuint64_t strides[3] = {(cuuint64_t)p.GC * 2, (cuuint64_t)p.gw * p.GC * 2, (cuuint64_t)p.gh * p.gw * p.GC * 2};
        cuuint32_t box[4] = {64, 8, 8, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (encode(&mG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.G), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return VP_EUNSUPPORTED;
    }
    // sub-lattice (r, c) of A: pixels (s*i + r, s*j + c); one map set per row parity r
    RowMaps maps[kMaxCls];
    memset(maps, 0, sizeof(maps));
    for (int r = 0; r < st; ++r) {
        bool used = false;
        for (int ky = 0; ky < kh; ++ky) used |= rp.row_r[ky] == r;
        for (int c = 0; c < st; ++c) {
            const int ci = cls_of_c[c];
            if (ci < 0) continue;
            const int hs = (p.ha - r + st - 1) / st, ws = (p.wa - c + st - 1) / st;
            if (!used || hs <= 0 || ws <= 0) { maps[r].m[ci] = mG; continue; }     // never dereferenced with a non-empty box inside the tensor
            const uint8_t* base = (const uint8_t*)p.A + ((int64_t)r * p.wa + c) * p.AC * 2;
            cuuint64_t dims[4] = {(cuuint64_t)p.AC, (cuuint64_t)ws, (cuuint64_t)hs, (cuuint64_t)p.n};
            cuuint64_t strides[3] = {(cuuint64_t)st * p.AC * 2, (cuuint64_t)st * p.wa * p.AC * 2, (cuuint64_t)p.ha * p.wa * p.AC * 2};
            cuuint32_t box[4] = {64, (cuuint32_t)rp.cls_wh[ci], 8, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            if (encode(&maps[r].m[ci], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<uint8_t*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return VP_EUNSUPPORTED;
        }
    }
    const int smem_bytes = smem_for_occupancy(kGStages * kGBytes + kHStages * rp.halo_stage_bytes + (2 * kGStages + 2 * kHStages + 1) * 8 + 16 + 1024, 1);
    if (smem_bytes > 227 * 1024) return VP_EUNSUPPORTED;
    static int attr_set = 0;
    if (attr_set < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(tapwgrad_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) { set_error("tapwgrad_row: cannot set %d bytes of dynamic smem: %s", smem_bytes, cudaGetErrorString(e)); return VP_ECUDA; }
        attr_set = smem_bytes;
    }
    if (!p.accumulate) zero_async(p.dWp, sizeof(float) * (size_t)p.taps.ntaps * p.GC * p.AC, s);
    dim3 grid((unsigned)sets, (unsigned)nsplit);
    launch_k(tapwgrad_row_kernel, grid, dim3(kRThreads), smem_bytes, s, mG, maps[0], maps[1], maps[2], maps[3], rp);
    VP_CHECK_LAUNCH("tapwgrad_row");
    return VP_OK;
}

}  // namespace vp
