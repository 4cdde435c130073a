// tcgen05 engine of the tap GEMM (sm_100a): TMA-fed implicit GEMM, bf16 operands, fp32 accumulation in TMEM.
//
//   D[n, gy*ds+doy, gx*ds+dox, :] = sum_t A[n, gy*as+ty_t, gx*as+tx_t, :] . Wp[widx_t][:, :]
//
// One CTA computes a 128 x BLOCK_N output tile; the 128 rows are a (bt x ht x wt) brick of the iteration grid,
// so that for every filter tap the A operand of the tile is ONE 4-D TMA box of the channels-last activation
// tensor (box {64 ch, wt, ht, bt}, element strides {1, as, as, 1}; out-of-bounds rows -- the conv zero padding
// -- are zero-filled by the TMA unit).  The box lands in shared memory as 128 rows of 128 B with the 128-byte
// swizzle, which is exactly the K-major SW128 canonical layout tcgen05.mma consumes.  B is a {64, BLOCK_N, 1}
// box of the packed weights viewed as [K, N, taps].
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = barrier init + TMEM alloc + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias/activation -> global, one output row per thread).
// Pipelines: smem full/empty mbarriers between TMA and MMA, one "accumulator ready" mbarrier MMA -> epilogue.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace vp {

// ---- host-side helpers shared by the tcgen05 launchers (declared in tc_common.cuh) ----

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

// Tensor map of the weight operand of a tap GEMM: element (tap, n, k) at Wp[tap*w_st + n*w_sn + k*w_sk] (bf16).
// K-major (w_sk == 1): dims {K, N, taps}, box {64, BN, 1}.  MN-major (w_sn == 1): dims {N, K, taps}, box {64, 64, 1}.
int encode_weight_map(CUtensorMap* m, const TapGemm& p, bool bmn, int BN) {
    EncodeTiledFn encode = get_encode();
    const int64_t s1 = bmn ? p.w_sk : p.w_sn;
    cuuint64_t dims[3] = {(cuuint64_t)(bmn ? p.N : p.K), (cuuint64_t)(bmn ? p.K : p.N), (cuuint64_t)kMaxTaps};
    cuuint64_t strides[2] = {(cuuint64_t)s1 * 2, (cuuint64_t)p.w_st * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)(bmn ? 64 : BN), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if ((strides[0] & 15) != 0 || (strides[1] & 15) != 0) return -1;
    // the tap extent is only an upper bound for the descriptor; taps actually addressed are < the real count
    CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p.Wp), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

int encode_out_map(CUtensorMap* m, void* D, int N, int hd, int wd, int n, int ds, int doy, int dox, int bw, int bh, int bb) {
    EncodeTiledFn encode = get_encode();
    const int gh = (hd - doy + ds - 1) / ds, gw = (wd - dox + ds - 1) / ds;
    if (gh <= 0 || gw <= 0) return -1;
    uint8_t* base = (uint8_t*)D + ((int64_t)doy * wd + dox) * N * 2;
    cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)gw, (cuuint64_t)gh, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)ds * N * 2, (cuuint64_t)ds * wd * N * 2, (cuuint64_t)hd * wd * N * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (((uintptr_t)base & 15) != 0) return -1;
    CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
int pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }

// SMs the persistent kernels size their grids for: the device's count, or the caller's cap (vp_set_sm_limit) while a
// collective runs next to them -- a persistent grid must be fully resident, so the SMs the collective's CTAs hold are left out.
static int g_sm_limit = 0;
void set_sm_limit(int n) { g_sm_limit = n > 0 ? n : 0; }
int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}


namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // bf16 elements = 128 bytes = one swizzle row
constexpr int kFwdThreads = 224;       // fwd/dgrad kernel: + a second TMA producer warp (B operand)
constexpr int kABytes = kBlockM * kBlockK * 2;

using namespace tc;

constexpr int kMaxPhases = 4;
constexpr int kBoxBytes = 64 * 128;  // {64 channels, 64 rows} bf16 box of an MN-major operand

// MN-major SWIZZLE_128B operand: 8-row (K) atoms of 1024 B, LBO = next 64-wide MN group (one box), 16 K rows per MMA = 2048 B
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(kBoxBytes >> 4) << 16;   // LBO: next 64-wide MN group
    d |= (uint64_t)(1024 >> 4) << 32;        // SBO: next 8 K rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct TcPhase {
    int gh, gw;            // iteration grid of this phase
    int doy, dox;          // output offset (scatter form)
    int tiles_w, tiles_h;  // bricks along gx, gy
    int tile_begin;        // first M-tile index of this phase in the launch-wide tile list
    TapList taps;
};

struct TcParams {
    void* D;
    const float* bias;
    int n, hd, wd, N;
    int as, ds;
    int act;
    float slope;
    int out_f32;
    int kblocks;           // K / 64
    int bt, ht, wt;        // tile brick, bt*ht*wt == 128
    int ntiles_n;          // N tiles
    int total_tiles;       // sum over phases of M-tiles, times ntiles_n (times ksplit)
    int nphases;
    int ksplit;            // > 1: the (tap, k-block) iterations of a tile are split over ksplit CTAs, partial tiles are
    int iters_per_split;   //      added into an fp32 workspace (D) with red.global.add and finished by splitk_finish_kernel
    int base_tiles;
    int tma_store;         // bf16 output, N % 64 == 0: epilogue through per-warp staging tiles + TMA stores (omaps)
    float* stat_parts;     // [gridDim.x][2][N] per-CTA BatchNorm partial sums of the stored output, or null
    TcPhase ph[kMaxPhases];
};

template <int BN, int STAGES>
struct SmemLayout {
    static constexpr int kBBytes = BN * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kOutOffset = STAGES * kStageBytes;                 // 4 warps x [32 rows][64 B] staging tiles
    static constexpr int kStatOffset = kOutOffset + 4 * 2048;               // sum[kStatMaxN], sumsq[kStatMaxN]
    static constexpr int kBarOffset = kStatOffset + 2 * kStatMaxN * 4;
    static constexpr int kTotal = kBarOffset + (2 * STAGES + 4) * 8 + 16;
};

struct TileCoord { int phase, n0, gy0, gx0, col0, split; };

__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int q, int BN) {
    TileCoord c;
    c.split = 0;
    if (p.ksplit > 1) { c.split = q / p.base_tiles; q -= c.split * p.base_tiles; }
    const int nt = q % p.ntiles_n;
    int mt = q / p.ntiles_n;
    int ph = 0;
#pragma unroll
    for (int i = 1; i < kMaxPhases; ++i)
        if (i < p.nphases && mt >= p.ph[i].tile_begin) ph = i;
    mt -= p.ph[ph].tile_begin;
    const int tw = mt % p.ph[ph].tiles_w; mt /= p.ph[ph].tiles_w;
    const int th = mt % p.ph[ph].tiles_h; mt /= p.ph[ph].tiles_h;
    c.phase = ph; c.n0 = mt * p.bt; c.gy0 = th * p.ht; c.gx0 = tw * p.wt; c.col0 = nt * BN;
    return c;
}

// Persistent, warp-specialised tap GEMM: grid = min(#tiles, #SMs); every CTA walks tiles q = blockIdx.x + i*gridDim.x
// (all output-parity phases of a transposed conv / strided dgrad are in ONE launch).  The smem ring runs across
// tile boundaries and the accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i
// overlaps the TMA/MMA main loop of tile i+1.
template <int BN, int STAGES, int CTAS_PER_SM, bool BMN>
__global__ void __launch_bounds__(kFwdThreads, CTAS_PER_SM) tapgemm_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                 const __grid_constant__ CUtensorMap mapB,
                                                                 const __grid_constant__ OutMaps omaps,
                                                                 const __grid_constant__ TcParams p) {
    using L = SmemLayout<BN, STAGES>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SW128 tiles: 1024-byte aligned
    uint64_t* full = (uint64_t*)(smem + L::kBarOffset);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;     // [2]
    uint64_t* acc_empty = acc_full + 2;      // [2]
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
    float* s_stat = (float*)(smem + L::kStatOffset);
    if (p.stat_parts)
        for (int i = threadIdx.x; i < 2 * kStatMaxN; i += kFwdThreads) s_stat[i] = 0.f;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int kAccCols = BN < 32 ? 32 : BN;          // columns per accumulator buffer
    constexpr int kTmemCols = 2 * kAccCols;              // power of two >= 64

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }   // full: A and B producers
            mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
            mbar_init(&acc_empty[0], 4); mbar_init(&acc_empty[1], 4);      // one arrive per epilogue warp
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();     // everything above overlapped the previous kernel's tail; global memory is touched only from here on

    if (warp == 0 || warp == 6) {
        // ===== TMA producers: warp 0 streams the A bricks, warp 6 the weight panels (two issuing threads) =====
        const bool is_a = warp == 0;
        if (elect_one()) {
            uint32_t g = 0;   // global k-iteration counter (ring position)
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x) {
                const TileCoord tc = decode_tile(p, q, BN);
                const TcPhase& ph = p.ph[tc.phase];
                int iters = ph.taps.ntaps * p.kblocks, it0 = 0;
                if (p.ksplit > 1) { it0 = tc.split * p.iters_per_split; iters = min(iters, it0 + p.iters_per_split); }
                for (int it = it0; it < iters; ++it, ++g) {
                    const int s = g % STAGES;
                    mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
                    const int t = it / p.kblocks;
                    const int kb = it - t * p.kblocks;
                    uint8_t* sa = smem + s * L::kStageBytes;
                    if (is_a) {
                        mbar_expect_tx(&full[s], kABytes);
                        tma_load_4d(sa, &mapA, &full[s], kb * kBlockK, tc.gx0 * p.as + ph.taps.tx[t], tc.gy0 * p.as + ph.taps.ty[t], tc.n0);
                    } else {
                        mbar_expect_tx(&full[s], L::kBBytes);
                        if constexpr (BMN) {      // weight read in place, N contiguous: 64 (n) x 64 (k) boxes = MN-major atoms
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j)
                                tma_load_3d(sa + kABytes + j * kBoxBytes, &mapB, &full[s], tc.col0 + 64 * j, kb * kBlockK, ph.taps.widx[t]);
                        } else {
                            tma_load_3d(sa + kABytes, &mapB, &full[s], kb * kBlockK, tc.col0, ph.taps.widx[t]);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = idesc_bf16_f32(kBlockM, BN < 16 ? 16 : BN) | (BMN ? (1u << 16) : 0u);
        if (elect_one()) {
            uint32_t g = 0, i = 0;
            for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++i) {
                const TileCoord tc = decode_tile(p, q, BN);
                int iters = p.ph[tc.phase].taps.ntaps * p.kblocks, it0 = 0;
                if (p.ksplit > 1) { it0 = tc.split * p.iters_per_split; iters = min(iters, it0 + p.iters_per_split); }
                const uint32_t buf = i & 1, use = i >> 1;
                mbar_wait(&acc_empty[buf], (use & 1) ^ 1);          // epilogue has drained this TMEM buffer
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * kAccCols;
                for (int it = it0; it < iters; ++it, ++g) {
                    const int s = g % STAGES;
                    mbar_wait(&full[s], (g / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * L::kStageBytes);
                    const uint64_t adesc = smem_desc_k_sw128(sa);
                    const uint64_t bdesc = BMN ? smem_desc_mn_sw128(sa + kABytes) : smem_desc_k_sw128(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)   // per UMMA_K=16: K-major +32 B (+2 in the addr>>4 field), MN-major +16 rows = 2048 B (+128)
                        tc_mma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * (BMN ? 128 : 2)), idesc, ((it - it0) | k) != 0);
                    tc_commit(&empty[s]);
                }
                tc_commit(&acc_full[buf]);
            }
        }
    } else if (warp >= 2 && warp <= 5) {
        // ===== epilogue: warps 2..5; warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
        const int lane_base = (warp & 3) * 32;
        const int r = lane_base + lane;                       // row of the tile = TMEM lane
        const int bw = r % p.wt;
        const int bh = (r / p.wt) % p.ht;
        const int bb = r / (p.wt * p.ht);
        uint32_t i = 0;
        for (int q = blockIdx.x; q < p.total_tiles; q += gridDim.x, ++i) {
            const TileCoord tc = decode_tile(p, q, BN);
            const TcPhase& ph = p.ph[tc.phase];
            const int n = tc.n0 + bb, gy = tc.gy0 + bh, gx = tc.gx0 + bw;
            const int oy = gy * p.ds + ph.doy, ox = gx * p.ds + ph.dox;
            const bool row_ok = n < p.n && gy < ph.gh && gx < ph.gw && oy < p.hd && ox < p.wd;
            const int64_t row_off = (((int64_t)n * p.hd + oy) * p.wd + ox) * p.N;
            const uint32_t buf = i & 1, use = i >> 1;
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
            constexpr int CH = BN >= 32 ? 32 : 16;
            if (BN >= 32 && p.tma_store) {
                // staging tile -> (BatchNorm partial sums) -> TMA store; the warp's 32 rows are a sub-brick of the tile
                uint8_t* st = smem + L::kOutOffset + (warp & 3) * 2048;
                const int r0 = lane_base;
                const uint32_t row_mask = __ballot_sync(0xffffffffu, row_ok);
                const int sx = tc.gx0 + r0 % p.wt, sy = tc.gy0 + (r0 / p.wt) % p.ht, sn = tc.n0 + r0 / (p.wt * p.ht);
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    if (tc.col0 + c >= p.N) break;
                    uint32_t v[32];
                    tmem_ld32(tmem_base + buf * kAccCols + ((uint32_t)lane_base << 16) + (uint32_t)c, v);
                    if (lane == 0) tma_store_wait_read<0>();       // the previous store has finished reading the staging tile
                    tmem_ld_wait();
                    __syncwarp();
                    stage_chunk32_sw64(st, lane, tc.col0 + c, v, p.bias, p.act, p.slope);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (p.stat_parts) stats_chunk32_sw64(st, lane, s_stat + tc.col0 + c, s_stat + kStatMaxN + tc.col0 + c, row_mask);
                    if (lane == 0) {
                        tma_store_4d(&omaps.m[tc.phase], st, tc.col0 + c, sx, sy, sn);
                        tma_store_commit();
                    }
                }
            } else {
#pragma unroll 1
            for (int c = 0; c < BN; c += CH) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + buf * kAccCols + ((uint32_t)lane_base << 16) + (uint32_t)c;
                if (CH == 32) tmem_ld32(taddr, v);
                else tmem_ld16(taddr, v);
                tmem_ld_wait();
                if (row_ok) {
                    if (p.ksplit > 1) {
                        float* out = (float*)p.D + row_off + tc.col0 + c;
                        if (tc.col0 + c + CH <= p.N && (p.N & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < CH; j += 4)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + j), "f"(__uint_as_float(v[j])),
                                             "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                             : "memory");
                        } else {
#pragma unroll
                            for (int j = 0; j < CH; ++j)
                                if (tc.col0 + c + j < p.N) atomicAdd(out + j, __uint_as_float(v[j]));
                        }
                    } else {
                        store_chunk<CH>(v, p.D, row_off, tc.col0 + c, p.N, p.bias, p.act, p.slope, p.out_f32 != 0);
                    }
                }
            }
            }
            // this warp's TMEM reads are complete (wait::ld above): hand the buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[buf])) : "memory");
        }
        if (p.tma_store && lane == 0) tma_store_wait_read<0>();
    }
    // teardown: everyone is done with TMEM before the allocating warp frees it
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
    if (p.stat_parts) {
        float* out = p.stat_parts + (size_t)blockIdx.x * 2 * p.N;
        for (int i = threadIdx.x; i < 2 * p.N; i += kFwdThreads) out[i] = s_stat[(i < p.N) ? i : (kStatMaxN + i - p.N)];
    }
}

// ---- host side ------------------------------------------------------------------------------------------
template <int BN, int STAGES, int CTAS, bool BMN = false>
int launch_cfg(const CUtensorMap& mA, const CUtensorMap& mB, const OutMaps& om, const TcParams& tp, cudaStream_t s, int* grid_out = nullptr) {
    using L = SmemLayout<BN, STAGES>;
    const int smem_bytes = smem_for_occupancy(L::kTotal + 1024, CTAS);  // + alignment slack; never more than CTAS CTAs per SM
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tapgemm_tc_kernel<BN, STAGES, CTAS, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) { set_error("tapgemm_tc: cannot set %d bytes of dynamic smem: %s", smem_bytes, cudaGetErrorString(e)); return VP_ECUDA; }
        attr_set = true;
    }
    const int slots = num_sms() * CTAS;
    const int grid = tp.total_tiles < slots ? tp.total_tiles : slots;
    if (grid_out) *grid_out = grid;
    launch_k(tapgemm_tc_kernel<BN, STAGES, CTAS, BMN>, dim3(grid), dim3(kFwdThreads), smem_bytes, s, mA, mB, om, tp);
    VP_CHECK_LAUNCH("tapgemm_tc");
    return VP_OK;
}

// experiment switch (tools/run_layer.py): VP_TC_VARIANT=1 -> one CTA per SM with deep rings and BN=256 tiles
int tc_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VP_TC_VARIANT"); v = e ? atoi(e) : 0; }
    return v;
}

}  // namespace

// ---- split-K support ------------------------------------------------------------------------------------------------
// one registered buffer per device (a process may drive several GPUs); selected by the current device of the calling thread
constexpr int kMaxDevices = 64;
static void* g_ws[kMaxDevices] = {};
static size_t g_ws_bytes[kMaxDevices] = {};
static int current_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
    return dev;
}
void set_splitk_workspace(void* ptr, size_t bytes) {
    const int d = current_device_slot();
    if (d >= 0) { g_ws[d] = ptr; g_ws_bytes[d] = bytes; }
}
void* splitk_workspace(size_t bytes) {
    const int d = current_device_slot();
    return (d >= 0 && bytes <= g_ws_bytes[d]) ? g_ws[d] : nullptr;
}

namespace {
__global__ void __launch_bounds__(256) splitk_finish_kernel(const float* __restrict__ ws, void* __restrict__ D, const float* __restrict__ bias, int act,
                                                            float slope, int64_t n, int N, int out_f32) {
    pdl_sync();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float x = ws[i];
        if (bias) x += bias[i % N];
        x = act_fwd(x, act, slope);
        if (out_f32) ((float*)D)[i] = x;
        else ((bf16*)D)[i] = __float2bfloat16_rn(x);
    }
}
}  // namespace
int launch_splitk_finish(const float* ws, void* D, const float* bias, int act, float slope, int64_t n, int N, bool out_f32, cudaStream_t s) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(splitk_finish_kernel, dim3((unsigned)blocks), dim3(256), 0, s, ws, D, bias, act, slope, n, N, out_f32 ? 1 : 0);
    VP_CHECK_LAUNCH("splitk_finish");
    return VP_OK;
}

bool tc_available() {
    static int cached = -1;
    if (cached < 0) {
        int dev = 0, maj = 0;
        cached = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess &&
            maj == 10 && get_encode() != nullptr)
            cached = 1;
    }
    return cached == 1;
}

// All phases share A, Wp, D, K, N, as, ds, bias, act; they differ in (gh, gw, doy, dox, taps).
int launch_tapgemm_tc_multi(const TapGemm* phases, int nphases, cudaStream_t s) {
    const TapGemm& p = phases[0];
    if (p.bn_y) {
        // fused BatchNorm-backward reduction: only the kernels that carry that epilogue may take the problem
        if (nphases == 1 && tc_variant() != 2) return launch_tapgemm_gwin(phases[0], s);
        return VP_EUNSUPPORTED;
    }
    // stride-1 tap sets go to the windowed (halo re-use) kernel first; VP_TC_VARIANT=2 disables it (A/B measurements)
    if (tc_variant() != 2 && nphases == 1) {
        const int rcg = launch_tapgemm_gwin(phases[0], s);
        if (rcg != VP_EUNSUPPORTED) return rcg;
    }
    if (tc_variant() != 2) {
        const int rc0 = launch_tapgemm_pair(phases, nphases, s);
        if (rc0 != VP_EUNSUPPORTED) return rc0;
        const int rc = launch_tapgemm_win(phases, nphases, s);
        if (rc != VP_EUNSUPPORTED) return rc;
    }
    // ---- eligibility -----------------------------------------------------------------------------------
    if (!tc_available()) { set_error("tcgen05 engine: needs an sm_100 device and cuTensorMapEncodeTiled"); return VP_EUNSUPPORTED; }
    if (nphases < 1 || nphases > kMaxPhases) { set_error("tcgen05 engine: %d phases", nphases); return VP_EUNSUPPORTED; }
    if (p.K % kBlockK != 0 || p.N < 1) { set_error("tcgen05 engine: K=%d must be a multiple of 64", p.K); return VP_EUNSUPPORTED; }
    for (int i = 0; i < nphases; ++i)
        if (phases[i].taps.ntaps < 1) { set_error("tcgen05 engine: phase without taps"); return VP_EUNSUPPORTED; }
    if (((uintptr_t)p.A & 15) || ((uintptr_t)p.Wp & 15) || ((uintptr_t)p.D & 15)) { set_error("tcgen05 engine: 16-byte alignment"); return VP_EUNSUPPORTED; }
    if (p.as < 1 || p.as > 8) { set_error("tcgen05 engine: gather stride %d", p.as); return VP_EUNSUPPORTED; }
    if (p.n <= 0) return VP_OK;
    EncodeTiledFn encode = get_encode();

    // ---- tile brick (from the largest phase grid) ----------------------------------------------------------
    int gh = 0, gw = 0;
    for (int i = 0; i < nphases; ++i) { gh = phases[i].gh > gh ? phases[i].gh : gh; gw = phases[i].gw > gw ? phases[i].gw : gw; }
    if (gh <= 0 || gw <= 0) return VP_OK;
    TcParams tp;
    memset(&tp, 0, sizeof(tp));
    int wt = pow2_floor(gw < kBlockM ? gw : kBlockM);
    if (wt * p.as > 256) wt = pow2_floor(256 / p.as);
    int ht = pow2_ceil(gh);
    if (ht > kBlockM / wt) ht = kBlockM / wt;
    if (ht * p.as > 256) ht = pow2_floor(256 / p.as);
    const int bt = kBlockM / (wt * ht);
    tp.bt = bt; tp.ht = ht; tp.wt = wt;
    const int tiles_b = (p.n + bt - 1) / bt;
    // 256-wide tiles (one CTA per SM, 4 stages): an M128 x N256 MMA reads 12 KB of operands per 128 issue cycles instead of
    // 8 KB per 64 -- measured +10..17 % on the N = 256 layers; everything else keeps two 128-wide CTAs per SM
    int64_t mt256 = 0;
    for (int i = 0; i < nphases; ++i)
        mt256 += (int64_t)((phases[i].gw + wt - 1) / wt) * ((phases[i].gh + ht - 1) / ht) * tiles_b;
    const bool wide = p.N % 256 == 0 && mt256 * (p.N / 256) * 3 >= num_sms() * 2 && tc_variant() != 5;
    const bool one_cta = tc_variant() == 1 || wide;
    const int BN = (one_cta && p.N % 256 == 0) ? 256 : (p.N % 128 == 0) ? 128 : (p.N >= 64 ? 64 : (p.N > 16 ? 32 : 16));
    tp.ntiles_n = (p.N + BN - 1) / BN;
    int64_t mtiles = 0;
    for (int i = 0; i < nphases; ++i) {
        TcPhase& ph = tp.ph[i];
        ph.gh = phases[i].gh; ph.gw = phases[i].gw; ph.doy = phases[i].doy; ph.dox = phases[i].dox; ph.taps = phases[i].taps;
        ph.tiles_w = (ph.gw + wt - 1) / wt;
        ph.tiles_h = (ph.gh + ht - 1) / ht;
        ph.tile_begin = (int)mtiles;
        mtiles += (int64_t)ph.tiles_w * ph.tiles_h * tiles_b;
    }
    if (mtiles * tp.ntiles_n > 0x7fffffff) { set_error("tcgen05 engine: too many tiles"); return VP_EUNSUPPORTED; }
    tp.total_tiles = (int)(mtiles * tp.ntiles_n);
    tp.nphases = nphases;
    tp.ksplit = 1; tp.base_tiles = tp.total_tiles;
    // skinny problems (the fc layers: a handful of output tiles, thousands of K iterations): split K over the idle SMs
    float* ws = nullptr;
    {
        const int iters = phases[0].taps.ntaps * (p.K / kBlockK);
        const int64_t out_elems = (int64_t)p.n * p.hd * p.wd * p.N;
        if (nphases == 1 && tp.total_tiles * 2 <= num_sms() && iters >= 16 && !p.stat_parts && tc_variant() != 3) {
            int ks = num_sms() / tp.total_tiles;
            if (ks > iters / 4) ks = iters / 4;
            ws = (float*)splitk_workspace((size_t)out_elems * sizeof(float));
            if (ks > 1 && ws) {
                tp.iters_per_split = (iters + ks - 1) / ks;
                tp.ksplit = (iters + tp.iters_per_split - 1) / tp.iters_per_split;
                tp.total_tiles = tp.base_tiles * tp.ksplit;
                zero_async(ws, (size_t)out_elems * sizeof(float), s);
            }
        }
    }

    // ---- tensor maps -----------------------------------------------------------------------------------
    CUtensorMap mA, mB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)p.K, (cuuint64_t)p.wa, (cuuint64_t)p.ha, (cuuint64_t)p.n};
        cuuint64_t strides[3] = {(cuuint64_t)p.K * 2, (cuuint64_t)p.wa * p.K * 2, (cuuint64_t)p.ha * p.wa * p.K * 2};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)(wt * p.as), (cuuint32_t)(ht * p.as), (cuuint32_t)bt};
        cuuint32_t estr[4] = {1, (cuuint32_t)p.as, (cuuint32_t)p.as, 1};
        CUresult r = encode(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.A), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("tcgen05 engine: cuTensorMapEncodeTiled(A) failed (%d)", (int)r); return VP_EUNSUPPORTED; }
    }
    const bool bmn = p.w_sn == 1 && p.w_sk != 1;
    if (!bmn && p.w_sk != 1) { set_error("tcgen05 engine: weight operand must be contiguous along K or along N"); return VP_EUNSUPPORTED; }
    if (bmn && (p.N % 64 != 0 || BN < 64)) { set_error("tcgen05 engine: MN-major weights need N %% 64 == 0"); return VP_EUNSUPPORTED; }
    if (encode_weight_map(&mB, p, bmn, BN)) { set_error("tcgen05 engine: cuTensorMapEncodeTiled(B) failed"); return VP_EUNSUPPORTED; }
    tp.D = p.D; tp.bias = p.bias; tp.n = p.n; tp.hd = p.hd; tp.wd = p.wd; tp.N = p.N;
    tp.as = p.as; tp.ds = p.ds; tp.act = p.act; tp.slope = p.slope;
    tp.out_f32 = (p.out_dtype == VP_F32); tp.kblocks = p.K / kBlockK;
    OutMaps om;
    memset(&om, 0, sizeof(om));
    int grid_used = 0;
    tp.tma_store = 0; tp.stat_parts = nullptr;
    if (tp.ksplit == 1 && !tp.out_f32 && p.N % 64 == 0 && BN >= 32 && tc_variant() != 4) {
        // warp q of the epilogue owns rows [32q, 32q+32) of the (bt x ht x wt) brick: a {bw, bh, bb} sub-brick
        const int bw = wt < 32 ? wt : 32, bh = ht < 32 / bw ? ht : 32 / bw, bb = 32 / (bw * bh);
        bool ok = true;
        for (int i = 0; i < nphases && ok; ++i)
            ok = encode_out_map(&om.m[i], p.D, p.N, p.hd, p.wd, p.n, p.ds, phases[i].doy, phases[i].dox, bw, bh, bb) == 0;
        if (ok) {
            tp.tma_store = 1;
            if (p.stat_parts && p.N <= kStatMaxN && p.bias == nullptr && p.act == VP_ACT_NONE) tp.stat_parts = p.stat_parts;
        }
    }
    if (p.stat_parts && !tp.stat_parts) { set_error("tcgen05 engine: epilogue statistics not available for this shape"); return VP_EUNSUPPORTED; }
    if (tp.stat_parts) {
        const int slots = num_sms() * (one_cta ? 1 : 2);
        const int g = tp.total_tiles < slots ? tp.total_tiles : slots;
        if (g > p.stat_capacity) { set_error("tcgen05 engine: statistics buffer holds %d parts, %d needed", p.stat_capacity, g); return VP_EINVAL; }
        if (p.stat_nparts) *p.stat_nparts = g;
    }
    if (tp.ksplit > 1) {
        tp.D = ws; tp.out_f32 = 1;
        int rc;
        if (bmn) rc = BN == 128 ? launch_cfg<128, 3, 2, true>(mA, mB, om, tp, s, &grid_used) : launch_cfg<64, 4, 2, true>(mA, mB, om, tp, s, &grid_used);
        else switch (BN) {
            case 128: rc = launch_cfg<128, 3, 2>(mA, mB, om, tp, s, &grid_used); break;
            case 64: rc = launch_cfg<64, 4, 2>(mA, mB, om, tp, s, &grid_used); break;
            case 32: rc = launch_cfg<32, 5, 2>(mA, mB, om, tp, s, &grid_used); break;
            default: rc = launch_cfg<16, 5, 2>(mA, mB, om, tp, s, &grid_used); break;
        }
        if (rc) return rc;
        return launch_splitk_finish(ws, p.D, p.bias, p.act, p.slope, (int64_t)p.n * p.hd * p.wd * p.N, p.N, p.out_dtype == VP_F32, s);
    }
    if (one_cta && bmn) {
        if (BN == 256) return launch_cfg<256, 4, 1, true>(mA, mB, om, tp, s, &grid_used);
        return BN == 128 ? launch_cfg<128, 6, 1, true>(mA, mB, om, tp, s, &grid_used) : launch_cfg<64, 8, 1, true>(mA, mB, om, tp, s, &grid_used);
    }
    if (one_cta) {
        switch (BN) {
            case 256: return launch_cfg<256, 4, 1>(mA, mB, om, tp, s, &grid_used);
            case 128: return launch_cfg<128, 6, 1>(mA, mB, om, tp, s, &grid_used);
            case 64: return launch_cfg<64, 8, 1>(mA, mB, om, tp, s, &grid_used);
            case 32: return launch_cfg<32, 8, 1>(mA, mB, om, tp, s, &grid_used);
            default: return launch_cfg<16, 8, 1>(mA, mB, om, tp, s, &grid_used);
        }
    }
    if (bmn) return BN == 128 ? launch_cfg<128, 3, 2, true>(mA, mB, om, tp, s, &grid_used) : launch_cfg<64, 4, 2, true>(mA, mB, om, tp, s, &grid_used);
    // default: two persistent CTAs per SM (two TMA issue streams, two epilogues in flight), <= 113 KB smem and
    // <= 256 TMEM columns each
    switch (BN) {
        case 128: return launch_cfg<128, 3, 2>(mA, mB, om, tp, s, &grid_used);
        case 64: return launch_cfg<64, 4, 2>(mA, mB, om, tp, s, &grid_used);
        case 32: return launch_cfg<32, 5, 2>(mA, mB, om, tp, s, &grid_used);
        default: return launch_cfg<16, 5, 2>(mA, mB, om, tp, s, &grid_used);
    }
}

int launch_tapgemm_tc(const TapGemm& p, cudaStream_t s) { return launch_tapgemm_tc_multi(&p, 1, s); }

// =====================================================================================================
// wgrad:  dWp[widx_t][gc][ac] += sum_{pixels} G[pix][gc] * A[pix (+) tap_t][ac]
//
// GEMM view per tap: D[M=gc, N=ac] = G^T[gc, pix] . A_t[pix, ac], the reduction runs over pixels.  Both operands
// are read straight from the channels-last tensors, i.e. with the GEMM's M/N index contiguous ("MN-major"):
// a TMA box of {64 channels, 64 pixels} is 64 rows (K) of 128 B (64 MN elements) with the 128-byte swizzle =
// the MN-major SW128 canonical layout (K-row stride 128 B inside an 8-row atom, SBO = 1024 B between atoms,
// LBO = one whole box = 8192 B between 64-channel groups).  One CTA owns (tap, 128-gc tile, BN-ac tile,
// pixel chunk); pixel chunks (split-K) are combined with fp32 red.global.add into the zeroed dWp.
// =====================================================================================================
namespace {

__host__ __device__ constexpr uint32_t idesc_bf16_f32_mn(int m, int n) {
    return idesc_bf16_f32(m, n) | (1u << 15) | (1u << 16);   // A and B MN-major
}

struct TcWgradParams {
    float* dWp;
    int64_t o_st, o_sg;        // element (tap, gc, ac) at dWp[tap*o_st + gc*o_sg + ac]
    int store_only;            // one pixel split and nothing to add to: plain vector stores instead of red.global.add
    int GC, AC;
    int as;
    int bt, ht, wt;            // pixel brick, bt*ht*wt == 64
    int tiles_w, tiles_h;
    int nbricks, bricks_per_split;
    int mtiles, ntiles;
    TapList taps;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kFwdThreads) tapwgrad_tc_kernel(const __grid_constant__ CUtensorMap mapG,
                                                               const __grid_constant__ CUtensorMap mapA,
                                                               const TcWgradParams p) {
    constexpr int kGBytes = 2 * kBoxBytes;            // 128 gc
    constexpr int kAStage = (BN / 64) * kBoxBytes;
    constexpr int kStage = kGBytes + kAStage;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + STAGES * kStage);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_ready = empty + STAGES;
    uint32_t* tmem_slot = (uint32_t*)(acc_ready + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    int id = blockIdx.x;
    const int nt = id % p.ntiles; id /= p.ntiles;
    const int mt = id % p.mtiles; id /= p.mtiles;
    const int tap = id;
    const int b0 = blockIdx.y * p.bricks_per_split;
    const int b1 = min(b0 + p.bricks_per_split, p.nbricks);
    const int iters = b1 - b0;
    if (iters <= 0) return;
    const int m0 = mt * 128, c0 = nt * BN;
    const int ty = p.taps.ty[tap], tx = p.taps.tx[tap];

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapG) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }   // full: G and A producers
            mbar_init(acc_ready, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();     // everything above overlapped the previous kernel's tail; global memory is touched only from here on

    if (warp == 0 || warp == 6) {
        // two TMA issuing threads: warp 0 streams the G (grid-side) boxes, warp 6 the shifted A boxes
        const bool is_g = warp == 0;
        if (elect_one()) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                int b = b0 + it;
                const int tw = b % p.tiles_w; b /= p.tiles_w;
                const int th = b % p.tiles_h; b /= p.tiles_h;
                const int n0 = b * p.bt, gy0 = th * p.ht, gx0 = tw * p.wt;
                uint8_t* sg = smem + s * kStage;
                uint8_t* sa = sg + kGBytes;
                if (is_g) {
                    mbar_expect_tx(&full[s], kGBytes);
                    tma_load_4d(sg, &mapG, &full[s], m0, gx0, gy0, n0);
                    tma_load_4d(sg + kBoxBytes, &mapG, &full[s], m0 + 64, gx0, gy0, n0);
                } else {
                    mbar_expect_tx(&full[s], kAStage);
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                        tma_load_4d(sa + j * kBoxBytes, &mapA, &full[s], c0 + 64 * j, gx0 * p.as + tx, gy0 * p.as + ty, n0);
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = idesc_bf16_f32_mn(128, BN);
        if (elect_one()) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sg = smem_u32(smem + s * kStage);
                const uint64_t adesc = smem_desc_mn_sw128(sg);
                const uint64_t bdesc = smem_desc_mn_sw128(sg + kGBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // 16 pixels (K rows) = two 8-row atoms = 2048 bytes: +128 in the (addr >> 4) field
                    tc_mma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (it | k) != 0);
                }
                tc_commit(&empty[s]);
            }
            tc_commit(acc_ready);
        }
    } else if (warp >= 2 && warp <= 5) {
        const int lane_base = (warp & 3) * 32;
        mbar_wait(acc_ready, 0);
        tc_fence_after();
        // acc_ready also says that every pipeline stage has been consumed (this CTA computes ONE output tile): the ring's
        // shared memory now serves as a per-warp transposition buffer, so that each store / reduction instruction of a warp
        // covers four full 128-byte row segments instead of 32 sixteen-byte pieces of 32 different rows
        constexpr int kTStride = 36;                                // floats per staged row: conflict-free 16-byte accesses
        float* stg = reinterpret_cast<float*>(smem) + (warp & 3) * (32 * kTStride);
        const int cr = lane >> 3, cc = (lane & 7) * 4;
        float* out = p.dWp + (int64_t)p.taps.widx[tap] * p.o_st + (int64_t)(m0 + lane_base + cr) * p.o_sg + c0 + cc;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)c, v);
            tmem_ld_wait();
            __syncwarp();                                           // the previous chunk has been read out of the buffer
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(stg + lane * kTStride + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            __syncwarp();
            if (c0 + c + 32 <= p.AC) {                              // always (AC % 64 == 0, BN | AC): kept as a guard
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = 4 * i + cr;
                    if (m0 + lane_base + row >= p.GC) continue;
                    const float4 val = *reinterpret_cast<const float4*>(stg + row * kTStride + cc);
                    float* o = out + (int64_t)(4 * i) * p.o_sg + c;
                    if (p.store_only) *reinterpret_cast<float4*>(o) = val;
                    else
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w)
                                     : "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
    }
}

template <int BN, int STAGES>
int launch_wgrad_cfg(const CUtensorMap& mG, const CUtensorMap& mA, const TcWgradParams& tp, dim3 grid, cudaStream_t s) {
    constexpr int smem_bytes = STAGES * (2 * kBoxBytes + (BN / 64) * kBoxBytes) + (2 * STAGES + 1) * 8 + 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tapwgrad_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) { set_error("tapwgrad_tc: cannot set %d bytes of dynamic smem: %s", smem_bytes, cudaGetErrorString(e)); return VP_ECUDA; }
        attr_set = true;
    }
    launch_k(tapwgrad_tc_kernel<BN, STAGES>, grid, dim3(kFwdThreads), smem_bytes, s, mG, mA, tp);
    VP_CHECK_LAUNCH("tapwgrad_tc");
    return VP_OK;
}

int encode_nhwc(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, int bw, int bh, int bb, int estride) {
    EncodeTiledFn encode = get_encode();
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(bw * estride), (cuuint32_t)(bh * estride), (cuuint32_t)bb};
    cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace

int launch_tapwgrad_tc(const TapWgrad& p, cudaStream_t s) {
    if (!tc_available()) { set_error("tcgen05 engine: needs an sm_100 device and cuTensorMapEncodeTiled"); return VP_EUNSUPPORTED; }
    if (p.GC % 64 != 0 || p.AC % 64 != 0) { set_error("tcgen05 wgrad: channel counts %d/%d must be multiples of 64", p.GC, p.AC); return VP_EUNSUPPORTED; }
    if (((uintptr_t)p.G & 15) || ((uintptr_t)p.A & 15) || p.as < 1 || p.as > 8) { set_error("tcgen05 wgrad: alignment/stride"); return VP_EUNSUPPORTED; }
    const int64_t P = (int64_t)p.n * p.gh * p.gw;
    if (P <= 0) return VP_OK;
    TcWgradParams tp;
    int wt = pow2_floor(p.gw < 64 ? p.gw : 64);
    if (wt * p.as > 256) wt = pow2_floor(256 / p.as);
    int ht = pow2_ceil(p.gh);
    if (ht > 64 / wt) ht = 64 / wt;
    if (ht * p.as > 256) ht = pow2_floor(256 / p.as);
    const int bt = 64 / (wt * ht);
    tp.bt = bt; tp.ht = ht; tp.wt = wt;
    tp.tiles_w = (p.gw + wt - 1) / wt;
    tp.tiles_h = (p.gh + ht - 1) / ht;
    const int64_t nbricks = (int64_t)tp.tiles_w * tp.tiles_h * ((p.n + bt - 1) / bt);
    if (nbricks > 0x7fffffff) { set_error("tcgen05 wgrad: too many bricks"); return VP_EUNSUPPORTED; }
    tp.nbricks = (int)nbricks;
    const int BN = (p.AC % 128 == 0) ? 128 : 64;
    tp.mtiles = (p.GC + 127) / 128;
    tp.ntiles = (p.AC + BN - 1) / BN;
    const int out_tiles = p.taps.ntaps * tp.mtiles * tp.ntiles;
    // split the pixel reduction so that ~2 waves of CTAs (2 per SM) are in flight, at least 4 bricks per CTA
    int nsplit = (148 * 4 + out_tiles - 1) / out_tiles;
    const int max_split = (tp.nbricks + 3) / 4;
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 65535) nsplit = 65535;
    tp.bricks_per_split = (tp.nbricks + nsplit - 1) / nsplit;
    nsplit = (tp.nbricks + tp.bricks_per_split - 1) / tp.bricks_per_split;
    tp.dWp = p.dWp; tp.o_st = p.o_st; tp.o_sg = p.o_sg; tp.GC = p.GC; tp.AC = p.AC; tp.as = p.as; tp.taps = p.taps;
    CUtensorMap mG, mA;
    int r = encode_nhwc(&mG, p.G, p.GC, p.gw, p.gh, p.n, wt, ht, bt, 1);
    if (r) { set_error("tcgen05 wgrad: cuTensorMapEncodeTiled(G) failed (%d)", r); return VP_EUNSUPPORTED; }
    r = encode_nhwc(&mA, p.A, p.AC, p.wa, p.ha, p.n, wt, ht, bt, p.as);
    if (r) { set_error("tcgen05 wgrad: cuTensorMapEncodeTiled(A) failed (%d)", r); return VP_EUNSUPPORTED; }
    // a single pixel split writes every element exactly once: no clearing pass, plain stores (the fc layers: K = batch only)
    tp.store_only = (nsplit == 1 && p.AC % 32 == 0) ? 1 : 0;
    if (!p.accumulate && !tp.store_only) zero_async(p.dWp, sizeof(float) * (size_t)p.taps.ntaps * p.GC * p.AC, s);
    dim3 grid((unsigned)out_tiles, (unsigned)nsplit);
    // short reductions (the fc layers: K = batch = a few bricks): a 2-stage ring is enough and lets 3 CTAs share an SM, which
    // hides the per-CTA prologue (barrier init, TMEM allocation) and epilogue behind the neighbours' main loops
    if (tp.bricks_per_split <= 4) return BN == 128 ? launch_wgrad_cfg<128, 2>(mG, mA, tp, grid, s) : launch_wgrad_cfg<64, 2>(mG, mA, tp, grid, s);
    if (BN == 128) return launch_wgrad_cfg<128, 4>(mG, mA, tp, grid, s);
    return launch_wgrad_cfg<64, 4>(mG, mA, tp, grid, s);
}

}  // namespace vp

#ifdef VP_DEBUG_PROBES   // debug builds only (VP_DEBUG_PROBES=1 python -m vae_play_b200.build --force): not part of the release ABI
// =====================================================================================================
// Hardware-semantics probe (debug entry point, used by tools/probe_umma.py): does tcgen05.mma accept a K-major
// SW128 A operand whose start address is NOT 1024-byte aligned (a window into a larger TMA-written tile), with an
// arbitrary stride between 8-row groups (SBO)?  D[m][n] should equal X[off + (m/8)*sbo_rows + m%8][n] for B = I.
// =====================================================================================================
namespace vp {
namespace {
__global__ void __launch_bounds__(128) umma_probe_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapI,
                                                         float* out, int off_rows, int sbo_rows, int base_offset) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sx = smem;                 // 256 rows x 128 B
    uint8_t* si = smem + 256 * 128;     // 64 rows x 128 B (identity)
    uint64_t* bar = (uint64_t*)(si + 64 * 128);
    uint64_t* done = bar + 1;
    uint32_t* slot = (uint32_t*)(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) { mbar_init(bar, 1); mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 256 * 128 + 64 * 128);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(sx)),
                     "l"(&mapX), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(si)),
                     "l"(&mapI), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
        mbar_wait(bar, 0);
        tc_fence_after();
        uint64_t adesc = 0;
        const uint32_t sa = smem_u32(sx) + off_rows * 128;
        adesc |= (uint64_t)((sa & 0x3FFFF) >> 4);
        adesc |= (uint64_t)1 << 16;
        adesc |= (uint64_t)((sbo_rows * 128) >> 4) << 32;
        adesc |= (uint64_t)1 << 46;
        adesc |= (uint64_t)(base_offset & 7) << 49;
        adesc |= (uint64_t)2 << 61;
        const uint64_t bdesc = smem_desc_k_sw128(smem_u32(si));
        constexpr uint32_t idesc = idesc_bf16_f32(128, 64);
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, k != 0);
        tc_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory"); }
}
}  // namespace
}  // namespace vp

// x: bf16 [256][64]; ident: bf16 [64][64]; out: fp32 [128][64]
extern "C" int vp_debug_umma_probe(const void* x, const void* ident, float* out, int off_rows, int sbo_rows, int base_offset, void* stream) {
    using namespace vp;
    if (!tc_available()) { set_error("probe: no tcgen05 device"); return VP_EUNSUPPORTED; }
    EncodeTiledFn encode = get_encode();
    CUtensorMap mX, mI;
    cuuint32_t estr[2] = {1, 1};
    {
        cuuint64_t dims[2] = {64, 256}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, 256};
        if (encode(&mX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return VP_ECUDA;
    }
    {
        cuuint64_t dims[2] = {64, 64}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, 64};
        if (encode(&mI, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ident), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return VP_ECUDA;
    }
    const int smem = 256 * 128 + 64 * 128 + 64 + 1024;
    cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mX, mI, out, off_rows, sbo_rows, base_offset);
    VP_CHECK_LAUNCH("umma_probe");
    return VP_OK;
}
#endif  // VP_DEBUG_PROBES
