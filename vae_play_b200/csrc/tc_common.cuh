// Device-side helpers shared by the tcgen05 kernels: mbarrier, TMA, tcgen05.mma / commit / ld, operand descriptors.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace vp {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode();
struct TapGemm;
int encode_weight_map(CUtensorMap* m, const TapGemm& p, bool bmn, int BN);
void set_splitk_workspace(void* ptr, size_t bytes);
void* splitk_workspace(size_t bytes);
int launch_splitk_finish(const float* ws, void* D, const float* bias, int act, float slope, int64_t n, int N, bool out_f32, cudaStream_t s);
// output tensor map of one phase: pixel (gy, gx) of the phase grid -> D[n, gy*ds + doy, gx*ds + dox, :], box {32 ch, bw, bh, bb}, SWIZZLE_64B
int encode_out_map(CUtensorMap* m, void* D, int N, int hd, int wd, int n, int ds, int doy, int dox, int bw, int bh, int bb);
int pow2_floor(int v);
int pow2_ceil(int v);

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (launch error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    printf("vaeplay_b200: mbarrier wait timed out (block %d,%d thread %d parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, parity);
    __trap();
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(smem)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// TMA store of a shared-memory tile (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(smem_u32(smem)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp receives columns [col, col+32) of TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B (SBO), start address advanced by
// 32 B per UMMA_K=16 step.  Bits: [0,14) addr>>4, [16,30) LBO>>4 (unused for swizzled K-major), [32,46) SBO>>4,
// [46,48) version=1 (Blackwell), [61,64) layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}


// ---- accumulator chunk -> global memory --------------------------------------------------------------------
// v: CH fp32 accumulator columns [cbase, cbase+CH) of one output row.  The fast path (no bias, no activation, whole
// chunk inside N) is a handful of pack + 16-byte store instructions; the generic per-element path (bias, activation,
// ragged N) is only taken by the few layers that need it.  (The first version ran the generic path for every element:
// ~6000 SASS instructions per tile, which made the EPILOGUE the bottleneck of the whole kernel.)
template <int CH>
__device__ __forceinline__ void store_chunk(const uint32_t* v, void* D, int64_t row_off, int cbase, int N, const float* bias, int act,
                                            float slope, bool out_f32) {
    const bool fast = (bias == nullptr) && (act == VP_ACT_NONE) && (cbase + CH <= N) && ((N & 7) == 0);
    if (out_f32) {
        float* out = (float*)D + row_off + cbase;
        if (fast) {
#pragma unroll
            for (int j = 0; j < CH; j += 4)
                *reinterpret_cast<float4*>(out + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                                                  __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                if (cbase + j < N) {
                    float x = __uint_as_float(v[j]);
                    if (bias) x += bias[cbase + j];
                    out[j] = act_fwd(x, act, slope);
                }
            }
        }
    } else {
        bf16* out = (bf16*)D + row_off + cbase;
        if (fast) {
#pragma unroll
            for (int j = 0; j < CH; j += 8) {
                uint4 pk;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
                __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(out + j) = pk;
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < CH; ++j) {
                if (cbase + j < N) {
                    float x = __uint_as_float(v[j]);
                    if (bias) x += bias[cbase + j];
                    out[j] = __float2bfloat16_rn(act_fwd(x, act, slope));
                }
            }
        }
    }
}

// ---- accumulator chunk -> per-warp staging tile for a TMA store ---------------------------------------------------
// A warp owns 32 consecutive rows of the output tile; per 64-column group it writes them as a [32 rows][128 B]
// SWIZZLE_128B tile (1024-byte aligned; chunk j of row r at ((j ^ (r & 7)) << 4): conflict-free 16-byte stores) which ONE
// cp.async.bulk.tensor store then moves to global memory as full 128-byte lines.  (Per-thread 16-byte global stores at
// a 128-byte row stride cost one L1 wavefront per row and capped the epilogue at ~1 TB/s.)
// v: 32 fp32 columns [cbase, cbase+32) of row `lane`.
__device__ __forceinline__ void stage_chunk32(uint8_t* stile, int lane, int cbase, const uint32_t* v, const float* bias, int act, float slope) {
    const int j0 = (cbase & 63) >> 3;
    uint8_t* row = stile + lane * 128;
    const bool plain = (bias == nullptr) && (act == VP_ACT_NONE);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float x = __uint_as_float(v[j * 8 + e]);
            if (!plain) {
                if (bias) x += bias[cbase + j * 8 + e];
                x = act_fwd(x, act, slope);
            }
            f[e] = x;
        }
        uint4 pk;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(row + (((j0 + j) ^ (lane & 7)) << 4)) = pk;
    }
}

// ---- 32-column staging tiles (SWIZZLE_64B) for the persistent GEMM kernels -------------------------------------------
// [32 rows][64 B] per warp, 512-byte aligned; chunk j (16 B = 8 bf16) of row r lives at ((j ^ ((r >> 1) & 3)) << 4).
// Half the shared memory of the 64-column tiles above, which is what lets two CTAs of tapgemm_tc_kernel share an SM.
__device__ __forceinline__ void stage_chunk32_sw64(uint8_t* stile, int lane, int cbase, const uint32_t* v, const float* bias, int act, float slope) {
    uint8_t* row = stile + lane * 64;
    const bool plain = (bias == nullptr) && (act == VP_ACT_NONE);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float x = __uint_as_float(v[j * 8 + e]);
            if (!plain) {
                if (bias) x += bias[cbase + j * 8 + e];
                x = act_fwd(x, act, slope);
            }
            f[e] = x;
        }
        uint4 pk;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(row + ((j ^ ((lane >> 1) & 3)) << 4)) = pk;
    }
}
// BatchNorm statistics of the tile that is about to be stored: per-column sum and sum of squares of the staged bf16 values
// (exactly the numbers a separate pass over the stored tensor would see), added to the CTA's shared accumulators.
// Lane = (row parity h, 32-bit word w): conflict-free 4-byte reads.  row_mask: bit r set when row r of the warp's sub-tile is an
// output pixel inside the tensor (rows beyond a ragged edge are computed from partly valid inputs and must not be counted).
__device__ __forceinline__ void stats_chunk32_sw64(const uint8_t* stile, int lane, float* s_sum, float* s_sq, uint32_t row_mask) {
    const int h = lane >> 4, w = lane & 15;
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int r = 2 * i + h;
        uint32_t word = *reinterpret_cast<const uint32_t*>(stile + r * 64 + (((w >> 2) ^ (i & 3)) << 4) + (w & 3) * 4);
        word = ((row_mask >> r) & 1u) ? word : 0u;
        const float lo = __uint_as_float(word << 16), hi = __uint_as_float(word & 0xffff0000u);
        s0 += lo; q0 = fmaf(lo, lo, q0);
        s1 += hi; q1 = fmaf(hi, hi, q1);
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    q0 += __shfl_xor_sync(0xffffffffu, q0, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
    if (h == 0) {
        atomicAdd(s_sum + 2 * w, s0); atomicAdd(s_sum + 2 * w + 1, s1);
        atomicAdd(s_sq + 2 * w, q0); atomicAdd(s_sq + 2 * w + 1, q1);
    }
}

// Fused BatchNorm-backward reduction of a dgrad epilogue: `stile` holds the 32 x 32 chunk of dL/da about to be stored, `ytile` the
// matching chunk of the block's pre-norm conv output (TMA-loaded with the same box / swizzle, hence the same addresses).  Per
// channel: d = da * (y*scale + shift > 0);  sum d  and  sum d*(y - mean)  go to the CTA's shared accumulators.  sc / sh / mu
// point at this chunk's first channel.  Same lane mapping as stats_chunk32_sw64 (lane = row parity, 32-bit word).
__device__ __forceinline__ void bnred_chunk32_sw64(const uint8_t* stile, const uint8_t* ytile, int lane, float* s_sum, float* s_q, uint32_t row_mask,
                                                   const float* __restrict__ sc, const float* __restrict__ sh, const float* __restrict__ mu) {
    const int h = lane >> 4, w = lane & 15;
    const float sc0 = sc[2 * w], sc1 = sc[2 * w + 1], sh0 = sh[2 * w], sh1 = sh[2 * w + 1], mu0 = mu[2 * w], mu1 = mu[2 * w + 1];
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int r = 2 * i + h;
        const int off = r * 64 + (((w >> 2) ^ (i & 3)) << 4) + (w & 3) * 4;
        uint32_t dw = *reinterpret_cast<const uint32_t*>(stile + off);
        const uint32_t yw = *reinterpret_cast<const uint32_t*>(ytile + off);
        dw = ((row_mask >> r) & 1u) ? dw : 0u;
        const float ylo = __uint_as_float(yw << 16), yhi = __uint_as_float(yw & 0xffff0000u);
        const float dlo = fmaf(ylo, sc0, sh0) > 0.f ? __uint_as_float(dw << 16) : 0.f;
        const float dhi = fmaf(yhi, sc1, sh1) > 0.f ? __uint_as_float(dw & 0xffff0000u) : 0.f;
        s0 += dlo; q0 = fmaf(dlo, ylo - mu0, q0);
        s1 += dhi; q1 = fmaf(dhi, yhi - mu1, q1);
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    q0 += __shfl_xor_sync(0xffffffffu, q0, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
    if (h == 0) {
        atomicAdd(s_sum + 2 * w, s0); atomicAdd(s_sum + 2 * w + 1, s1);
        atomicAdd(s_q + 2 * w, q0); atomicAdd(s_q + 2 * w + 1, q1);
    }
}

// same for a 64-column [32 rows][128 B] SWIZZLE_128B staging tile: lane = 32-bit word (two channels), all 32 rows
__device__ __forceinline__ void stats_group64_sw128(const uint8_t* stile, int lane, float* s_sum, float* s_sq) {
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const uint32_t word = *reinterpret_cast<const uint32_t*>(stile + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
        const float lo = __uint_as_float(word << 16), hi = __uint_as_float(word & 0xffff0000u);
        s0 += lo; q0 = fmaf(lo, lo, q0);
        s1 += hi; q1 = fmaf(hi, hi, q1);
    }
    atomicAdd(s_sum + 2 * lane, s0); atomicAdd(s_sum + 2 * lane + 1, s1);
    atomicAdd(s_sq + 2 * lane, q0); atomicAdd(s_sq + 2 * lane + 1, q1);
}

// the fused BatchNorm-backward reduction for a 64-column [32 rows][128 B] SWIZZLE_128B staging tile (thin-layer kernels):
// lane = 32-bit word (two channels), all 32 rows; see bnred_chunk32_sw64
__device__ __forceinline__ void bnred_group64_sw128(const uint8_t* stile, const uint8_t* ytile, int lane, float* s_sum, float* s_q, uint32_t row_mask,
                                                    const float* __restrict__ sc, const float* __restrict__ sh, const float* __restrict__ mu) {
    const float sc0 = sc[2 * lane], sc1 = sc[2 * lane + 1], sh0 = sh[2 * lane], sh1 = sh[2 * lane + 1], mu0 = mu[2 * lane], mu1 = mu[2 * lane + 1];
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const int off = r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4;
        uint32_t dw = *reinterpret_cast<const uint32_t*>(stile + off);
        const uint32_t yw = *reinterpret_cast<const uint32_t*>(ytile + off);
        dw = ((row_mask >> r) & 1u) ? dw : 0u;
        const float ylo = __uint_as_float(yw << 16), yhi = __uint_as_float(yw & 0xffff0000u);
        const float dlo = fmaf(ylo, sc0, sh0) > 0.f ? __uint_as_float(dw << 16) : 0.f;
        const float dhi = fmaf(yhi, sc1, sh1) > 0.f ? __uint_as_float(dw & 0xffff0000u) : 0.f;
        s0 += dlo; q0 = fmaf(dlo, ylo - mu0, q0);
        s1 += dhi; q1 = fmaf(dhi, yhi - mu1, q1);
    }
    atomicAdd(s_sum + 2 * lane, s0); atomicAdd(s_sum + 2 * lane + 1, s1);
    atomicAdd(s_q + 2 * lane, q0); atomicAdd(s_q + 2 * lane + 1, q1);
}

struct OutMaps { CUtensorMap m[4]; };      // one output tensor map per output-parity phase (TMA-store epilogue)
constexpr int kStatMaxN = 512;             // widest layer whose BatchNorm statistics are taken in the GEMM epilogue

}  // namespace tc
}  // namespace vp
