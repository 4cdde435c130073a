// Bulk-copy (TMA 1-D) pipelined versions of the BatchNorm/activation streaming passes.
//
// The register-staged kernels in norm.cu top out near 2 TB/s: with ~100 registers per thread an SM holds only
// ~48 KB of loads in flight.  Here one elected thread streams contiguous row slabs into a 4-stage shared-memory
// ring with cp.async.bulk (completion on an mbarrier), so 64..128 KB per CTA are in flight regardless of register
// pressure, and all 256 threads only compute.  Eligible when the normalisation has one group (BatchNorm), a row is
// a multiple of 16 bytes and at most 2 KB, and the tensors are 16-byte aligned; everything else stays in norm.cu.
#include "common.cuh"

namespace vp {
namespace {

#ifndef VP_NORM_STAGES
#define VP_NORM_STAGES 4
#endif
// ring depth: 4 stages (6 measured no better: the two-tensor passes are issue-bound, not bound by bytes in flight)
constexpr int STAGES = VP_NORM_STAGES;
// one-tensor passes (stats, apply): 256 threads, 16 KB tiles, 2 CTAs/SM.  Two-tensor passes (the backward ones): 256 threads
// and 8 KB tiles per tensor, three CTAs per SM -- each thread then owns 4 rows of a tile, which halves the per-tile overhead
// (barrier wait, index arithmetic) per element against the 512-thread version (ncu: these passes are issue-bound, 0.7 warp
// instructions per element; measured -20 us per step).
#ifndef VP_NORM_TWO_NT
#define VP_NORM_TWO_NT 256
#endif
template <int MODE> struct Cfg {
    static constexpr bool TWO = (MODE == 2 || MODE == 3);
    static constexpr int NT = TWO ? VP_NORM_TWO_NT : 256;
    static constexpr int CTAS = (TWO && VP_NORM_TWO_NT == 256) ? 3 : 2;
    static constexpr int TILE_BYTES = TWO ? 8 * 1024 : 16 * 1024;   // per tensor per stage
    static constexpr int V = TWO ? 4 : 8;   // channels per thread; the backward passes carry 8 per-channel vectors
};
constexpr int VMAX = 8;

enum { M_STATS = 0, M_APPLY = 1, M_BWD_REDUCE = 2, M_BWD_APPLY = 3 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void bulk_load(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
                 "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename T, int VV> struct SV;
template <int VV> struct SV<float, VV> {
    static __device__ __forceinline__ void ld(const float* p, float* v) {
#pragma unroll
        for (int q = 0; q < VV / 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    }
    static __device__ __forceinline__ void st(float* p, const float* v) {
#pragma unroll
        for (int q = 0; q < VV / 4; ++q) *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
};
template <> struct SV<bf16, 8> {
    static __device__ __forceinline__ void ld(const bf16* p, float* v) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[2 * j] = __uint_as_float(w[j] << 16); v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
    }
    static __device__ __forceinline__ void st(bf16* p, const float* v) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]); w[j] = *reinterpret_cast<uint32_t*>(&h); }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <> struct SV<bf16, 4> {
    static __device__ __forceinline__ void ld(const bf16* p, float* v) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void st(bf16* p, const float* v) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    }
};

struct StreamArgs {
    const void* x;
    const void* da;
    void* out;                 // a (APPLY), dx (BWD_APPLY), optional dxo (BWD_REDUCE)
    const float *mean, *invstd, *scale, *shift;
    double* sums;              // [2][C]
    float *dgamma, *dbeta;
    int64_t rows;
    int C, act;
    float slope;
    int rows_per_tile;         // TILE_BYTES / row bytes
    int64_t ntiles;
};

// RELU = true compiles the activation derivative down to one compare/select per element; the generic switch in
// act_grad() costs ~4x the instructions (measured: 36 instructions per element, issue-bound at 1.7 TB/s).
template <bool RELU>
__device__ __forceinline__ float dact(float pre, int act, float slope) {
    if (RELU) return pre > 0.f ? 1.f : 0.f;
    return act_grad(pre, act, slope);
}

template <int MODE, typename T, bool RELU>
__global__ void __launch_bounds__(Cfg<MODE>::NT, Cfg<MODE>::CTAS) norm_stream_kernel(const StreamArgs p) {
    constexpr bool TWO = Cfg<MODE>::TWO;     // streams x and da
    constexpr int NT = Cfg<MODE>::NT;
    constexpr int TILE_BYTES = Cfg<MODE>::TILE_BYTES;
    constexpr int V = Cfg<MODE>::V;
    constexpr int STAGE_BYTES = TILE_BYTES * (TWO ? 2 : 1);
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    float* red = reinterpret_cast<float*>(full + STAGES);                  // [2][NT*V] for the reductions

    const int C = p.C;
    const int tpr = C / V;                        // threads per row
    const int rlanes = NT / tpr;                  // rows handled per pass over the block (tpr <= 256)
    const int cl = (threadIdx.x % tpr) * V;
    const int rl = threadIdx.x / tpr;
    const bool active = rl < rlanes;              // NT not a multiple of tpr leaves a few idle threads
    const int64_t row_bytes = (int64_t)C * sizeof(T);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_sync();

    float sc[V], sh[V], mu[V], is[V], m1[V], m2[V], a[V], b[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { sc[j] = 1.f; sh[j] = 0.f; mu[j] = 0.f; is[j] = 0.f; m1[j] = 0.f; m2[j] = 0.f; a[j] = 0.f; b[j] = 0.f; }
    if (active && MODE != M_STATS) {
        if (p.scale) { SV<float, V>::ld(p.scale + cl, sc); SV<float, V>::ld(p.shift + cl, sh); }
        if (p.mean && (MODE == M_BWD_REDUCE || MODE == M_BWD_APPLY)) { SV<float, V>::ld(p.mean + cl, mu); SV<float, V>::ld(p.invstd + cl, is); }
        if (MODE == M_BWD_APPLY) {
            const float inv_m = 1.f / (float)p.rows;
#pragma unroll
            for (int j = 0; j < V; ++j) { m1[j] = (float)p.sums[cl + j] * inv_m; m2[j] = (float)p.sums[C + cl + j] * inv_m; }
            if (blockIdx.x == 0 && rl == 0) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    if (p.dbeta) p.dbeta[cl + j] = (float)p.sums[cl + j];
                    if (p.dgamma) p.dgamma[cl + j] = (float)p.sums[C + cl + j];
                }
            }
#pragma unroll
            for (int j = 0; j < V; ++j) {       // m1 <- ka = -scale*invstd*m2,  m2 <- kb = -scale*m1 - ka*mean
                const float ka = -sc[j] * is[j] * m2[j];
                const float kb = -sc[j] * m1[j] - ka * mu[j];
                m1[j] = ka; m2[j] = kb;
            }
        }
    }

    // tiles owned by this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t my_tiles = (p.ntiles > blockIdx.x) ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](int64_t i) {   // thread 0 only
        const int s = (int)(i % STAGES);
        const int64_t tile = blockIdx.x + i * gridDim.x;
        const int64_t r0 = tile * p.rows_per_tile;
        const int64_t nrows = min((int64_t)p.rows_per_tile, p.rows - r0);
        const uint32_t bytes = (uint32_t)(nrows * row_bytes);
        uint8_t* dst = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full[s], bytes * (TWO ? 2 : 1));
        bulk_load(dst, (const uint8_t*)p.x + r0 * row_bytes, bytes, &full[s]);
        if (TWO) bulk_load(dst + TILE_BYTES, (const uint8_t*)p.da + r0 * row_bytes, bytes, &full[s]);
    };
    if (threadIdx.x == 0)
        for (int64_t i = 0; i < STAGES - 1 && i < my_tiles; ++i) issue(i);

    for (int64_t i = 0; i < my_tiles; ++i) {
        const int s = (int)(i % STAGES);
        // the stage freed at the end of iteration i-1 receives tile i+STAGES-1
        if (threadIdx.x == 0 && i + STAGES - 1 < my_tiles) issue(i + STAGES - 1);
        mbar_wait(&full[s], (uint32_t)((i / STAGES) & 1));
        const int64_t tile = blockIdx.x + i * gridDim.x;
        const int64_t r0 = tile * p.rows_per_tile;
        const int nrows = (int)min((int64_t)p.rows_per_tile, p.rows - r0);
        // 32-bit offsets inside the tile, one pointer bump per row step: the streaming passes are issue-bound, not
        // byte-bound (ncu: 0.7 warp instructions per element with 64-bit index arithmetic in the row loop)
        const T* xs = reinterpret_cast<const T*>(smem + s * STAGE_BYTES) + (rl * C + cl);
        const T* ds = reinterpret_cast<const T*>(smem + s * STAGE_BYTES + TILE_BYTES) + (rl * C + cl);
        T* outp = (MODE == M_APPLY || MODE == M_BWD_APPLY || MODE == M_BWD_REDUCE) && p.out
                      ? reinterpret_cast<T*>(p.out) + ((r0 + rl) * C + cl) : nullptr;
        const int rstep = rlanes * C;
        if (active) {
#pragma unroll 4
            for (int r = rl; r < nrows; r += rlanes, xs += rstep, ds += rstep, outp += rstep) {
                float v[V];
                SV<T, V>::ld(xs, v);
                if (MODE == M_STATS) {
#pragma unroll
                    for (int j = 0; j < V; ++j) { a[j] += v[j]; b[j] = fmaf(v[j], v[j], b[j]); }
                } else if (MODE == M_APPLY) {
#pragma unroll
                    for (int j = 0; j < V; ++j) { const float t = fmaf(v[j], sc[j], sh[j]); v[j] = RELU ? fmaxf(t, 0.f) : act_fwd(t, p.act, p.slope); }
                    SV<T, V>::st(outp, v);
                } else {
                    float d[V];
                    SV<T, V>::ld(ds, d);
                    if (MODE == M_BWD_REDUCE) {
                        // b accumulates sum d*(x - mean); the common factor invstd is applied once, after the loop
#pragma unroll
                        for (int j = 0; j < V; ++j) {
                            const float pre = fmaf(v[j], sc[j], sh[j]);
                            d[j] = RELU ? (pre > 0.f ? d[j] : 0.f) : d[j] * act_grad(pre, p.act, p.slope);
                            a[j] += d[j];
                            b[j] = fmaf(d[j], v[j] - mu[j], b[j]);
                        }
                        if (p.out) SV<T, V>::st(outp, d);
                    } else {
                        // dx = scale*(dd - m1 - xhat*m2) = scale*dd + (ka*x + kb), ka/kb per channel (held in m1/m2 from here on)
#pragma unroll
                        for (int j = 0; j < V; ++j) {
                            const float pre = fmaf(v[j], sc[j], sh[j]);
                            const float dd = RELU ? (pre > 0.f ? d[j] : 0.f) : d[j] * act_grad(pre, p.act, p.slope);
                            d[j] = fmaf(sc[j], dd, fmaf(m1[j], v[j], m2[j]));
                        }
                        SV<T, V>::st(outp, d);
                    }
                }
            }
        }
        __syncthreads();     // everyone is done with stage s before it is refilled in the next iteration
    }

    if (MODE == M_STATS || MODE == M_BWD_REDUCE) {
        const int cpb = tpr * V;   // == C
        if (active) {
#pragma unroll
            for (int j = 0; j < V; ++j) { red[rl * cpb + cl + j] = a[j]; red[NT * V + rl * cpb + cl + j] = b[j]; }
        }
        __syncthreads();
        for (int cc = threadIdx.x; cc < C; cc += NT) {
            double t1 = 0, t2 = 0;
            for (int i = 0; i < rlanes; ++i) { t1 += red[i * cpb + cc]; t2 += red[NT * V + i * cpb + cc]; }
            if (MODE == M_BWD_REDUCE && p.invstd) t2 *= (double)p.invstd[cc];
            atomicAdd(p.sums + cc, t1);
            if (MODE == M_STATS || p.mean) atomicAdd(p.sums + C + cc, t2);
        }
    }
}

template <int MODE, typename T, bool RELU>
int launch_stream(const StreamArgs& a, cudaStream_t s) {
    constexpr bool TWO = Cfg<MODE>::TWO;
    constexpr int NT = Cfg<MODE>::NT;
    constexpr int TILE_BYTES = Cfg<MODE>::TILE_BYTES;
    constexpr int V = Cfg<MODE>::V;
    constexpr int smem = STAGES * TILE_BYTES * (TWO ? 2 : 1) + STAGES * 8 + 2 * NT * V * 4 + 128;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(norm_stream_kernel<MODE, T, RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
            set_error("norm_stream: cannot set %d bytes of shared memory", smem);
            return VP_ECUDA;
        }
        attr = true;
    }
    const int ctas_per_sm = Cfg<MODE>::CTAS;       // ~97 KB / ~82 KB of smem per CTA
    int64_t grid = (int64_t)num_sms() * ctas_per_sm;
    if (grid > a.ntiles) grid = a.ntiles;
    launch_k(norm_stream_kernel<MODE, T, RELU>, dim3((unsigned)grid), dim3(NT), smem, s, a);
    VP_CHECK_LAUNCH("norm_stream");
    return VP_OK;
}

}  // namespace

// Returns VP_EUNSUPPORTED when the shape is not eligible (caller falls back to the register-staged kernels).
int norm_stream(int mode, int dtype, const void* x, const void* da, void* out, const float* mean, const float* invstd,
                const float* scale, const float* shift, double* sums, float* dgamma, float* dbeta, int64_t rows, int c, int act,
                float slope, cudaStream_t s) {
    const int es = dtype == VP_F32 ? 4 : 2;
    const int64_t row_bytes = (int64_t)c * es;
    if (c % VMAX != 0 || c / 4 > 256 || row_bytes > 2048 || rows * row_bytes < (1 << 20)) return VP_EUNSUPPORTED;
    if (((uintptr_t)x & 15) || ((uintptr_t)da & 15) || ((uintptr_t)out & 15)) return VP_EUNSUPPORTED;
    if (act & 16) {
        // "derivative from the OUTPUT" (a plain activation backward, no norm): for ReLU the output's sign is the pre-activation's,
        // so the reduce pass with scale = 1, shift = 0 is exactly dy = da * (a > 0), sums[0:C] = column sums of dy
        if (mode != M_BWD_REDUCE || (act & 15) != VP_ACT_RELU || mean || scale || !out) return VP_EUNSUPPORTED;
        act = VP_ACT_RELU;
    }
    StreamArgs a;
    a.x = x; a.da = da; a.out = out; a.mean = mean; a.invstd = invstd; a.scale = scale; a.shift = shift; a.sums = sums;
    a.dgamma = dgamma; a.dbeta = dbeta; a.rows = rows; a.C = c; a.act = act; a.slope = slope;
    const int tile_bytes = (mode == M_BWD_REDUCE || mode == M_BWD_APPLY) ? Cfg<2>::TILE_BYTES : Cfg<0>::TILE_BYTES;
    a.rows_per_tile = (int)(tile_bytes / row_bytes);
    a.ntiles = (rows + a.rows_per_tile - 1) / a.rows_per_tile;
    const bool relu = (act == VP_ACT_RELU);
#define VP_STREAM(MODE)                                                                                             \
    return dtype == VP_F32 ? (relu ? launch_stream<MODE, float, true>(a, s) : launch_stream<MODE, float, false>(a, s)) \
                           : (relu ? launch_stream<MODE, bf16, true>(a, s) : launch_stream<MODE, bf16, false>(a, s))
    switch (mode) {
        case M_STATS: VP_STREAM(M_STATS);
        case M_APPLY: VP_STREAM(M_APPLY);
        case M_BWD_REDUCE: VP_STREAM(M_BWD_REDUCE);
        default: VP_STREAM(M_BWD_APPLY);
    }
#undef VP_STREAM
}

}  // namespace vp
