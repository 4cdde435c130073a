// BatchNorm / InstanceNorm (train mode) + activation over channels-last rows, forward and backward.
// HBM-bound passes: each thread owns 4 consecutive channels of a row slab (coalesced row reads),
// per-channel partial sums are combined in shared memory and accumulated into double scratch.
#include "common.cuh"

namespace vp {
namespace {

constexpr int NT = 256;
constexpr int V = 8;                 // channels per thread: one 16-byte load of bf16, two of fp32
constexpr int CPB = 64;              // channels per block (8 threads x 8)
constexpr int RL = NT / (CPB / V);   // 32 row lanes
constexpr int ROWS_PER_BLOCK = 512;

template <typename T> __device__ __forceinline__ void ld4(const T* p, int c, int C, float v[V]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, int c, int C, float v[V]) {
    if (c + V - 1 < C && ((C & 3) == 0)) {
        const float4 t0 = *reinterpret_cast<const float4*>(p + c);
        const float4 t1 = *reinterpret_cast<const float4*>(p + c + 4);
        v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = (c + j < C) ? p[c + j] : 0.f;
    }
}
template <> __device__ __forceinline__ void ld4<bf16>(const bf16* p, int c, int C, float v[V]) {
    if (c + V - 1 < C && ((C & 7) == 0)) {
        const uint4 t = *reinterpret_cast<const uint4*>(p + c);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < V; ++j) {
            v[2 * j] = __uint_as_float(w[j] << 16);
            v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
        }
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = (c + j < C) ? __bfloat162float(p[c + j]) : 0.f;
    }
}
template <typename T> __device__ __forceinline__ void st4(T* p, int c, int C, const float v[V]);
template <> __device__ __forceinline__ void st4<float>(float* p, int c, int C, const float v[V]) {
    if (c + V - 1 < C && ((C & 3) == 0)) {
        *reinterpret_cast<float4*>(p + c) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) if (c + j < C) p[c + j] = v[j];
    }
}
template <> __device__ __forceinline__ void st4<bf16>(bf16* p, int c, int C, const float v[V]) {
    if (c + V - 1 < C && ((C & 7) == 0)) {
        uint4 t;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
        t.x = *reinterpret_cast<uint32_t*>(&h0); t.y = *reinterpret_cast<uint32_t*>(&h1);
        t.z = *reinterpret_cast<uint32_t*>(&h2); t.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(p + c) = t;
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) if (c + j < C) p[c + j] = __float2bfloat16_rn(v[j]);
    }
}

// grid: (channel tiles, row slabs per group, groups)
template <typename T>
__global__ void __launch_bounds__(NT) stats_kernel(const T* __restrict__ x, double* __restrict__ sums,
                                                   int64_t groups, int64_t rpg, int C) {
    __shared__ float s1[RL][CPB + 1], s2[RL][CPB + 1];
    const int cl = (threadIdx.x % (CPB / V)) * V;
    const int rl = threadIdx.x / (CPB / V);
    const int c = blockIdx.x * CPB + cl;
    const int64_t g = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.y * ROWS_PER_BLOCK;
    const int64_t r1 = min(r0 + ROWS_PER_BLOCK, rpg);
    float a[V] = {0, 0, 0, 0, 0, 0, 0, 0}, b[V] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c < C) {
        for (int64_t r = r0 + rl; r < r1; r += RL) {
            float v[V];
            ld4<T>(x + (g * rpg + r) * C, c, C, v);
#pragma unroll
            for (int j = 0; j < V; ++j) { a[j] += v[j]; b[j] = fmaf(v[j], v[j], b[j]); }
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) { s1[rl][cl + j] = a[j]; s2[rl][cl + j] = b[j]; }
    __syncthreads();
    if (threadIdx.x < CPB) {
        const int cc = blockIdx.x * CPB + threadIdx.x;
        if (cc < C) {
            double t1 = 0, t2 = 0;
#pragma unroll
            for (int i = 0; i < RL; ++i) { t1 += s1[i][threadIdx.x]; t2 += s2[i][threadIdx.x]; }
            atomicAdd(sums + g * C + cc, t1);
            atomicAdd(sums + (groups + g) * C + cc, t2);
        }
    }
}

__global__ void finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* running_mean, float* running_var,
                                float momentum, float eps, float* mean, float* invstd, float* scale,
                                float* shift, int64_t groups, int64_t rpg, int C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups * C) return;
    const int c = (int)(i % C);
    const double m = (double)rpg;
    const double mu = sums[i] / m;
    double var = sums[groups * C + i] / m - mu * mu;
    if (var < 0) var = 0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f;
    const float b = beta ? beta[c] : 0.f;
    mean[i] = (float)mu;
    invstd[i] = is;
    const float sc = g * is;
    scale[i] = sc;
    shift[i] = b - (float)mu * sc;
    if (running_mean && groups == 1) {
        const double unb = var * m / (m > 1 ? m - 1 : 1);
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
}

// grid: (row slabs over all rows, channel tiles)
template <typename T>
__global__ void __launch_bounds__(NT) apply_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                   const float* __restrict__ shift, T* __restrict__ a,
                                                   int64_t rows, int64_t rpg, int C, int act, float slope) {
    const int cl = (threadIdx.x % (CPB / V)) * V;
    const int rl = threadIdx.x / (CPB / V);
    const int c = blockIdx.y * CPB + cl;
    if (c >= C) return;
    const int64_t r0 = (int64_t)blockIdx.x * ROWS_PER_BLOCK;
    const int64_t r1 = min(r0 + ROWS_PER_BLOCK, rows);
    for (int64_t r = r0 + rl; r < r1; r += RL) {
        const int64_t g = r / rpg;
        float v[V], sc[V] = {1, 1, 1, 1, 1, 1, 1, 1}, sh[V] = {0, 0, 0, 0, 0, 0, 0, 0};
        ld4<T>(x + r * C, c, C, v);
        if (scale) { ld4<float>(scale + g * C, c, C, sc); ld4<float>(shift + g * C, c, C, sh); }
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = act_fwd(fmaf(v[j], sc[j], sh[j]), act, slope);
        st4<T>(a + r * C, c, C, v);
    }
}

// pass 1 of backward: d = da*act'(pre), sums of d and d*xhat.  grid like stats_kernel.
template <typename T>
__global__ void __launch_bounds__(NT) bwd_reduce_kernel(const T* __restrict__ x, const T* __restrict__ da,
                                                        const float* __restrict__ mean,
                                                        const float* __restrict__ invstd,
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ shift,
                                                        double* __restrict__ sums, T* __restrict__ dxo,
                                                        int64_t groups, int64_t rpg, int C, int act, float slope) {
    __shared__ float s1[RL][CPB + 1], s2[RL][CPB + 1];
    const int cl = (threadIdx.x % (CPB / V)) * V;
    const int rl = threadIdx.x / (CPB / V);
    const int c = blockIdx.x * CPB + cl;
    const int64_t g = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.y * ROWS_PER_BLOCK;
    const int64_t r1 = min(r0 + ROWS_PER_BLOCK, rpg);
    float a[V] = {0, 0, 0, 0, 0, 0, 0, 0}, b[V] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c < C) {
        float sc[V] = {1, 1, 1, 1, 1, 1, 1, 1}, sh[V] = {0, 0, 0, 0, 0, 0, 0, 0}, mu[V] = {0, 0, 0, 0, 0, 0, 0, 0}, is[V] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (scale) { ld4<float>(scale + g * C, c, C, sc); ld4<float>(shift + g * C, c, C, sh); }
        if (mean) { ld4<float>(mean + g * C, c, C, mu); ld4<float>(invstd + g * C, c, C, is); }
        for (int64_t r = r0 + rl; r < r1; r += RL) {
            float v[V], d[V];
            ld4<T>(x + (g * rpg + r) * C, c, C, v);
            ld4<T>(da + (g * rpg + r) * C, c, C, d);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                d[j] *= act_grad(fmaf(v[j], sc[j], sh[j]), act, slope);
                a[j] += d[j];
                b[j] = fmaf(d[j], (v[j] - mu[j]) * is[j], b[j]);
            }
            if (dxo) st4<T>(dxo + (g * rpg + r) * C, c, C, d);
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) { s1[rl][cl + j] = a[j]; s2[rl][cl + j] = b[j]; }
    __syncthreads();
    if (threadIdx.x < CPB) {
        const int cc = blockIdx.x * CPB + threadIdx.x;
        if (cc < C) {
            double t1 = 0, t2 = 0;
#pragma unroll
            for (int i = 0; i < RL; ++i) { t1 += s1[i][threadIdx.x]; t2 += s2[i][threadIdx.x]; }
            atomicAdd(sums + g * C + cc, t1);
            if (mean) atomicAdd(sums + (groups + g) * C + cc, t2);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(NT) bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ da,
                                                       const float* __restrict__ mean,
                                                       const float* __restrict__ invstd,
                                                       const float* __restrict__ scale,
                                                       const float* __restrict__ shift,
                                                       const double* __restrict__ sums, T* __restrict__ dx,
                                                       float* dgamma, float* dbeta, int64_t groups, int64_t rpg,
                                                       int C, int act, float slope) {
    const int cl = (threadIdx.x % (CPB / V)) * V;
    const int rl = threadIdx.x / (CPB / V);
    const int c = blockIdx.y * CPB + cl;
    if (c >= C) return;
    const int64_t rows = groups * rpg;
    const int64_t r0 = (int64_t)blockIdx.x * ROWS_PER_BLOCK;
    const int64_t r1 = min(r0 + ROWS_PER_BLOCK, rows);
    const float inv_m = 1.f / (float)rpg;
    if (blockIdx.x == 0 && rl == 0 && groups == 1) {
#pragma unroll
        for (int j = 0; j < V; ++j)
            if (c + j < C) {
                if (dbeta) dbeta[c + j] = (float)sums[c + j];
                if (dgamma) dgamma[c + j] = (float)sums[C + c + j];
            }
    }
    for (int64_t r = r0 + rl; r < r1; r += RL) {
        const int64_t g = r / rpg;
        float sc[V], sh[V], mu[V], is[V], m1[V], m2[V], v[V], d[V];
        ld4<float>(scale + g * C, c, C, sc);
        ld4<float>(shift + g * C, c, C, sh);
        ld4<float>(mean + g * C, c, C, mu);
        ld4<float>(invstd + g * C, c, C, is);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            m1[j] = (c + j < C) ? (float)sums[g * C + c + j] * inv_m : 0.f;
            m2[j] = (c + j < C) ? (float)sums[(groups + g) * C + c + j] * inv_m : 0.f;
        }
        ld4<T>(x + r * C, c, C, v);
        ld4<T>(da + r * C, c, C, d);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float dd = d[j] * act_grad(fmaf(v[j], sc[j], sh[j]), act, slope);
            const float xh = (v[j] - mu[j]) * is[j];
            d[j] = sc[j] * (dd - m1[j] - xh * m2[j]);
        }
        st4<T>(dx + r * C, c, C, d);
    }
}

__global__ void colsum_finish_kernel(const double* __restrict__ s, float* __restrict__ out, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) out[i] = (float)s[i];
}

}  // namespace
}  // namespace vp

using namespace vp;

extern "C" int vp_norm_stats(const void* x, double* sums, int dtype, int64_t groups, int64_t rpg, int c,
                             void* stream) {
    VP_CHECK_ARG(x && sums && groups > 0 && rpg > 0 && c > 0, "vp_norm_stats: bad arguments");
    VP_CHECK_ARG(groups <= 65535, "vp_norm_stats: too many groups");
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * groups * c, (cudaStream_t)stream);
    dim3 grid((c + CPB - 1) / CPB, (unsigned)((rpg + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), (unsigned)groups);
    if (dtype == VP_F32) stats_kernel<float><<<grid, NT, 0, (cudaStream_t)stream>>>((const float*)x, sums, groups, rpg, c);
    else stats_kernel<bf16><<<grid, NT, 0, (cudaStream_t)stream>>>((const bf16*)x, sums, groups, rpg, c);
    VP_CHECK_LAUNCH("vp_norm_stats");
    return VP_OK;
}

extern "C" int vp_norm_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, float momentum, float eps, float* mean, float* invstd,
                                float* scale, float* shift, int64_t groups, int64_t rpg, int c, void* stream) {
    VP_CHECK_ARG(sums && mean && invstd && scale && shift && groups > 0 && rpg > 0 && c > 0,
                 "vp_norm_finalize: bad arguments");
    const int64_t n = groups * c;
    finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        sums, gamma, beta, running_mean, running_var, momentum, eps, mean, invstd, scale, shift, groups, rpg, c);
    VP_CHECK_LAUNCH("vp_norm_finalize");
    return VP_OK;
}

extern "C" int vp_norm_apply_act(const void* x, const float* scale, const float* shift, void* a, int dtype,
                                 int64_t groups, int64_t rpg, int c, int act, float slope, void* stream) {
    VP_CHECK_ARG(x && a && groups > 0 && rpg > 0 && c > 0, "vp_norm_apply_act: bad arguments");
    const int64_t rows = groups * rpg;
    dim3 grid((unsigned)((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), (c + CPB - 1) / CPB);
    if (dtype == VP_F32)
        apply_kernel<float><<<grid, NT, 0, (cudaStream_t)stream>>>((const float*)x, scale, shift, (float*)a, rows, rpg, c, act, slope);
    else
        apply_kernel<bf16><<<grid, NT, 0, (cudaStream_t)stream>>>((const bf16*)x, scale, shift, (bf16*)a, rows, rpg, c, act, slope);
    VP_CHECK_LAUNCH("vp_norm_apply_act");
    return VP_OK;
}

extern "C" int vp_norm_bwd_reduce(const void* x, const void* da, const float* mean, const float* invstd,
                                  const float* scale, const float* shift, double* sums, void* dxo, int dtype,
                                  int64_t groups, int64_t rpg, int c, int act, float slope, void* stream) {
    VP_CHECK_ARG(x && da && sums && groups > 0 && rpg > 0 && c > 0, "vp_norm_bwd_reduce: bad arguments");
    VP_CHECK_ARG(groups <= 65535, "vp_norm_bwd_reduce: too many groups");
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * groups * c, (cudaStream_t)stream);
    dim3 grid((c + CPB - 1) / CPB, (unsigned)((rpg + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), (unsigned)groups);
    if (dtype == VP_F32)
        bwd_reduce_kernel<float><<<grid, NT, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)da, mean, invstd, scale, shift, sums, (float*)dxo, groups, rpg, c, act, slope);
    else
        bwd_reduce_kernel<bf16><<<grid, NT, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)da, mean, invstd, scale, shift, sums, (bf16*)dxo, groups, rpg, c, act, slope);
    VP_CHECK_LAUNCH("vp_norm_bwd_reduce");
    return VP_OK;
}

extern "C" int vp_norm_bwd_apply(const void* x, const void* da, const float* mean, const float* invstd,
                                 const float* scale, const float* shift, const double* sums, void* dx,
                                 float* dgamma, float* dbeta, int dtype, int64_t groups, int64_t rpg, int c,
                                 int act, float slope, void* stream) {
    VP_CHECK_ARG(x && da && mean && invstd && scale && shift && sums && dx && groups > 0 && rpg > 0 && c > 0,
                 "vp_norm_bwd_apply: bad arguments");
    const int64_t rows = groups * rpg;
    dim3 grid((unsigned)((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), (c + CPB - 1) / CPB);
    if (dtype == VP_F32)
        bwd_apply_kernel<float><<<grid, NT, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)da, mean, invstd, scale, shift, sums, (float*)dx, dgamma, dbeta, groups, rpg, c, act, slope);
    else
        bwd_apply_kernel<bf16><<<grid, NT, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)da, mean, invstd, scale, shift, sums, (bf16*)dx, dgamma, dbeta, groups, rpg, c, act, slope);
    VP_CHECK_LAUNCH("vp_norm_bwd_apply");
    return VP_OK;
}

extern "C" int vp_colsum(const void* x, float* out, double* scratch_c, int dtype, int64_t rows, int c, void* stream) {
    VP_CHECK_ARG(x && out && scratch_c && rows > 0 && c > 0, "vp_colsum: bad arguments");
    // scratch_c: double [2][c], zero on entry; reuse the statistics kernel (sum column)
    int rc = vp_norm_stats(x, scratch_c, dtype, 1, rows, c, stream);
    if (rc) return rc;
    colsum_finish_kernel<<<(c + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scratch_c, out, c);
    VP_CHECK_LAUNCH("vp_colsum");
    return VP_OK;
}
