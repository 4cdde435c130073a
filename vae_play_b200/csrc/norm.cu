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

template <typename T, int VV> struct Vec;
template <int VV> struct Vec<float, VV> {
    static __device__ __forceinline__ void ld(const float* p, int c, int C, float* v) {
        if (c + VV - 1 < C && ((C & 3) == 0)) {
#pragma unroll
            for (int q = 0; q < VV / 4; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(p + c + 4 * q);
                v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < VV; ++j) v[j] = (c + j < C) ? p[c + j] : 0.f;
        }
    }
    static __device__ __forceinline__ void st(float* p, int c, int C, const float* v) {
        if (c + VV - 1 < C && ((C & 3) == 0)) {
#pragma unroll
            for (int q = 0; q < VV / 4; ++q)
                *reinterpret_cast<float4*>(p + c + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < VV; ++j) if (c + j < C) p[c + j] = v[j];
        }
    }
};
template <int VV> struct Vec<bf16, VV> {
    static __device__ __forceinline__ void ld(const bf16* p, int c, int C, float* v) {
        if (c + VV - 1 < C && ((C & (VV - 1)) == 0)) {
            uint32_t w[VV / 2];
            if (VV == 8) {
                const uint4 t = *reinterpret_cast<const uint4*>(p + c);
                w[0] = t.x; w[1] = t.y; w[VV / 2 - 2] = t.z; w[VV / 2 - 1] = t.w;
            } else {
                const uint2 t = *reinterpret_cast<const uint2*>(p + c);
                w[0] = t.x; w[1] = t.y;
            }
#pragma unroll
            for (int j = 0; j < VV / 2; ++j) {
                v[2 * j] = __uint_as_float(w[j] << 16);
                v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
            }
        } else {
#pragma unroll
            for (int j = 0; j < VV; ++j) v[j] = (c + j < C) ? __bfloat162float(p[c + j]) : 0.f;
        }
    }
    static __device__ __forceinline__ void st(bf16* p, int c, int C, const float* v) {
        if (c + VV - 1 < C && ((C & (VV - 1)) == 0)) {
            uint32_t w[VV / 2];
#pragma unroll
            for (int j = 0; j < VV / 2; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                w[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            if (VV == 8) *reinterpret_cast<uint4*>(p + c) = make_uint4(w[0], w[1], w[VV / 2 - 2], w[VV / 2 - 1]);
            else *reinterpret_cast<uint2*>(p + c) = make_uint2(w[0], w[1]);
        } else {
#pragma unroll
            for (int j = 0; j < VV; ++j) if (c + j < C) p[c + j] = __float2bfloat16_rn(v[j]);
        }
    }
};
template <typename T> __device__ __forceinline__ void ld4(const T* p, int c, int C, float* v) { Vec<T, V>::ld(p, c, C, v); }
template <typename T> __device__ __forceinline__ void st4(T* p, int c, int C, const float* v) { Vec<T, V>::st(p, c, C, v); }

constexpr int UNR = 4;   // rows in flight per thread (16-byte loads issued back to back before use)

// Reduction kernels use 4 channels per thread (8-byte bf16 loads) to stay at <= 64 registers: four 256-thread
// blocks per SM with 4 rows in flight per thread keeps ~64 KB of loads outstanding per SM.
constexpr int VR = 4;

// grid: (channel tiles, row slabs per group, groups)
template <typename T>
__global__ void __launch_bounds__(NT, 4) stats_kernel(const T* __restrict__ x, double* __restrict__ sums,
                                                      int64_t groups, int64_t rpg, int C, int tpr) {
    pdl_sync();
    // tpr = threads per row (power of two <= 16): narrow tensors (C = 1, 3, ...) put more threads on the row axis
    __shared__ float s1[NT * VR], s2[NT * VR];
    const int cpb = tpr * VR;            // channels per block
    const int rlanes = NT / tpr;         // row lanes
    const int cl = (threadIdx.x % tpr) * VR;
    const int rl = threadIdx.x / tpr;
    const int c = blockIdx.x * cpb + cl;
    const int64_t g = blockIdx.z;
    float a[VR] = {0, 0, 0, 0}, b[VR] = {0, 0, 0, 0};
    // each block walks several row slabs: few blocks => few same-address double atomics at the end
    for (int64_t r0 = (int64_t)blockIdx.y * ROWS_PER_BLOCK; r0 < rpg && c < C; r0 += (int64_t)gridDim.y * ROWS_PER_BLOCK) {
        const int64_t r1 = min(r0 + ROWS_PER_BLOCK, rpg);
        const T* base = x + g * rpg * C;
        for (int64_t r = r0 + rl; r < r1; r += (int64_t)rlanes * UNR) {
            float v[UNR][VR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int64_t rr = r + (int64_t)u * rlanes;
                if (rr < r1) Vec<T, VR>::ld(base + rr * C, c, C, v[u]);
                else {
#pragma unroll
                    for (int j = 0; j < VR; ++j) v[u][j] = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u)
#pragma unroll
                for (int j = 0; j < VR; ++j) { a[j] += v[u][j]; b[j] = fmaf(v[u][j], v[u][j], b[j]); }
        }
    }
#pragma unroll
    for (int j = 0; j < VR; ++j) { s1[rl * cpb + cl + j] = a[j]; s2[rl * cpb + cl + j] = b[j]; }
    __syncthreads();
    if ((int)threadIdx.x < cpb) {
        const int cc = blockIdx.x * cpb + threadIdx.x;
        if (cc < C) {
            double t1 = 0, t2 = 0;
            for (int i = 0; i < rlanes; ++i) { t1 += s1[i * cpb + threadIdx.x]; t2 += s2[i * cpb + threadIdx.x]; }
            atomicAdd(sums + g * C + cc, t1);
            atomicAdd(sums + (groups + g) * C + cc, t2);
        }
    }
}

__global__ void finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* running_mean, float* running_var,
                                float momentum, float eps, float* mean, float* invstd, float* scale,
                                float* shift, int64_t groups, int64_t rpg, int C, long long* nbt) {
    pdl_sync();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && nbt) *nbt += 1;          // BatchNorm.num_batches_tracked, kept on the device by torch
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

// grid: (row slabs per group, channel tiles, groups): per-channel parameters are loaded once per thread
template <typename T>
__global__ void __launch_bounds__(NT) apply_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                   const float* __restrict__ shift, T* __restrict__ a,
                                                   int64_t rpg, int C, int act, float slope) {
    pdl_sync();
    const int cl = (threadIdx.x % (CPB / V)) * V;
    const int rl = threadIdx.x / (CPB / V);
    const int c = blockIdx.y * CPB + cl;
    if (c >= C) return;
    const int64_t g = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.x * ROWS_PER_BLOCK;
    const int64_t r1 = min(r0 + ROWS_PER_BLOCK, rpg);
    float sc[V] = {1, 1, 1, 1, 1, 1, 1, 1}, sh[V] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (scale) { ld4<float>(scale + g * C, c, C, sc); ld4<float>(shift + g * C, c, C, sh); }
    const T* xb = x + g * rpg * C;
    T* ab = a + g * rpg * C;
    for (int64_t r = r0 + rl; r < r1; r += RL * UNR) {
        float v[UNR][V];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t rr = r + (int64_t)u * RL;
            if (rr < r1) ld4<T>(xb + rr * C, c, C, v[u]);
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t rr = r + (int64_t)u * RL;
            if (rr < r1) {
#pragma unroll
                for (int j = 0; j < V; ++j) v[u][j] = act_fwd(fmaf(v[u][j], sc[j], sh[j]), act, slope);
                st4<T>(ab + rr * C, c, C, v[u]);
            }
        }
    }
}

// pass 1 of backward: d = da*act'(pre), sums of d and d*xhat.  grid like stats_kernel.
template <typename T>
__global__ void __launch_bounds__(NT, 3) bwd_reduce_kernel(const T* __restrict__ x, const T* __restrict__ da,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ invstd,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           double* __restrict__ sums, T* __restrict__ dxo,
                                                           int64_t groups, int64_t rpg, int C, int act, float slope, int tpr) {
    pdl_sync();
    __shared__ float s1[NT * VR], s2[NT * VR];
    const int cpb = tpr * VR;
    const int rlanes = NT / tpr;
    const int cl = (threadIdx.x % tpr) * VR;
    const int rl = threadIdx.x / tpr;
    const int c = blockIdx.x * cpb + cl;
    const int64_t g = blockIdx.z;
    float a[VR] = {0, 0, 0, 0}, b[VR] = {0, 0, 0, 0};
    float sc[VR] = {1, 1, 1, 1}, sh[VR] = {0, 0, 0, 0}, mu[VR] = {0, 0, 0, 0}, is[VR] = {0, 0, 0, 0};
    if (c < C) {
        if (scale) { Vec<float, VR>::ld(scale + g * C, c, C, sc); Vec<float, VR>::ld(shift + g * C, c, C, sh); }
        if (mean) { Vec<float, VR>::ld(mean + g * C, c, C, mu); Vec<float, VR>::ld(invstd + g * C, c, C, is); }
    }
    for (int64_t r0 = (int64_t)blockIdx.y * ROWS_PER_BLOCK; r0 < rpg && c < C; r0 += (int64_t)gridDim.y * ROWS_PER_BLOCK) {
        const int64_t r1 = min(r0 + ROWS_PER_BLOCK, rpg);
        const T* xb = x + g * rpg * C;
        const T* db = da + g * rpg * C;
        for (int64_t r = r0 + rl; r < r1; r += (int64_t)rlanes * UNR) {
            float v[UNR][VR], d[UNR][VR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int64_t rr = r + (int64_t)u * rlanes;
                if (rr < r1) { Vec<T, VR>::ld(xb + rr * C, c, C, v[u]); Vec<T, VR>::ld(db + rr * C, c, C, d[u]); }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int64_t rr = r + (int64_t)u * rlanes;
                if (rr < r1) {
#pragma unroll
                    for (int j = 0; j < VR; ++j) {
                        d[u][j] *= act_grad(fmaf(v[u][j], sc[j], sh[j]), act, slope);
                        a[j] += d[u][j];
                        b[j] = fmaf(d[u][j], (v[u][j] - mu[j]) * is[j], b[j]);
                    }
                    if (dxo) Vec<T, VR>::st(dxo + (g * rpg + rr) * C, c, C, d[u]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < VR; ++j) { s1[rl * cpb + cl + j] = a[j]; s2[rl * cpb + cl + j] = b[j]; }
    __syncthreads();
    if ((int)threadIdx.x < cpb) {
        const int cc = blockIdx.x * cpb + threadIdx.x;
        if (cc < C) {
            double t1 = 0, t2 = 0;
            for (int i = 0; i < rlanes; ++i) { t1 += s1[i * cpb + threadIdx.x]; t2 += s2[i * cpb + threadIdx.x]; }
            atomicAdd(sums + g * C + cc, t1);
            if (mean) atomicAdd(sums + (groups + g) * C + cc, t2);
        }
    }
}

// pass 2: grid (row slabs per group, channel tiles, groups)
template <typename T>
__global__ void __launch_bounds__(NT) bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ da,
                                                       const float* __restrict__ mean,
                                                       const float* __restrict__ invstd,
                                                       const float* __restrict__ scale,
                                                       const float* __restrict__ shift,
                                                       const double* __restrict__ sums, T* __restrict__ dx,
                                                       float* dgamma, float* dbeta, int64_t groups, int64_t rpg,
                                                       int C, int act, float slope) {
    pdl_sync();
    const int cl = (threadIdx.x % (CPB / V)) * V;
    const int rl = threadIdx.x / (CPB / V);
    const int c = blockIdx.y * CPB + cl;
    if (c >= C) return;
    const int64_t g = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.x * ROWS_PER_BLOCK;
    const int64_t r1 = min(r0 + ROWS_PER_BLOCK, rpg);
    const float inv_m = 1.f / (float)rpg;
    float sc[V], sh[V], mu[V], is[V], m1[V], m2[V];
    ld4<float>(scale + g * C, c, C, sc);
    ld4<float>(shift + g * C, c, C, sh);
    ld4<float>(mean + g * C, c, C, mu);
    ld4<float>(invstd + g * C, c, C, is);
#pragma unroll
    for (int j = 0; j < V; ++j) {
        m1[j] = (c + j < C) ? (float)sums[g * C + c + j] * inv_m : 0.f;
        m2[j] = (c + j < C) ? (float)sums[(groups + g) * C + c + j] * inv_m : 0.f;
    }
    if (blockIdx.x == 0 && rl == 0 && groups == 1) {
#pragma unroll
        for (int j = 0; j < V; ++j)
            if (c + j < C) {
                if (dbeta) dbeta[c + j] = (float)sums[c + j];
                if (dgamma) dgamma[c + j] = (float)sums[C + c + j];
            }
    }
    const T* xb = x + g * rpg * C;
    const T* db = da + g * rpg * C;
    T* ob = dx + g * rpg * C;
    for (int64_t r = r0 + rl; r < r1; r += RL * (UNR / 2)) {
        float v[UNR / 2][V], d[UNR / 2][V];
#pragma unroll
        for (int u = 0; u < UNR / 2; ++u) {
            const int64_t rr = r + (int64_t)u * RL;
            if (rr < r1) { ld4<T>(xb + rr * C, c, C, v[u]); ld4<T>(db + rr * C, c, C, d[u]); }
        }
#pragma unroll
        for (int u = 0; u < UNR / 2; ++u) {
            const int64_t rr = r + (int64_t)u * RL;
            if (rr < r1) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float dd = d[u][j] * act_grad(fmaf(v[u][j], sc[j], sh[j]), act, slope);
                    const float xh = (v[u][j] - mu[j]) * is[j];
                    d[u][j] = sc[j] * (dd - m1[j] - xh * m2[j]);
                }
                st4<T>(ob + rr * C, c, C, d[u]);
            }
        }
    }
}

__global__ void colsum_finish_kernel(const double* __restrict__ s, float* __restrict__ out, int C) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) out[i] = (float)s[i];
}

}  // namespace
}  // namespace vp

namespace vp {
int norm_stream(int mode, int dtype, const void* x, const void* da, void* out, const float* mean, const float* invstd,
                const float* scale, const float* shift, double* sums, float* dgamma, float* dbeta, int64_t rows, int c, int act,
                float slope, cudaStream_t s);
}
using namespace vp;

// threads per row of the reduction kernels: enough 4-channel threads to cover C, at most 16 (64 channels per block)
static int threads_per_row(int c) {
    int t = 1;
    while (t < 16 && t * VR < c) t *= 2;
    return t;
}

// number of row-slab blocks of the reduction kernels: about 4 CTAs per SM in total
static unsigned slab_blocks(int64_t rpg, int ctiles, int64_t groups) {
    const int64_t slabs = (rpg + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
    int64_t cap = (148 * 4) / ((int64_t)ctiles * groups);
    if (cap < 1) cap = 1;
    return (unsigned)(slabs < cap ? slabs : cap);
}

extern "C" int vp_norm_stats(const void* x, double* sums, int dtype, int64_t groups, int64_t rpg, int c,
                             void* stream) {
    VP_CHECK_ARG(x && sums && groups > 0 && rpg > 0 && c > 0, "vp_norm_stats: bad arguments");
    VP_CHECK_ARG(groups <= 65535, "vp_norm_stats: too many groups");
    zero_async(sums, sizeof(double) * 2 * groups * c, (cudaStream_t)stream);
    if (groups == 1) {
        const int rc = norm_stream(0, dtype, x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, sums, nullptr, nullptr, rpg, c, 0,
                                   0.f, (cudaStream_t)stream);
        if (rc != VP_EUNSUPPORTED) return rc;
    }
    const int tpr = threads_per_row(c);
    const int ctiles = (c + tpr * VR - 1) / (tpr * VR);
    dim3 grid(ctiles, slab_blocks(rpg, ctiles, groups), (unsigned)groups);
    if (dtype == VP_F32) launch_k(stats_kernel<float>, dim3(grid), dim3(NT), 0, (cudaStream_t)stream, (const float*)x, sums, groups, rpg, c, tpr);
    else launch_k(stats_kernel<bf16>, dim3(grid), dim3(NT), 0, (cudaStream_t)stream, (const bf16*)x, sums, groups, rpg, c, tpr);
    VP_CHECK_LAUNCH("vp_norm_stats");
    return VP_OK;
}

extern "C" int vp_norm_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, float momentum, float eps, float* mean, float* invstd,
                                float* scale, float* shift, int64_t groups, int64_t rpg, int c, int64_t* num_batches_tracked,
                                void* stream) {
    VP_CHECK_ARG(sums && mean && invstd && scale && shift && groups > 0 && rpg > 0 && c > 0,
                 "vp_norm_finalize: bad arguments");
    const int64_t n = groups * c;
    launch_k(finalize_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
        sums, gamma, beta, running_mean, running_var, momentum, eps, mean, invstd, scale, shift, groups, rpg, c, (long long*)num_batches_tracked);
    VP_CHECK_LAUNCH("vp_norm_finalize");
    return VP_OK;
}

namespace vp {
namespace {
// block = 16 channels x 64 part lanes: every part row is read by one lane, all loads of a thread are independent
__global__ void __launch_bounds__(1024) finalize_parts_kernel(const float* __restrict__ parts, int nparts, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* running_mean, float* running_var, float momentum,
                                                             float eps, float* mean, float* invstd, float* scale, float* shift, int64_t rows, int C,
                                                             long long* nbt) {
    pdl_sync();
    if (nbt && blockIdx.x == 0 && threadIdx.x == 0) *nbt += 1;
    __shared__ double sh1[64][17], sh2[64][17];
    const int cl = threadIdx.x & 15, pl = threadIdx.x >> 4;
    const int c = blockIdx.x * 16 + cl;
    double a1 = 0, a2 = 0;
    if (c < C) {
#pragma unroll 5
        for (int i = pl; i < nparts; i += 64) { a1 += (double)parts[(size_t)i * 2 * C + c]; a2 += (double)parts[(size_t)i * 2 * C + C + c]; }
    }
    sh1[pl][cl] = a1; sh2[pl][cl] = a2;
    __syncthreads();
    if (pl != 0 || c >= C) return;
    double s1 = 0, s2 = 0;
#pragma unroll 8
    for (int i = 0; i < 64; ++i) { s1 += sh1[i][cl]; s2 += sh2[i][cl]; }
    const double m = (double)rows;
    const double mu = s1 / m;
    double var = s2 / m - mu * mu;
    if (var < 0) var = 0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f;
    const float b = beta ? beta[c] : 0.f;
    mean[c] = (float)mu;
    invstd[c] = is;
    const float sc = g * is;
    scale[c] = sc;
    shift[c] = b - (float)mu * sc;
    if (running_mean) {
        const double unb = var * m / (m > 1 ? m - 1 : 1);
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
}
}  // namespace
}  // namespace vp

namespace vp {
namespace {
// parts[nparts][2][C] fp32 (sum d, sum d*(x - mean)) -> sums[2][C] double = (sum d, invstd * sum d*(x - mean)), fixed order
__global__ void __launch_bounds__(256) bwd_finish_parts_kernel(const float* __restrict__ parts, int nparts, const float* __restrict__ invstd,
                                                               double* __restrict__ sums, int C) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double a = 0, b = 0;
    for (int i = 0; i < nparts; ++i) { a += (double)parts[(size_t)i * 2 * C + c]; b += (double)parts[(size_t)i * 2 * C + C + c]; }
    sums[c] = a;
    sums[C + c] = b * (double)invstd[c];
}
}  // namespace
}  // namespace vp

/* The `sums` of vp_norm_bwd_reduce (double [2][c]: sum d, sum d*xhat) from the per-CTA partial sums a fused data-gradient
 * epilogue delivered (vp_conv_dgrad_cl_bnred): added in double in a fixed order, the common factor invstd applied here. */
extern "C" int vp_norm_bwd_finish_parts(const float* parts, int nparts, const float* invstd, double* sums, int c, void* stream) {
    VP_CHECK_ARG(parts && nparts > 0 && invstd && sums && c > 0, "vp_norm_bwd_finish_parts: bad arguments");
    launch_k(bwd_finish_parts_kernel, dim3((c + 255) / 256), dim3(256), 0, (cudaStream_t)stream, parts, nparts, invstd, sums, c);
    VP_CHECK_LAUNCH("vp_norm_bwd_finish_parts");
    return VP_OK;
}

/* vp_norm_finalize for BatchNorm statistics delivered as per-CTA partial sums by a GEMM epilogue (vp_conv_fwd_cl_stats,
 * vp_thin_conv_fwd_stats): parts[nparts][2][c] fp32, added in double in a fixed order (deterministic). */
extern "C" int vp_norm_finalize_parts(const float* parts, int nparts, const float* gamma, const float* beta, float* running_mean,
                                      float* running_var, float momentum, float eps, float* mean, float* invstd, float* scale,
                                      float* shift, int64_t rows, int c, int64_t* num_batches_tracked, void* stream) {
    VP_CHECK_ARG(parts && nparts > 0 && mean && invstd && scale && shift && rows > 0 && c > 0, "vp_norm_finalize_parts: bad arguments");
    launch_k(finalize_parts_kernel, dim3((c + 15) / 16), dim3(1024), 0, (cudaStream_t)stream, parts, nparts, gamma, beta, running_mean, running_var, momentum, eps,
                                                                         mean, invstd, scale, shift, rows, c, (long long*)num_batches_tracked);
    VP_CHECK_LAUNCH("vp_norm_finalize_parts");
    return VP_OK;
}

extern "C" int vp_norm_apply_act(const void* x, const float* scale, const float* shift, void* a, int dtype,
                                 int64_t groups, int64_t rpg, int c, int act, float slope, void* stream) {
    VP_CHECK_ARG(x && a && groups > 0 && rpg > 0 && c > 0, "vp_norm_apply_act: bad arguments");
    VP_CHECK_ARG(groups <= 65535, "vp_norm_apply_act: too many groups");
    if (groups == 1) {
        const int rc = norm_stream(1, dtype, x, nullptr, a, nullptr, nullptr, scale, shift, nullptr, nullptr, nullptr, rpg, c, act, slope,
                                   (cudaStream_t)stream);
        if (rc != VP_EUNSUPPORTED) return rc;
    }
    dim3 grid((unsigned)((rpg + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), (c + CPB - 1) / CPB, (unsigned)groups);
    if (dtype == VP_F32)
        launch_k(apply_kernel<float>, dim3(grid), dim3(NT), 0, (cudaStream_t)stream, (const float*)x, scale, shift, (float*)a, rpg, c, act, slope);
    else
        launch_k(apply_kernel<bf16>, dim3(grid), dim3(NT), 0, (cudaStream_t)stream, (const bf16*)x, scale, shift, (bf16*)a, rpg, c, act, slope);
    VP_CHECK_LAUNCH("vp_norm_apply_act");
    return VP_OK;
}

namespace vp {
namespace {
template <typename T> struct AccT { typedef float type; };
template <> struct AccT<float> { typedef double type; };   // fp32 check mode: double accumulation

template <typename T>
__device__ __forceinline__ void ldv4(const T* p, float* v);
template <> __device__ __forceinline__ void ldv4<float>(const float* p, float* v) {
    const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ldv4<bf16>(const bf16* p, float* v) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}
template <typename T>
__device__ __forceinline__ void stv4(T* p, const float* v);
template <> __device__ __forceinline__ void stv4<float>(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
template <> __device__ __forceinline__ void stv4<bf16>(bf16* p, const float* v) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
}

// Activation backward of a layer WITHOUT normalisation and with very few channels (the decoder's 1-channel output conv:
// 1 M rows x 1 channel): d = da * act'(.), per-channel sum(d) for the bias gradient.  The tensor is walked as a flat
// array, 8 elements per thread; element i belongs to channel i % C (C <= 4).
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_flat_kernel(const T* __restrict__ x, const T* __restrict__ da, T* __restrict__ dxo, double* sums,
                                                           int64_t n8, int C, int act, float slope) {
    pdl_sync();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        float v[8], d[8];
        ldv4<T>(x + i * 8, v); ldv4<T>(x + i * 8 + 4, v + 4);
        ldv4<T>(da + i * 8, d); ldv4<T>(da + i * 8 + 4, d + 4);
        const int c0 = (int)((i * 8) % C);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            d[j] *= act_grad(v[j], act, slope);
            const int c = (c0 + j) % C;
            acc[0] += c == 0 ? d[j] : 0.f; acc[1] += c == 1 ? d[j] : 0.f; acc[2] += c == 2 ? d[j] : 0.f; acc[3] += c == 3 ? d[j] : 0.f;
        }
        stv4<T>(dxo + i * 8, d); stv4<T>(dxo + i * 8 + 4, d + 4);
    }
    __shared__ float red[8][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = warp_sum(acc[c]);
    if ((threadIdx.x & 31) == 0)
        for (int c = 0; c < 4; ++c) red[threadIdx.x >> 5][c] = acc[c];
    __syncthreads();
    if ((int)threadIdx.x < C) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        atomicAdd(sums + threadIdx.x, t);
    }
}
}  // namespace
}  // namespace vp

extern "C" int vp_norm_bwd_reduce(const void* x, const void* da, const float* mean, const float* invstd,
                                  const float* scale, const float* shift, double* sums, void* dxo, int dtype,
                                  int64_t groups, int64_t rpg, int c, int act, float slope, void* stream) {
    VP_CHECK_ARG(x && da && sums && groups > 0 && rpg > 0 && c > 0, "vp_norm_bwd_reduce: bad arguments");
    VP_CHECK_ARG(groups <= 65535, "vp_norm_bwd_reduce: too many groups");
    zero_async(sums, sizeof(double) * 2 * groups * c, (cudaStream_t)stream);
    if (groups == 1 && !mean && !scale && dxo && c <= 4 && (rpg * c) % 8 == 0 && (((uintptr_t)x | (uintptr_t)da | (uintptr_t)dxo) & 15) == 0) {
        const int64_t n8 = rpg * c / 8;
        int64_t blocks = (n8 + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        if (dtype == VP_F32) launch_k(act_bwd_flat_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const float*)x, (const float*)da, (float*)dxo, sums, n8, c, act, slope);
        else launch_k(act_bwd_flat_kernel<bf16>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)da, (bf16*)dxo, sums, n8, c, act, slope);
        VP_CHECK_LAUNCH("vp_norm_bwd_reduce");
        return VP_OK;
    }
    if (groups == 1) {
        const int rc = norm_stream(2, dtype, x, da, dxo, mean, invstd, scale, shift, sums, nullptr, nullptr, rpg, c, act, slope,
                                   (cudaStream_t)stream);
        if (rc != VP_EUNSUPPORTED) return rc;
    }
    const int tpr = threads_per_row(c);
    const int ctiles = (c + tpr * VR - 1) / (tpr * VR);
    dim3 grid(ctiles, slab_blocks(rpg, ctiles, groups), (unsigned)groups);
    if (dtype == VP_F32)
        launch_k(bwd_reduce_kernel<float>, dim3(grid), dim3(NT), 0, (cudaStream_t)stream, (const float*)x, (const float*)da, mean, invstd, scale, shift, sums, (float*)dxo, groups, rpg, c, act, slope, tpr);
    else
        launch_k(bwd_reduce_kernel<bf16>, dim3(grid), dim3(NT), 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)da, mean, invstd, scale, shift, sums, (bf16*)dxo, groups, rpg, c, act, slope, tpr);
    VP_CHECK_LAUNCH("vp_norm_bwd_reduce");
    return VP_OK;
}

extern "C" int vp_norm_bwd_apply(const void* x, const void* da, const float* mean, const float* invstd,
                                 const float* scale, const float* shift, const double* sums, void* dx,
                                 float* dgamma, float* dbeta, int dtype, int64_t groups, int64_t rpg, int c,
                                 int act, float slope, void* stream) {
    VP_CHECK_ARG(x && da && mean && invstd && scale && shift && sums && dx && groups > 0 && rpg > 0 && c > 0,
                 "vp_norm_bwd_apply: bad arguments");
    VP_CHECK_ARG(groups <= 65535, "vp_norm_bwd_apply: too many groups");
    if (groups == 1) {
        const int rc = norm_stream(3, dtype, x, da, dx, mean, invstd, scale, shift, const_cast<double*>(sums), dgamma, dbeta, rpg, c, act,
                                   slope, (cudaStream_t)stream);
        if (rc != VP_EUNSUPPORTED) return rc;
    }
    dim3 grid((unsigned)((rpg + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), (c + CPB - 1) / CPB, (unsigned)groups);
    if (dtype == VP_F32)
        launch_k(bwd_apply_kernel<float>, dim3(grid), dim3(NT), 0, (cudaStream_t)stream, (const float*)x, (const float*)da, mean, invstd, scale, shift, sums, (float*)dx, dgamma, dbeta, groups, rpg, c, act, slope);
    else
        launch_k(bwd_apply_kernel<bf16>, dim3(grid), dim3(NT), 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)da, mean, invstd, scale, shift, sums, (bf16*)dx, dgamma, dbeta, groups, rpg, c, act, slope);
    VP_CHECK_LAUNCH("vp_norm_bwd_apply");
    return VP_OK;
}

extern "C" int vp_colsum(const void* x, float* out, double* scratch_c, int dtype, int64_t rows, int c, void* stream) {
    VP_CHECK_ARG(x && out && scratch_c && rows > 0 && c > 0, "vp_colsum: bad arguments");
    // scratch_c: double [2][c], zero on entry; reuse the statistics kernel (sum column)
    int rc = vp_norm_stats(x, scratch_c, dtype, 1, rows, c, stream);
    if (rc) return rc;
    launch_k(colsum_finish_kernel, dim3((c + 255) / 256), dim3(256), 0, (cudaStream_t)stream, scratch_c, out, c);
    VP_CHECK_LAUNCH("vp_colsum");
    return VP_OK;
}

// =====================================================================================================================
// BatchNorm over FEW rows and many channels in ONE launch per direction: the BatchNorm1d behind the fc layers
// (models/networks.py:66,89: [batch, 1024] and [batch, 8*8*C]).  A block owns 32 channels; its 32 row slices each walk
// the rows of those channels (second pass served by L2), so statistics + finalize + apply -- five launches and three
// memsets of the generic path -- become one kernel, and the same for the backward pair.
// =====================================================================================================================
namespace vp {
namespace {
constexpr int SM_CG = 8, SM_RS = 32, SM_V = 4;      // 8 channel groups x 4 channels, 32 row slices -> 256 threads, 32 channels per block

template <typename T>
__global__ void __launch_bounds__(256) bn_rows_fwd_kernel(const T* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* running_mean, float* running_var, float momentum, float eps, T* __restrict__ a,
                                                          float* mean, float* invstd, float* scale, float* shift, int rows, int C, int act,
                                                          float slope, long long* nbt) {
    pdl_sync();
    if (nbt && blockIdx.x == 0 && threadIdx.x == 0) *nbt += 1;
    typedef typename AccT<T>::type Acc;
    __shared__ double red[2][SM_RS][SM_CG * SM_V];
    __shared__ float par[2][SM_CG * SM_V];
    const int cg = threadIdx.x % SM_CG, rs = threadIdx.x / SM_CG;
    const int c0 = blockIdx.x * (SM_CG * SM_V) + cg * SM_V;
    const bool on = c0 < C;
    Acc s[SM_V] = {0, 0, 0, 0}, q[SM_V] = {0, 0, 0, 0};
    if (on)
        for (int r = rs; r < rows; r += SM_RS) {
            float v[SM_V];
            ldv4<T>(y + (int64_t)r * C + c0, v);
#pragma unroll
            for (int j = 0; j < SM_V; ++j) { s[j] += v[j]; q[j] += (Acc)v[j] * v[j]; }
        }
#pragma unroll
    for (int j = 0; j < SM_V; ++j) { red[0][rs][cg * SM_V + j] = (double)s[j]; red[1][rs][cg * SM_V + j] = (double)q[j]; }
    __syncthreads();
    if (threadIdx.x < SM_CG * SM_V) {
        const int c = blockIdx.x * (SM_CG * SM_V) + threadIdx.x;
        if (c < C) {
            double s1 = 0, s2 = 0;
#pragma unroll
            for (int i = 0; i < SM_RS; ++i) { s1 += red[0][i][threadIdx.x]; s2 += red[1][i][threadIdx.x]; }
            const double m = (double)rows, mu = s1 / m;
            double var = s2 / m - mu * mu;
            if (var < 0) var = 0;
            const float is = (float)(1.0 / sqrt(var + (double)eps));
            const float sc = (gamma ? gamma[c] : 1.f) * is;
            const float sh = (beta ? beta[c] : 0.f) - (float)mu * sc;
            mean[c] = (float)mu; invstd[c] = is; scale[c] = sc; shift[c] = sh;
            par[0][threadIdx.x] = sc; par[1][threadIdx.x] = sh;
            if (running_mean) {
                const double unb = var * m / (m > 1 ? m - 1 : 1);
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
            }
        }
    }
    __syncthreads();
    if (!on) return;
    float sc[SM_V], sh[SM_V];
#pragma unroll
    for (int j = 0; j < SM_V; ++j) { sc[j] = par[0][cg * SM_V + j]; sh[j] = par[1][cg * SM_V + j]; }
    for (int r = rs; r < rows; r += SM_RS) {
        float v[SM_V];
        ldv4<T>(y + (int64_t)r * C + c0, v);
#pragma unroll
        for (int j = 0; j < SM_V; ++j) v[j] = act_fwd(fmaf(v[j], sc[j], sh[j]), act, slope);
        stv4<T>(a + (int64_t)r * C + c0, v);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_rows_bwd_kernel(const T* __restrict__ y, const T* __restrict__ da, const float* __restrict__ mean,
                                                          const float* __restrict__ invstd, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, T* __restrict__ dy, float* dgamma, float* dbeta, int rows,
                                                          int C, int act, float slope) {
    pdl_sync();
    typedef typename AccT<T>::type Acc;
    __shared__ double red[2][SM_RS][SM_CG * SM_V];
    __shared__ float par[2][SM_CG * SM_V];
    const int cg = threadIdx.x % SM_CG, rs = threadIdx.x / SM_CG;
    const int c0 = blockIdx.x * (SM_CG * SM_V) + cg * SM_V;
    const bool on = c0 < C;
    float sc[SM_V] = {1, 1, 1, 1}, sh[SM_V] = {0, 0, 0, 0}, mu[SM_V] = {0, 0, 0, 0}, is[SM_V] = {0, 0, 0, 0};
    if (on) {
#pragma unroll
        for (int j = 0; j < SM_V; ++j) { sc[j] = scale[c0 + j]; sh[j] = shift[c0 + j]; mu[j] = mean[c0 + j]; is[j] = invstd[c0 + j]; }
    }
    Acc s[SM_V] = {0, 0, 0, 0}, q[SM_V] = {0, 0, 0, 0};
    if (on)
        for (int r = rs; r < rows; r += SM_RS) {
            float v[SM_V], d[SM_V];
            ldv4<T>(y + (int64_t)r * C + c0, v);
            ldv4<T>(da + (int64_t)r * C + c0, d);
#pragma unroll
            for (int j = 0; j < SM_V; ++j) {
                const float dd = d[j] * act_grad(fmaf(v[j], sc[j], sh[j]), act, slope);
                s[j] += dd;
                q[j] += (Acc)dd * ((v[j] - mu[j]) * is[j]);
            }
        }
#pragma unroll
    for (int j = 0; j < SM_V; ++j) { red[0][rs][cg * SM_V + j] = (double)s[j]; red[1][rs][cg * SM_V + j] = (double)q[j]; }
    __syncthreads();
    if (threadIdx.x < SM_CG * SM_V) {
        const int c = blockIdx.x * (SM_CG * SM_V) + threadIdx.x;
        if (c < C) {
            double s1 = 0, s2 = 0;
#pragma unroll
            for (int i = 0; i < SM_RS; ++i) { s1 += red[0][i][threadIdx.x]; s2 += red[1][i][threadIdx.x]; }
            if (dbeta) dbeta[c] = (float)s1;
            if (dgamma) dgamma[c] = (float)s2;
            par[0][threadIdx.x] = (float)(s1 / rows); par[1][threadIdx.x] = (float)(s2 / rows);
        }
    }
    __syncthreads();
    if (!on) return;
    float m1[SM_V], m2[SM_V];
#pragma unroll
    for (int j = 0; j < SM_V; ++j) { m1[j] = par[0][cg * SM_V + j]; m2[j] = par[1][cg * SM_V + j]; }
    for (int r = rs; r < rows; r += SM_RS) {
        float v[SM_V], d[SM_V];
        ldv4<T>(y + (int64_t)r * C + c0, v);
        ldv4<T>(da + (int64_t)r * C + c0, d);
#pragma unroll
        for (int j = 0; j < SM_V; ++j) {
            const float dd = d[j] * act_grad(fmaf(v[j], sc[j], sh[j]), act, slope);
            d[j] = sc[j] * (dd - m1[j] - (v[j] - mu[j]) * is[j] * m2[j]);
        }
        stv4<T>(dy + (int64_t)r * C + c0, d);
    }
}
}  // namespace
}  // namespace vp

/* Train-mode BatchNorm (+ activation) over x [rows, c] with few rows (<= 8192) in ONE launch: statistics, finalize
 * (mean / invstd / scale / shift as vp_norm_finalize, running statistics blended in place) and a = act(x*scale + shift).
 * c must be a multiple of 4 and the tensors 16-byte aligned.  The BatchNorm1d of models/networks.py:66,89. */
extern "C" int vp_bn_rows_fwd(const void* x, const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum,
                              float eps, void* a, float* mean, float* invstd, float* scale, float* shift, int dtype, int64_t rows, int c,
                              int act, float slope, int64_t* num_batches_tracked, void* stream) {
    VP_CHECK_ARG(x && a && mean && invstd && scale && shift && rows > 0 && rows <= 8192 && c > 0 && c % 4 == 0, "vp_bn_rows_fwd: bad arguments");
    VP_CHECK_ARG((((uintptr_t)x | (uintptr_t)a) & 15) == 0, "vp_bn_rows_fwd: 16-byte alignment");
    const unsigned grid = (unsigned)((c + SM_CG * SM_V - 1) / (SM_CG * SM_V));
    if (dtype == VP_F32)
        launch_k(bn_rows_fwd_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float*)x, gamma, beta, running_mean, running_var, momentum, eps,
                                                                         (float*)a, mean, invstd, scale, shift, (int)rows, c, act, slope, (long long*)num_batches_tracked);
    else
        launch_k(bn_rows_fwd_kernel<bf16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, gamma, beta, running_mean, running_var, momentum, eps,
                                                                        (bf16*)a, mean, invstd, scale, shift, (int)rows, c, act, slope, (long long*)num_batches_tracked);
    VP_CHECK_LAUNCH("vp_bn_rows_fwd");
    return VP_OK;
}

/* Backward of the above in one launch: dx, dgamma, dbeta (nullable) from x, da and the saved statistics. */
extern "C" int vp_bn_rows_bwd(const void* x, const void* da, const float* mean, const float* invstd, const float* scale, const float* shift,
                              void* dx, float* dgamma, float* dbeta, int dtype, int64_t rows, int c, int act, float slope, void* stream) {
    VP_CHECK_ARG(x && da && dx && mean && invstd && scale && shift && rows > 0 && rows <= 8192 && c > 0 && c % 4 == 0,
                 "vp_bn_rows_bwd: bad arguments");
    VP_CHECK_ARG((((uintptr_t)x | (uintptr_t)da | (uintptr_t)dx) & 15) == 0, "vp_bn_rows_bwd: 16-byte alignment");
    const unsigned grid = (unsigned)((c + SM_CG * SM_V - 1) / (SM_CG * SM_V));
    if (dtype == VP_F32)
        launch_k(bn_rows_bwd_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float*)x, (const float*)da, mean, invstd, scale, shift, (float*)dx,
                                                                         dgamma, dbeta, (int)rows, c, act, slope);
    else
        launch_k(bn_rows_bwd_kernel<bf16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)da, mean, invstd, scale, shift, (bf16*)dx,
                                                                        dgamma, dbeta, (int)rows, c, act, slope);
    VP_CHECK_LAUNCH("vp_bn_rows_bwd");
    return VP_OK;
}
