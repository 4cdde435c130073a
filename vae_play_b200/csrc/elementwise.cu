// Reparameterisation + KL (Philox eps, bit-compatible with Tensor.normal_() on CUDA), reconstruction
// losses, layout conversion and weight (un)packing.  All single-pass, coalesced, warp-shuffle reductions.
#include "common.cuh"

namespace vp {
namespace {

// ---- Philox4x32-10 -------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// curand_normal.h:70-87 (_curand_box_muller), device branch: logf + __sincosf
__device__ __forceinline__ float2 box_muller(uint32_t x, uint32_t y) {
    const float u = x * 2.3283064e-10f + (2.3283064e-10f / 2);
    const float v = y * (2.3283064e-10f * 6.2831855f) + ((2.3283064e-10f * 6.2831855f) / 2);
    const float s = sqrtf(-2.0f * logf(u));
    float2 r;
    __sincosf(v, &r.x, &r.y);
    r.x *= s;
    r.y *= s;
    return r;
}

// Element li of an n-element Tensor.normal_() draw (ATen DistributionTemplates.h:65-91):
//   thread = li % (grid*256), slot = li / (grid*256); counter = offset/4 + slot/4; component = slot%4
__device__ __forceinline__ float aten_normal_element(int64_t li, int64_t nthreads, uint64_t seed, uint64_t offset) {
    const uint64_t thread = (uint64_t)(li % nthreads);
    const uint64_t slot = (uint64_t)(li / nthreads);
    const uint64_t cnt = offset / 4 + slot / 4;
    const uint4 ctr = make_uint4((uint32_t)cnt, (uint32_t)(cnt >> 32), (uint32_t)thread, (uint32_t)(thread >> 32));
    const uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const int comp = (int)(slot & 3);
    const float2 n = (comp < 2) ? box_muller(r.x, r.y) : box_muller(r.z, r.w);
    const float v = (comp & 1) ? n.y : n.x;
    return v * 1.0f + 0.0f;  // transformation::normal(rand, mean=0, std=1)
}

__host__ __device__ inline int64_t aten_threads(int64_t n, int num_sms) {
    int64_t grid = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms * 8;  // maxThreadsPerMultiProcessor(2048) / 256
    if (grid > cap) grid = cap;
    return grid * 256;
}

__global__ void philox_normal_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t offset,
                                     const uint64_t* __restrict__ offset_dev, int64_t nthreads) {
    pdl_sync();
    if (offset_dev) offset += *offset_dev;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = aten_normal_element(i, nthreads, seed, offset);
}

__global__ void philox_advance_kernel(uint64_t* o, uint64_t inc) {
    pdl_sync(); *o += inc; }

// one warp per row: z = eps*exp(.5*lv)+mu ; kl_row = -0.5*sum(-exp(lv) - mu^2 + lv + 1)
template <typename TZ>
__global__ void __launch_bounds__(256) reparam_kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                             int64_t ld, const float* __restrict__ eps_in,
                                                             uint64_t seed, uint64_t offset,
                                                             const uint64_t* __restrict__ offset_dev, int64_t nthreads,
                                                             TZ* __restrict__ z, float* __restrict__ eps_out,
                                                             float* __restrict__ kl, int64_t rows, int zdim) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    if (offset_dev) offset += *offset_dev;
    float acc = 0.f;
    for (int j = lane; j < zdim; j += 32) {
        const float m = mu[row * ld + j], l = lv[row * ld + j];
        const int64_t li = row * zdim + j;
        const float e = eps_in ? eps_in[li] : aten_normal_element(li, nthreads, seed, offset);
        const float sd = expf(0.5f * l);
        Cvt<TZ>::st(z + li, fmaf(e, sd, m));
        if (eps_out) eps_out[li] = e;
        acc += -expf(l) - m * m + l + 1.f;
    }
    acc = warp_sum(acc);
    if (lane == 0 && kl) kl[row] = -0.5f * acc;
}

template <typename TZ, typename TO>
__global__ void __launch_bounds__(256) reparam_kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                             int64_t ld, const float* __restrict__ eps,
                                                             const TZ* __restrict__ dz, const float* __restrict__ dkl,
                                                             TO* __restrict__ dmu, TO* __restrict__ dlv,
                                                             int64_t ldo, int64_t rows, int zdim) {
    pdl_sync();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * zdim) return;
    const int64_t row = i / zdim;
    const int j = (int)(i - row * zdim);
    const float m = mu[row * ld + j], l = lv[row * ld + j];
    const float g = dz ? Cvt<TZ>::ld(dz + i) : 0.f;
    const float k = dkl ? dkl[row] : 0.f;
    const float e = eps[i];
    Cvt<TO>::st(dmu + row * ldo + j, g + k * m);
    Cvt<TO>::st(dlv + row * ldo + j, g * e * 0.5f * expf(0.5f * l) + k * 0.5f * (expf(l) - 1.f));
}

// ---- reconstruction losses ------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_double(double v) {
    __shared__ double sh[32];
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    __syncthreads();
    return v;
}

__global__ void __launch_bounds__(256) recon_fwd_kernel(const float* __restrict__ x, const float* __restrict__ xt,
                                                        int64_t n, int kind, double* acc, unsigned int* counter,
                                                        float* loss) {
    pdl_sync();
    float s = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(xt)) & 15) == 0) {
            const float4 a = *reinterpret_cast<const float4*>(x + i);
            const float4 b = *reinterpret_cast<const float4*>(xt + i);
            const float d0 = b.x - a.x, d1 = b.y - a.y, d2 = b.z - a.z, d3 = b.w - a.w;
            s += kind == 0 ? (d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) : (fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3));
        } else {
            for (int64_t k = i; k < n && k < i + 4; ++k) {
                const float d = xt[k] - x[k];
                s += kind == 0 ? d * d : fabsf(d);
            }
        }
    }
    const double t = block_sum_double((double)s);
    if (threadIdx.x == 0) {
        atomicAdd(acc, t);
        __threadfence();
        const unsigned int done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {
            const double total = atomicAdd(acc, 0.0);
            *loss = (float)(total / (double)n);
            *acc = 0.0;  // self-cleaning: ready for the next call / graph replay
            *counter = 0;
        }
    }
}

__global__ void __launch_bounds__(256) recon_bwd_kernel(const float* __restrict__ x, const float* __restrict__ xt,
                                                        int64_t n, int kind, const float* __restrict__ gscale,
                                                        float* __restrict__ dxt) {
    pdl_sync();
    const float g = (gscale ? *gscale : 1.f) / (float)n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float d = xt[i] - x[i];
        dxt[i] = kind == 0 ? 2.f * d * g : (d > 0.f ? g : (d < 0.f ? -g : 0.f));
    }
}

// acc[row] = {sum bce, sum p*t, sum p, sum t}; block handles a slab of one sample
__global__ void __launch_bounds__(256) bce_dice_fwd_kernel(const float* __restrict__ z, const float* __restrict__ t,
                                                           int64_t rows, int64_t per, float wbce, double* acc,
                                                           unsigned int* counter, float* loss) {
    pdl_sync();
    const int64_t row = blockIdx.y;
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
        const float zz = z[row * per + i], tt = t[row * per + i];
        const float p = 1.f / (1.f + expf(-zz));
        s0 += fmaxf(zz, 0.f) - zz * tt + log1pf(expf(-fabsf(zz)));
        s1 += p * tt;
        s2 += p;
        s3 += tt;
    }
    const double a0 = block_sum_double(s0), a1 = block_sum_double(s1), a2 = block_sum_double(s2), a3 = block_sum_double(s3);
    if (threadIdx.x == 0) {
        atomicAdd(acc + row * 4 + 0, a0);
        atomicAdd(acc + row * 4 + 1, a1);
        atomicAdd(acc + row * 4 + 2, a2);
        atomicAdd(acc + row * 4 + 3, a3);
        __threadfence();
        const unsigned int done = atomicAdd(counter, 1u);
        if (done == gridDim.x * gridDim.y - 1) {
            double bce = 0, dice = 0;
            for (int64_t r = 0; r < rows; ++r) {
                bce += atomicAdd(acc + r * 4 + 0, 0.0);
                const double inter = atomicAdd(acc + r * 4 + 1, 0.0), sp = atomicAdd(acc + r * 4 + 2, 0.0),
                             st = atomicAdd(acc + r * 4 + 3, 0.0);
                dice += (2.0 * inter + 1.0) / (sp + st + 1.0);
            }
            *loss = (float)(wbce * bce / ((double)rows * (double)per) + (1.0 - dice / (double)rows));
            *counter = 0;
        }
    }
}

__global__ void __launch_bounds__(256) bce_dice_bwd_kernel(const float* __restrict__ z, const float* __restrict__ t,
                                                           int64_t rows, int64_t per, float wbce,
                                                           const double* __restrict__ acc, const float* __restrict__ gscale,
                                                           float* __restrict__ dz) {
    pdl_sync();
    const int64_t row = blockIdx.y;
    const float g = gscale ? *gscale : 1.f;
    const double inter = acc[row * 4 + 1], den = acc[row * 4 + 2] + acc[row * 4 + 3] + 1.0;
    const double num = 2.0 * inter + 1.0;
    const float kb = wbce / ((float)rows * (float)per);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
        const float zz = z[row * per + i], tt = t[row * per + i];
        const float p = 1.f / (1.f + expf(-zz));
        const float ddice = (float)(-(2.0 * tt * den - num) / (den * den) / (double)rows);
        dz[row * per + i] = g * (kb * (p - tt) + ddice * p * (1.f - p));
    }
}

// ---- layout / packing -------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int n, int c,
                                                           int64_t hw) {
    pdl_sync();
    // y[(n*hw + p)*c + ch] = x[(n*c + ch)*hw + p]; tile-transpose through shared memory
    __shared__ float tile[32][33];
    const int img = blockIdx.z;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int ch = c0 + i;
        const int64_t p = p0 + tx;
        tile[i][tx] = (ch < c && p < hw) ? x[((int64_t)img * c + ch) * hw + p] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int64_t p = p0 + i;
        const int ch = c0 + tx;
        if (ch < c && p < hw) Cvt<T>::st(y + ((int64_t)img * hw + p) * c + ch, tile[tx][i]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int n, int c,
                                                           int64_t hw) {
    pdl_sync();
    __shared__ float tile[32][33];
    const int img = blockIdx.z;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int64_t p = p0 + i;
        const int ch = c0 + tx;
        tile[i][tx] = (ch < c && p < hw) ? Cvt<T>::ld(x + ((int64_t)img * hw + p) * c + ch) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int ch = c0 + i;
        const int64_t p = p0 + tx;
        if (ch < c && p < hw) y[((int64_t)img * c + ch) * hw + p] = tile[tx][i];
    }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, int64_t n) {
    pdl_sync();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        Cvt<TD>::st(d + i, Cvt<TS>::ld(s + i));
}

// ---- zero fill (replaces cudaMemsetAsync inside the library) -------------------------------------------------------------------
__global__ void __launch_bounds__(256) zero_fill_kernel(uint32_t* __restrict__ p, int64_t n4, int64_t nwords) {
    pdl_sync();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) reinterpret_cast<uint4*>(p)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) p[i] = 0u;
}

// ---- channel padding: layers whose channel counts are not multiples of 64 run on the 64-multiple tcgen05 kernels -------------
// dst[r, j] = j < c ? src[r, j] : 0   (activations, channels-last rows)
template <typename T>
__global__ void __launch_bounds__(256) pad_channels_kernel(const T* __restrict__ src, int c, T* __restrict__ dst, int cp, int64_t rows) {
    pdl_sync();
    const int64_t total = rows * cp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cp;
        const int j = (int)(i - r * cp);
        Cvt<T>::st(dst + i, j < c ? Cvt<T>::ld(src + r * c + j) : 0.f);
    }
}
// the same for 2-byte elements with c % 8 == 0 and cp % 8 == 0 (32 -> 64 channels): one 16-byte chunk per thread, no divisions by
// non-powers of two in the common case
__global__ void __launch_bounds__(256) pad_channels16_kernel(const uint4* __restrict__ src, int c8, uint4* __restrict__ dst, int cp8, int64_t rows) {
    pdl_sync();
    const int64_t total = rows * cp8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cp8;
        const int j = (int)(i - r * cp8);
        dst[i] = j < c8 ? src[r * c8 + j] : make_uint4(0u, 0u, 0u, 0u);
    }
}
// bf16, any c, cp % 8 == 0 (RGB 3 -> 64): one 16-byte OUTPUT chunk per thread, its up-to-8 source elements gathered one by one
__global__ void __launch_bounds__(256) pad_channels_out16_kernel(const bf16* __restrict__ src, int c, uint4* __restrict__ dst, int cp8, int64_t rows) {
    pdl_sync();
    const int64_t total = rows * cp8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cp8;
        const int j0 = (int)(i - r * cp8) * 8;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (j0 < c) {
            const bf16* sp = src + r * c + j0;
            unsigned short e[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = (j0 + k < c) ? __bfloat16_as_ushort(sp[k]) : (unsigned short)0;
            o = make_uint4((uint32_t)e[0] | ((uint32_t)e[1] << 16), (uint32_t)e[2] | ((uint32_t)e[3] << 16), (uint32_t)e[4] | ((uint32_t)e[5] << 16),
                           (uint32_t)e[6] | ((uint32_t)e[7] << 16));
        }
        dst[i] = o;
    }
}
// weight [d0][d1][taps] with element strides (s0, s1, st)  ->  bf16 [d0p][taps][d1p] (channels-last element order), zero padded
__global__ void __launch_bounds__(256) pad_weight_cl_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int d0, int d1, int taps, int64_t s0,
                                                            int64_t s1, int64_t st, int d0p, int d1p) {
    pdl_sync();
    const int64_t total = (int64_t)d0p * taps * d1p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int i1 = (int)(i % d1p);
        const int t = (int)((i / d1p) % taps);
        const int i0 = (int)(i / ((int64_t)d1p * taps));
        dst[i] = __float2bfloat16_rn((i0 < d0 && i1 < d1) ? w[i0 * s0 + i1 * s1 + t * st] : 0.f);
    }
}
// the inverse gather for the fp32 weight gradient: dw[i0*s0 + i1*s1 + t*st] = dwp[(i0*taps + t)*d1p + i1]
__global__ void __launch_bounds__(256) unpad_wgrad_cl_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int d0, int d1, int taps, int64_t s0,
                                                             int64_t s1, int64_t st, int d1p) {
    pdl_sync();
    const int64_t total = (int64_t)d0 * taps * d1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int i1 = (int)(i % d1);
        const int t = (int)((i / d1) % taps);
        const int i0 = (int)(i / ((int64_t)d1 * taps));
        dw[i0 * s0 + i1 * s1 + t * st] = dwp[((int64_t)i0 * taps + t) * d1p + i1];
    }
}

// dst (bf16) = src (fp32); src <- 0 when zero_src: the gradient bucket goes onto the wire in bf16 and is cleared for the next step
// in the same pass (n a multiple of 4, 16-byte aligned: bucket slots are)
__global__ void __launch_bounds__(256) pack_grads_kernel(float* __restrict__ src, bf16* __restrict__ dst, int64_t n4, int zero_src) {
    pdl_sync();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        reinterpret_cast<uint2*>(dst)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        if (zero_src) reinterpret_cast<float4*>(src)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__global__ void __launch_bounds__(256) axpy_kernel(float alpha, const float* __restrict__ a, float* __restrict__ sum, int64_t n) {
    pdl_sync();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        sum[i] = fmaf(alpha, a[i], sum[i]);
}

// Wp[t][n][k] = w[n*sn + k*sk + t]; one thread per (n,k) reads its taps (contiguous in torch layout)
template <typename T>
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, T* __restrict__ wp, int taps, int N,
                                                          int K, int64_t sn, int64_t sk, int64_t st) {
    pdl_sync();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * K) return;
    const int n = (int)(i / K), k = (int)(i - (int64_t)n * K);
    const float* src = w + n * sn + k * sk;
    for (int t = 0; t < taps; ++t) Cvt<T>::st(wp + ((int64_t)t * N + n) * K + k, src[t * st]);
}

// Tiled variants for the torch conv layouts (taps contiguous in the source, stride_t == 1): a 32(k) x T tile goes
// through shared memory so that both the fp32 side (contiguous along t) and the packed side (contiguous along k)
// are accessed in full segments.  grid = (K/32, N).
template <typename T>
__global__ void __launch_bounds__(256) pack_weight_tiled_kernel(const float* __restrict__ w, T* __restrict__ wp, int taps, int N,
                                                                int K, int64_t sn, int64_t sk) {
    pdl_sync();
    extern __shared__ float tile[];   // [32][taps+1]
    const int n = blockIdx.y, k0 = blockIdx.x * 32;
    const int ld = taps + 1;
    for (int i = threadIdx.x; i < 32 * taps; i += blockDim.x) {
        const int kk = i / taps, tt = i - kk * taps;
        tile[kk * ld + tt] = (k0 + kk < K) ? w[n * sn + (int64_t)(k0 + kk) * sk + tt] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * taps; i += blockDim.x) {
        const int tt = i >> 5, kk = i & 31;
        if (k0 + kk < K) Cvt<T>::st(wp + ((int64_t)tt * N + n) * K + k0 + kk, tile[kk * ld + tt]);
    }
}

__global__ void __launch_bounds__(256) unpack_wgrad_tiled_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int taps,
                                                                 int N, int K, int64_t sn, int64_t sk) {
    pdl_sync();
    extern __shared__ float tile[];
    const int n = blockIdx.y, k0 = blockIdx.x * 32;
    const int ld = taps + 1;
    for (int i = threadIdx.x; i < 32 * taps; i += blockDim.x) {
        const int tt = i >> 5, kk = i & 31;
        tile[kk * ld + tt] = (k0 + kk < K) ? dwp[((int64_t)tt * N + n) * K + k0 + kk] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * taps; i += blockDim.x) {
        const int kk = i / taps, tt = i - kk * taps;
        if (k0 + kk < K) dw[n * sn + (int64_t)(k0 + kk) * sk + tt] = tile[kk * ld + tt];
    }
}

__global__ void __launch_bounds__(256) unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int taps,
                                                           int N, int K, int64_t sn, int64_t sk, int64_t st) {
    pdl_sync();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * K) return;
    const int n = (int)(i / K), k = (int)(i - (int64_t)n * K);
    float* dst = dw + n * sn + k * sk;
    for (int t = 0; t < taps; ++t) dst[t * st] = dwp[((int64_t)t * N + n) * K + k];
}

__global__ void __launch_bounds__(256) sum_into_kernel(const float* __restrict__ v, int64_t n, float scale, float* acc) {
    pdl_sync();
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)v[i];
    s = block_sum_double(s);
    if (threadIdx.x == 0) *acc += scale * (float)s;
}

__global__ void __launch_bounds__(256) fill_from_kernel(const float* __restrict__ g, float scale, float* __restrict__ out, int64_t n) {
    pdl_sync();
    const float v = scale * (g ? *g : 1.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = v;
}

}  // namespace

int zero_async(void* ptr, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return VP_OK;
    if (((uintptr_t)ptr & 3) || (bytes & 3)) {      // never the case for the library's fp32 / double scratch; keep the exact semantics anyway
        return cudaMemsetAsync(ptr, 0, bytes, s) == cudaSuccess ? VP_OK : VP_ECUDA;
    }
    const int64_t nwords = (int64_t)(bytes / 4);
    const int64_t n4 = ((uintptr_t)ptr & 15) == 0 ? nwords / 4 : 0;
    int64_t blocks = (nwords / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    launch_k(zero_fill_kernel, dim3((unsigned)blocks), dim3(256), 0, s, (uint32_t*)ptr, n4, nwords);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("zero_async: %s", cudaGetErrorString(e)); return VP_ECUDA; }
    count_launch();
    return VP_OK;
}

namespace {

inline unsigned grid_for(int64_t n, int per_thread = 1) {
    int64_t b = (n + 256LL * per_thread - 1) / (256LL * per_thread);
    const int64_t cap = 148 * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace
}  // namespace vp

using namespace vp;

extern "C" int vp_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                                int num_sms, void* stream) {
    VP_CHECK_ARG(out && n >= 0 && num_sms > 0, "vp_philox_normal: bad arguments");
    if (n == 0) return VP_OK;
    launch_k(philox_normal_kernel, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, out, n, seed, offset, offset_dev, aten_threads(n, num_sms));
    VP_CHECK_LAUNCH("vp_philox_normal");
    return VP_OK;
}

extern "C" int vp_philox_advance(uint64_t* offset_dev, uint64_t inc, void* stream) {
    VP_CHECK_ARG(offset_dev, "vp_philox_advance: null");
    launch_k(philox_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, offset_dev, inc);
    VP_CHECK_LAUNCH("vp_philox_advance");
    return VP_OK;
}

extern "C" int vp_reparam_kl_fwd(const float* mu, const float* logvar, int64_t ld, const float* eps_in, uint64_t seed,
                                 uint64_t offset, const uint64_t* offset_dev, int num_sms, void* z, int z_dtype,
                                 float* eps_out, float* kl, int64_t rows, int zdim, void* stream) {
    VP_CHECK_ARG(mu && logvar && z && rows >= 0 && zdim > 0 && ld >= zdim && num_sms > 0, "vp_reparam_kl_fwd: bad arguments");
    if (rows == 0) return VP_OK;
    const int64_t nthreads = aten_threads(rows * zdim, num_sms);
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (z_dtype == VP_F32)
        launch_k(reparam_kl_fwd_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, mu, logvar, ld, eps_in, seed, offset, offset_dev, nthreads, (float*)z, eps_out, kl, rows, zdim);
    else
        launch_k(reparam_kl_fwd_kernel<bf16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, mu, logvar, ld, eps_in, seed, offset, offset_dev, nthreads, (bf16*)z, eps_out, kl, rows, zdim);
    VP_CHECK_LAUNCH("vp_reparam_kl_fwd");
    return VP_OK;
}

extern "C" int vp_reparam_kl_bwd(const float* mu, const float* logvar, int64_t ld, const float* eps, const void* dz,
                                 int dz_dtype, const float* dkl, void* dmu, void* dlogvar, int out_dtype, int64_t ld_out,
                                 int64_t rows, int zdim, void* stream) {
    VP_CHECK_ARG(mu && logvar && eps && dmu && dlogvar && rows >= 0 && zdim > 0, "vp_reparam_kl_bwd: bad arguments");
    if (rows == 0) return VP_OK;
    const unsigned grid = (unsigned)((rows * zdim + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (dz_dtype == VP_F32 && out_dtype == VP_F32)
        launch_k(reparam_kl_bwd_kernel<float, float>, dim3(grid), dim3(256), 0, s, mu, logvar, ld, eps, (const float*)dz, dkl, (float*)dmu, (float*)dlogvar, ld_out, rows, zdim);
    else if (dz_dtype == VP_F32)
        launch_k(reparam_kl_bwd_kernel<float, bf16>, dim3(grid), dim3(256), 0, s, mu, logvar, ld, eps, (const float*)dz, dkl, (bf16*)dmu, (bf16*)dlogvar, ld_out, rows, zdim);
    else if (out_dtype == VP_F32)
        launch_k(reparam_kl_bwd_kernel<bf16, float>, dim3(grid), dim3(256), 0, s, mu, logvar, ld, eps, (const bf16*)dz, dkl, (float*)dmu, (float*)dlogvar, ld_out, rows, zdim);
    else
        launch_k(reparam_kl_bwd_kernel<bf16, bf16>, dim3(grid), dim3(256), 0, s, mu, logvar, ld, eps, (const bf16*)dz, dkl, (bf16*)dmu, (bf16*)dlogvar, ld_out, rows, zdim);
    VP_CHECK_LAUNCH("vp_reparam_kl_bwd");
    return VP_OK;
}

extern "C" int vp_recon_loss_fwd(const float* x, const float* xt, int64_t n, int kind, double* loss_acc,
                                 unsigned int* counter, float* loss, void* stream) {
    VP_CHECK_ARG(x && xt && loss_acc && counter && loss && n > 0 && (kind == 0 || kind == 1), "vp_recon_loss_fwd: bad arguments");
    launch_k(recon_fwd_kernel, dim3(grid_for(n, 4)), dim3(256), 0, (cudaStream_t)stream, x, xt, n, kind, loss_acc, counter, loss);
    VP_CHECK_LAUNCH("vp_recon_loss_fwd");
    return VP_OK;
}

extern "C" int vp_recon_loss_bwd(const float* x, const float* xt, int64_t n, int kind, const float* gscale, float* dxt,
                                 void* stream) {
    VP_CHECK_ARG(x && xt && dxt && n > 0 && (kind == 0 || kind == 1), "vp_recon_loss_bwd: bad arguments");
    launch_k(recon_bwd_kernel, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, x, xt, n, kind, gscale, dxt);
    VP_CHECK_LAUNCH("vp_recon_loss_bwd");
    return VP_OK;
}

extern "C" int vp_bce_dice_fwd(const float* logits, const float* target, int64_t rows, int64_t per, float bce_weight,
                               double* acc, unsigned int* counter, float* loss, void* stream) {
    VP_CHECK_ARG(logits && target && acc && counter && loss && rows > 0 && rows <= 65535 && per > 0, "vp_bce_dice_fwd: bad arguments");
    dim3 grid((unsigned)((per + 1023) / 1024 < 64 ? (per + 1023) / 1024 : 64), (unsigned)rows);
    zero_async(acc, sizeof(double) * 4 * rows, (cudaStream_t)stream);
    launch_k(bce_dice_fwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, logits, target, rows, per, bce_weight, acc, counter, loss);
    VP_CHECK_LAUNCH("vp_bce_dice_fwd");
    return VP_OK;
}

extern "C" int vp_bce_dice_bwd(const float* logits, const float* target, int64_t rows, int64_t per, float bce_weight,
                               const double* acc, const float* gscale, float* dlogits, void* stream) {
    VP_CHECK_ARG(logits && target && acc && dlogits && rows > 0 && rows <= 65535 && per > 0, "vp_bce_dice_bwd: bad arguments");
    dim3 grid((unsigned)((per + 1023) / 1024 < 64 ? (per + 1023) / 1024 : 64), (unsigned)rows);
    launch_k(bce_dice_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, logits, target, rows, per, bce_weight, acc, gscale, dlogits);
    VP_CHECK_LAUNCH("vp_bce_dice_bwd");
    return VP_OK;
}

extern "C" int vp_nchw_to_nhwc(const float* x, void* y, int dtype, int n, int c, int h, int w, void* stream) {
    VP_CHECK_ARG(x && y && n > 0 && c > 0 && h > 0 && w > 0 && n <= 65535, "vp_nchw_to_nhwc: bad arguments");
    const int64_t hw = (int64_t)h * w;
    dim3 grid((unsigned)((hw + 31) / 32), (c + 31) / 32, n);
    if (dtype == VP_F32) launch_k(nchw_to_nhwc_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, (float*)y, n, c, hw);
    else launch_k(nchw_to_nhwc_kernel<bf16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, (bf16*)y, n, c, hw);
    VP_CHECK_LAUNCH("vp_nchw_to_nhwc");
    return VP_OK;
}

extern "C" int vp_nhwc_to_nchw(const void* x, float* y, int dtype, int n, int c, int h, int w, void* stream) {
    VP_CHECK_ARG(x && y && n > 0 && c > 0 && h > 0 && w > 0 && n <= 65535, "vp_nhwc_to_nchw: bad arguments");
    const int64_t hw = (int64_t)h * w;
    dim3 grid((unsigned)((hw + 31) / 32), (c + 31) / 32, n);
    if (dtype == VP_F32) launch_k(nhwc_to_nchw_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float*)x, y, n, c, hw);
    else launch_k(nhwc_to_nchw_kernel<bf16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, y, n, c, hw);
    VP_CHECK_LAUNCH("vp_nhwc_to_nchw");
    return VP_OK;
}

extern "C" int vp_cast(const void* src, int sd, void* dst, int dd, int64_t n, void* stream) {
    VP_CHECK_ARG(src && dst && n >= 0, "vp_cast: bad arguments");
    if (n == 0) return VP_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(n);
    if (sd == VP_F32 && dd == VP_BF16) launch_k(cast_kernel<float, bf16>, dim3(g), dim3(256), 0, s, (const float*)src, (bf16*)dst, n);
    else if (sd == VP_BF16 && dd == VP_F32) launch_k(cast_kernel<bf16, float>, dim3(g), dim3(256), 0, s, (const bf16*)src, (float*)dst, n);
    else if (sd == VP_F32 && dd == VP_F32) launch_k(cast_kernel<float, float>, dim3(g), dim3(256), 0, s, (const float*)src, (float*)dst, n);
    else launch_k(cast_kernel<bf16, bf16>, dim3(g), dim3(256), 0, s, (const bf16*)src, (bf16*)dst, n);
    VP_CHECK_LAUNCH("vp_cast");
    return VP_OK;
}

extern "C" int vp_pad_channels(const void* src, int c, void* dst, int cp, int64_t rows, int dtype, void* stream) {
    VP_CHECK_ARG(src && dst && c > 0 && cp >= c && rows >= 0, "vp_pad_channels: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_pad_channels: bad dtype %d", dtype);
    if (rows == 0) return VP_OK;
    if (dtype == VP_BF16 && c % 8 == 0 && cp % 8 == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
        launch_k(pad_channels16_kernel, dim3(grid_for(rows * (cp / 8))), dim3(256), 0, (cudaStream_t)stream, (const uint4*)src, c / 8, (uint4*)dst, cp / 8, rows);
        VP_CHECK_LAUNCH("vp_pad_channels");
        return VP_OK;
    }
    if (dtype == VP_BF16 && cp % 8 == 0 && ((uintptr_t)dst & 15) == 0) {
        launch_k(pad_channels_out16_kernel, dim3(grid_for(rows * (cp / 8))), dim3(256), 0, (cudaStream_t)stream, (const bf16*)src, c, (uint4*)dst, cp / 8, rows);
        VP_CHECK_LAUNCH("vp_pad_channels");
        return VP_OK;
    }
    if (dtype == VP_F32) launch_k(pad_channels_kernel<float>, dim3(grid_for(rows * cp)), dim3(256), 0, (cudaStream_t)stream, (const float*)src, c, (float*)dst, cp, rows);
    else launch_k(pad_channels_kernel<bf16>, dim3(grid_for(rows * cp)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)src, c, (bf16*)dst, cp, rows);
    VP_CHECK_LAUNCH("vp_pad_channels");
    return VP_OK;
}

extern "C" int vp_pad_weight_cl(const float* w, void* dst_bf16, int d0, int d1, int taps, int64_t s0, int64_t s1, int64_t st, int d0p, int d1p,
                                void* stream) {
    VP_CHECK_ARG(w && dst_bf16 && d0 > 0 && d1 > 0 && taps > 0 && d0p >= d0 && d1p >= d1, "vp_pad_weight_cl: bad arguments");
    launch_k(pad_weight_cl_kernel, dim3(grid_for((int64_t)d0p * taps * d1p)), dim3(256), 0, (cudaStream_t)stream, w, (bf16*)dst_bf16, d0, d1, taps, s0, s1,
             st, d0p, d1p);
    VP_CHECK_LAUNCH("vp_pad_weight_cl");
    return VP_OK;
}

extern "C" int vp_unpad_wgrad_cl(const float* dwp, float* dw, int d0, int d1, int taps, int64_t s0, int64_t s1, int64_t st, int d1p, void* stream) {
    VP_CHECK_ARG(dwp && dw && d0 > 0 && d1 > 0 && taps > 0 && d1p >= d1, "vp_unpad_wgrad_cl: bad arguments");
    launch_k(unpad_wgrad_cl_kernel, dim3(grid_for((int64_t)d0 * taps * d1)), dim3(256), 0, (cudaStream_t)stream, dwp, dw, d0, d1, taps, s0, s1, st, d1p);
    VP_CHECK_LAUNCH("vp_unpad_wgrad_cl");
    return VP_OK;
}

extern "C" int vp_pack_grads_bf16(float* src, void* dst_bf16, int64_t n, int zero_src, void* stream) {
    VP_CHECK_ARG(src && dst_bf16 && n >= 0 && (n & 3) == 0 && (((uintptr_t)src | (uintptr_t)dst_bf16) & 15) == 0,
                 "vp_pack_grads_bf16: n must be a multiple of 4 and the buffers 16-byte aligned");
    if (n == 0) return VP_OK;
    static bool carve_set = false;       // meant to run NEXT TO the TMA kernels of the backward pass: see vp_rmsprop_step_wire
    if (!carve_set) {
        cudaFuncSetAttribute(pack_grads_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carve_set = true;
    }
    launch_k(pack_grads_kernel, dim3(grid_for(n / 4)), dim3(256), 0, (cudaStream_t)stream, src, (bf16*)dst_bf16, n / 4, zero_src);
    VP_CHECK_LAUNCH("vp_pack_grads_bf16");
    return VP_OK;
}

extern "C" int vp_axpy(float alpha, const float* a, float* sum, int64_t n, void* stream) {
    VP_CHECK_ARG(a && sum && n >= 0, "vp_axpy: bad arguments");
    if (n == 0) return VP_OK;
    launch_k(axpy_kernel, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, alpha, a, sum, n);
    VP_CHECK_LAUNCH("vp_axpy");
    return VP_OK;
}

extern "C" int vp_pack_weight(const float* w, void* wp, int dtype, int taps, int n, int k, int64_t sn, int64_t sk,
                              int64_t st, void* stream) {
    VP_CHECK_ARG(w && wp && taps > 0 && n > 0 && k > 0, "vp_pack_weight: bad arguments");
    if (st == 1 && taps >= 4 && k >= 32 && n <= 65535) {
        dim3 grid((k + 31) / 32, n);
        const size_t smem = sizeof(float) * 32 * (taps + 1);
        if (dtype == VP_F32) launch_k(pack_weight_tiled_kernel<float>, dim3(grid), dim3(256), smem, (cudaStream_t)stream, w, (float*)wp, taps, n, k, sn, sk);
        else launch_k(pack_weight_tiled_kernel<bf16>, dim3(grid), dim3(256), smem, (cudaStream_t)stream, w, (bf16*)wp, taps, n, k, sn, sk);
        VP_CHECK_LAUNCH("vp_pack_weight(tiled)");
        return VP_OK;
    }
    const unsigned g = (unsigned)(((int64_t)n * k + 255) / 256);
    if (dtype == VP_F32) launch_k(pack_weight_kernel<float>, dim3(g), dim3(256), 0, (cudaStream_t)stream, w, (float*)wp, taps, n, k, sn, sk, st);
    else launch_k(pack_weight_kernel<bf16>, dim3(g), dim3(256), 0, (cudaStream_t)stream, w, (bf16*)wp, taps, n, k, sn, sk, st);
    VP_CHECK_LAUNCH("vp_pack_weight");
    return VP_OK;
}

extern "C" int vp_unpack_wgrad(const float* dwp, float* dw, int taps, int n, int k, int64_t sn, int64_t sk, int64_t st,
                               void* stream) {
    VP_CHECK_ARG(dwp && dw && taps > 0 && n > 0 && k > 0, "vp_unpack_wgrad: bad arguments");
    if (st == 1 && taps >= 4 && k >= 32 && n <= 65535) {
        dim3 grid((k + 31) / 32, n);
        const size_t smem = sizeof(float) * 32 * (taps + 1);
        launch_k(unpack_wgrad_tiled_kernel, dim3(grid), dim3(256), smem, (cudaStream_t)stream, dwp, dw, taps, n, k, sn, sk);
        VP_CHECK_LAUNCH("vp_unpack_wgrad(tiled)");
        return VP_OK;
    }
    const unsigned g = (unsigned)(((int64_t)n * k + 255) / 256);
    launch_k(unpack_wgrad_kernel, dim3(g), dim3(256), 0, (cudaStream_t)stream, dwp, dw, taps, n, k, sn, sk, st);
    VP_CHECK_LAUNCH("vp_unpack_wgrad");
    return VP_OK;
}

extern "C" int vp_sum_into(const float* v, int64_t n, float scale, float* acc, void* stream) {
    VP_CHECK_ARG(v && acc && n >= 0, "vp_sum_into: bad arguments");
    launch_k(sum_into_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, v, n, scale, acc);
    VP_CHECK_LAUNCH("vp_sum_into");
    return VP_OK;
}

extern "C" int vp_fill_from(const float* g, float scale, float* out, int64_t n, void* stream) {
    VP_CHECK_ARG(out && n >= 0, "vp_fill_from: bad arguments");
    if (n == 0) return VP_OK;
    launch_k(fill_from_kernel, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, g, scale, out, n);
    VP_CHECK_LAUNCH("vp_fill_from");
    return VP_OK;
}

// ---- multi-tensor RMSprop ---------------------------------------------------------------------------------------
namespace vp {
namespace {
constexpr int kOptMax = 64;
struct OptTable {
    float* p[kOptMax];
    const float* g[kOptMax];
    float* sq[kOptMax];
    bf16* sh[kOptMax];          // optional bf16 copy of the updated parameter (same element order), or null
    const bf16* wire[kOptMax];  // optional: the gradient is read from this bf16 buffer (data-parallel wire format) instead of g
    int64_t n[kOptMax];
    int blk0[kOptMax + 1];      // rmsprop: tensor i is served by the CTAs [blk0[i], blk0[i+1]) of a flat grid (no idle CTAs)
};
__global__ void __launch_bounds__(256) rmsprop_kernel(const __grid_constant__ OptTable t, int ntensors, float lr, float alpha, float eps, float wd,
                                                      int zero_g) {
    pdl_wait();     // no early trigger: a dependent grid starts only once this grid has COMPLETED -- kernels that follow may read
                    // the fp32 master weights ahead of their own wait (thin_tc.cu)
    int ti = 0;
    while (ti + 1 < ntensors && (int)blockIdx.x >= t.blk0[ti + 1]) ++ti;
    const int bid = blockIdx.x - t.blk0[ti], nblk = t.blk0[ti + 1] - t.blk0[ti];
    float* __restrict__ p = t.p[ti];
    bf16* __restrict__ sh = t.sh[ti];
    const float* __restrict__ g = t.g[ti];
    float* __restrict__ sq = t.sq[ti];
    const int64_t n = t.n[ti];
    const bf16* __restrict__ wire = t.wire[ti];
    const int64_t n4 = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)sq) & 15) == 0 && (((uintptr_t)sh | (uintptr_t)wire) & 7) == 0 ? n / 4 : 0;
    const int64_t stride = (int64_t)nblk * blockDim.x;
    for (int64_t i = (int64_t)bid * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pv = reinterpret_cast<float4*>(p)[i];
        float4 gv0;
        if (wire) {
            const uint2 w2 = reinterpret_cast<const uint2*>(wire)[i];
            gv0 = make_float4(__uint_as_float(w2.x << 16), __uint_as_float(w2.x & 0xffff0000u), __uint_as_float(w2.y << 16), __uint_as_float(w2.y & 0xffff0000u));
        } else {
            gv0 = reinterpret_cast<const float4*>(g)[i];
        }
        float4 sv = reinterpret_cast<float4*>(sq)[i];
        float pe[4] = {pv.x, pv.y, pv.z, pv.w}, ge[4] = {gv0.x, gv0.y, gv0.z, gv0.w}, se[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gg = ge[j] + wd * pe[j];
            se[j] = alpha * se[j] + (1.f - alpha) * gg * gg;
            pe[j] -= lr * (gg / (sqrtf(se[j]) + eps));
        }
        reinterpret_cast<float4*>(p)[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
        reinterpret_cast<float4*>(sq)[i] = make_float4(se[0], se[1], se[2], se[3]);
        if (zero_g) reinterpret_cast<float4*>(const_cast<float*>(g))[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sh) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(pe[0], pe[1]), h1 = __floats2bfloat162_rn(pe[2], pe[3]);
            reinterpret_cast<uint2*>(sh)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        }
    }
    for (int64_t i = n4 * 4 + (int64_t)bid * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gg = (wire ? __bfloat162float(wire[i]) : g[i]) + wd * p[i];
        const float s2 = alpha * sq[i] + (1.f - alpha) * gg * gg;
        sq[i] = s2;
        const float pn = p[i] - lr * (gg / (sqrtf(s2) + eps));
        p[i] = pn;
        if (zero_g) const_cast<float*>(g)[i] = 0.f;
        if (sh) sh[i] = __float2bfloat16_rn(pn);
    }
}
}  // namespace
}  // namespace vp

extern "C" int vp_rmsprop_step_shadow(void* const* params, void* const* grads, void* const* sq, void* const* shadows, const int64_t* numel,
                                      int count, float lr, float alpha, float eps, float weight_decay, int zero_grads, void* stream);

extern "C" int vp_rmsprop_step(void* const* params, const void* const* grads, void* const* sq, const int64_t* numel, int count, float lr,
                               float alpha, float eps, float weight_decay, void* stream) {
    return vp_rmsprop_step_shadow(params, const_cast<void* const*>(grads), sq, nullptr, numel, count, lr, alpha, eps, weight_decay, 0, stream);
}

extern "C" int vp_rmsprop_step_wire(void* const* params, void* const* grads, void* const* sq, void* const* shadows, const void* const* wire_grads,
                                    const int64_t* numel, int count, float lr, float alpha, float eps, float weight_decay, int zero_grads, void* stream);
extern "C" int vp_rmsprop_step_shadow(void* const* params, void* const* grads, void* const* sq, void* const* shadows, const int64_t* numel,
                                      int count, float lr, float alpha, float eps, float weight_decay, int zero_grads, void* stream) {
    return vp_rmsprop_step_wire(params, grads, sq, shadows, nullptr, numel, count, lr, alpha, eps, weight_decay, zero_grads, stream);
}

/* vp_rmsprop_step_shadow whose gradients are read from bf16 buffers (wire_grads[i] non-NULL): the data-parallel exchange in
 * bf16 (vp_pack_grads_bf16 -> all-reduce -> this), without a conversion pass back to fp32. */
extern "C" int vp_rmsprop_step_wire(void* const* params, void* const* grads, void* const* sq, void* const* shadows, const void* const* wire_grads,
                                    const int64_t* numel, int count, float lr, float alpha, float eps, float weight_decay, int zero_grads, void* stream) {
    VP_CHECK_ARG(params && grads && sq && numel && count >= 0, "vp_rmsprop_step: bad arguments");
    for (int base = 0; base < count; base += kOptMax) {
        OptTable t;
        const int m = count - base < kOptMax ? count - base : kOptMax;
        // CTAs in proportion to each tensor's size: ~4 float4 iterations per thread at least, at most 2 CTAs per SM per tensor
        const int cap = 2 * num_sms();
        int total = 0;
        for (int i = 0; i < m; ++i) {
            t.p[i] = (float*)params[base + i]; t.g[i] = (const float*)grads[base + i]; t.sq[i] = (float*)sq[base + i];
            t.n[i] = numel[base + i];
            t.sh[i] = shadows ? (bf16*)shadows[base + i] : nullptr;
            t.wire[i] = wire_grads ? (const bf16*)wire_grads[base + i] : nullptr;
            int64_t nb = (numel[base + i] / 4 + 1023) / 1024;
            nb = nb < 1 ? 1 : (nb > cap ? cap : nb);
            t.blk0[i] = total;
            total += (int)nb;
        }
        t.blk0[m] = total;
        // Ask for the maximum shared-memory carve-out although the kernel uses none: an SM whose L1 / shared split was set for a
        // kernel without shared memory cannot take CTAs of the TMA kernels (75-220 KB each) until it has drained, i.e. the
        // optimiser could never run NEXT TO the backward pass it is meant to overlap (measured: the first BatchNorm pass of the
        // encoder backward started only when this kernel had finished).
        static bool carve_set = false;
        if (!carve_set) {
            cudaFuncSetAttribute(rmsprop_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            carve_set = true;
        }
        launch_k(rmsprop_kernel, dim3((unsigned)total), dim3(256), 0, (cudaStream_t)stream, t, m, lr, alpha, eps, weight_decay, zero_grads);
        VP_CHECK_LAUNCH("vp_rmsprop_step");
    }
    return VP_OK;
}

// ---- Adam (torch.optim.Adam without amsgrad; train_BE.py:131, train_Style_GAN.py) as ONE multi-tensor kernel --------------------
namespace vp {
namespace {
struct AdamTable {
    float* p[kOptMax];
    const float* g[kOptMax];
    float* m[kOptMax];
    float* v[kOptMax];
    bf16* sh[kOptMax];
    int64_t n[kOptMax];
    int blk0[kOptMax + 1];      // tensor i is served by the CTAs [blk0[i], blk0[i+1]) of a flat grid (as rmsprop_kernel)
};
__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamTable t, int ntensors, float lr, float b1, float b2, float eps, float wd,
                                                   int64_t step, const unsigned long long* __restrict__ step_dev, int zero_g) {
    pdl_wait();     // no early trigger (see rmsprop_kernel)
    int ti = 0;
    while (ti + 1 < ntensors && (int)blockIdx.x >= t.blk0[ti + 1]) ++ti;
    const int bid = blockIdx.x - t.blk0[ti], nblk = t.blk0[ti + 1] - t.blk0[ti];
    float* __restrict__ p = t.p[ti];
    float* __restrict__ g = const_cast<float*>(t.g[ti]);
    float* __restrict__ m = t.m[ti];
    float* __restrict__ v = t.v[ti];
    bf16* __restrict__ sh = t.sh[ti];
    const int64_t n = t.n[ti];
    const double tt = (double)(step_dev ? (int64_t)*step_dev : step);
    const float bc1 = (float)(1.0 - pow((double)b1, tt)), bc2s = (float)sqrt(1.0 - pow((double)b2, tt));
    const float step_size = lr / bc1;
    const int64_t stride = (int64_t)nblk * blockDim.x;
    const int64_t n4 = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && ((uintptr_t)sh & 7) == 0) ? n / 4 : 0;
    for (int64_t i = (int64_t)bid * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 p4 = reinterpret_cast<float4*>(p)[i], g4 = reinterpret_cast<float4*>(g)[i];
        const float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
        float pe[4] = {p4.x, p4.y, p4.z, p4.w}, ge[4] = {g4.x, g4.y, g4.z, g4.w}, me[4] = {m4.x, m4.y, m4.z, m4.w}, ve[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gg = ge[j] + wd * pe[j];
            me[j] = b1 * me[j] + (1.f - b1) * gg;
            ve[j] = b2 * ve[j] + (1.f - b2) * gg * gg;
            pe[j] = pe[j] - step_size * (me[j] / (sqrtf(ve[j]) / bc2s + eps));
        }
        reinterpret_cast<float4*>(m)[i] = make_float4(me[0], me[1], me[2], me[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
        reinterpret_cast<float4*>(p)[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
        if (zero_g) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sh) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(pe[0], pe[1]), h1 = __floats2bfloat162_rn(pe[2], pe[3]);
            reinterpret_cast<uint2*>(sh)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        }
    }
    for (int64_t i = n4 * 4 + (int64_t)bid * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float pv = p[i];
        const float gg = g[i] + wd * pv;
        const float mn = b1 * m[i] + (1.f - b1) * gg;
        const float vn = b2 * v[i] + (1.f - b2) * gg * gg;
        m[i] = mn; v[i] = vn;
        const float pn = pv - step_size * (mn / (sqrtf(vn) / bc2s + eps));
        p[i] = pn;
        if (zero_g) g[i] = 0.f;
        if (sh) sh[i] = __float2bfloat16_rn(pn);
    }
}
}  // namespace
}  // namespace vp

/* torch.optim.Adam (no amsgrad) over fp32 masters: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 * p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps), g <- g + wd p first.  t = `step` (1-based), or *step_dev when
 * non-NULL (a device counter the caller advances: CUDA-graph replay).  shadows / zero_grads as vp_rmsprop_step_shadow. */
extern "C" int vp_adam_step(void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq, void* const* shadows,
                            const int64_t* numel, int count, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                            const uint64_t* step_dev, int zero_grads, void* stream) {
    VP_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && numel && count >= 0 && (step >= 1 || step_dev), "vp_adam_step: bad arguments");
    for (int base = 0; base < count; base += kOptMax) {
        AdamTable t;
        const int mcount = count - base < kOptMax ? count - base : kOptMax;
        const int cap = 2 * num_sms();
        int total = 0;
        for (int i = 0; i < mcount; ++i) {
            t.p[i] = (float*)params[base + i]; t.g[i] = (const float*)grads[base + i];
            t.m[i] = (float*)exp_avg[base + i]; t.v[i] = (float*)exp_avg_sq[base + i];
            t.sh[i] = shadows ? (bf16*)shadows[base + i] : nullptr;
            t.n[i] = numel[base + i];
            int64_t nb = (numel[base + i] / 4 + 1023) / 1024;
            nb = nb < 1 ? 1 : (nb > cap ? cap : nb);
            t.blk0[i] = total;
            total += (int)nb;
        }
        t.blk0[mcount] = total;
        launch_k(adam_kernel, dim3((unsigned)total), dim3(256), 0, (cudaStream_t)stream, t, mcount, lr, beta1, beta2, eps, weight_decay, step,
                 (const unsigned long long*)step_dev, zero_grads);
        VP_CHECK_LAUNCH("vp_adam_step");
    }
    return VP_OK;
}

// ---- batched transpose (same dtype): dst[b][c][r] = src[b][r][c] ---------------------------------------------------------
// The NCHW-flatten Linear layers (models/networks.py:65,74-75 and :88,110) see the 8x8 map in (c, y, x) order while the
// activations are channels-last: a [B][64][C] <-> [B][C][64] transpose on either side lets them run as plain Linear layers
// on the module's own weight.
namespace vp {
namespace {
template <typename T>
__global__ void __launch_bounds__(256) transpose_bt_kernel(const T* __restrict__ src, T* __restrict__ dst, int rows, int cols) {
    pdl_sync();
    __shared__ T tile[32][33];
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const T* s = src + (int64_t)b * rows * cols;
    T* d = dst + (int64_t)b * rows * cols;
    for (int i = ty; i < 32; i += 8)
        if (r0 + i < rows && c0 + tx < cols) tile[i][tx] = s[(int64_t)(r0 + i) * cols + c0 + tx];
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < cols && r0 + tx < rows) d[(int64_t)(c0 + i) * rows + r0 + tx] = tile[tx][i];
}

// bf16, rows and cols multiples of 64: 64 x 64 tiles, every global access is a 4-byte pair, the 2 x 2 blocks are
// transposed in registers on the way out (4x fewer, 2x wider memory instructions than the generic kernel)
__global__ void __launch_bounds__(256) transpose_bt64_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int rows, int cols) {
    pdl_sync();
    __shared__ uint32_t tile[64][33];
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint32_t* s = src + ((int64_t)b * rows * cols >> 1);
    uint32_t* d = dst + ((int64_t)b * rows * cols >> 1);
    for (int i = ty; i < 64; i += 8) tile[i][tx] = s[((int64_t)(r0 + i) * cols + c0 >> 1) + tx];     // row r0+i, column pair tx
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        // output rows (= source columns) c0 + 2i, c0 + 2i + 1; output column pair tx = source rows r0 + 2tx, r0 + 2tx + 1
        const uint32_t a = tile[2 * tx][i], bb = tile[2 * tx + 1][i];
        const uint32_t lo = (a & 0xffffu) | (bb << 16), hi = (a >> 16) | (bb & 0xffff0000u);
        d[((int64_t)(c0 + 2 * i) * rows + r0 >> 1) + tx] = lo;
        d[((int64_t)(c0 + 2 * i + 1) * rows + r0 >> 1) + tx] = hi;
    }
}
}  // namespace
}  // namespace vp

extern "C" int vp_transpose_bt(const void* src, void* dst, int dtype, int batch, int rows, int cols, void* stream) {
    VP_CHECK_ARG(src && dst && batch > 0 && rows > 0 && cols > 0 && batch <= 65535, "vp_transpose_bt: bad arguments");
    if (dtype == VP_BF16 && rows % 64 == 0 && cols % 64 == 0 && (((uintptr_t)src | (uintptr_t)dst) & 3) == 0) {
        launch_k(transpose_bt64_kernel, dim3(cols / 64, rows / 64, batch), dim3(256), 0, (cudaStream_t)stream, (const uint32_t*)src, (uint32_t*)dst, rows, cols);
        VP_CHECK_LAUNCH("vp_transpose_bt");
        return VP_OK;
    }
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
    if (dtype == VP_F32) launch_k(transpose_bt_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float*)src, (float*)dst, rows, cols);
    else launch_k(transpose_bt_kernel<bf16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16*)src, (bf16*)dst, rows, cols);
    VP_CHECK_LAUNCH("vp_transpose_bt");
    return VP_OK;
}
