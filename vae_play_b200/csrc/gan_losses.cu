// The remaining terms of VaeGan.loss (reference models/networks.py:264-281) and of the train.py step (:62-67) as single
// coalesced warp-shuffle kernels on fp32 tensors: feature MSE between discriminator layers, -log(s*p + c) on the sigmoid
// scores, smooth-L1 on the auxiliary parameters and the per-sample KL.  All are a few KB..MB: latency-bound, one launch each.
#include "common.cuh"

namespace vp {
namespace {

// out[r] += 0.5 * sum_j (a[r,j] - b[r,j])^2 over the column slice of this CTA (grid = splits x rows; out zeroed by the caller: a
// batch of 64 rows x 65 536 features must not be left to 64 warps)                                                     networks.py:273
__global__ void __launch_bounds__(256) feat_mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t rows,
                                                           int64_t cols) {
    pdl_sync();
    __shared__ float part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = blockIdx.y;
    const int64_t chunk = ((cols + gridDim.x - 1) / gridDim.x + 3) & ~(int64_t)3;
    const int64_t c0 = (int64_t)blockIdx.x * chunk, c1 = c0 + chunk < cols ? c0 + chunk : cols;
    const float* pa = a + r * cols;
    const float* pb = b + r * cols;
    float s = 0.f;
    int64_t j0 = c0;
    if (c0 < c1 && (((uintptr_t)(pa + c0) | (uintptr_t)(pb + c0)) & 15) == 0) {
        const int64_t n4 = (c1 - c0) / 4;
        const float4* a4 = reinterpret_cast<const float4*>(pa + c0);
        const float4* b4 = reinterpret_cast<const float4*>(pb + c0);
#pragma unroll 4
        for (int64_t j = threadIdx.x; j < n4; j += 256) {
            const float4 x = a4[j], y = b4[j];
            const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
            s = fmaf(d0, d0, s); s = fmaf(d1, d1, s); s = fmaf(d2, d2, s); s = fmaf(d3, d3, s);
        }
        j0 = c0 + n4 * 4;
    }
    for (int64_t j = j0 + threadIdx.x; j < c1; j += 256) { const float d = pa[j] - pb[j]; s = fmaf(d, d, s); }
    s = warp_sum(s);
    if (lane == 0) part[warp] = s;
    __syncthreads();
    if (warp == 0) {
        float t = lane < 8 ? part[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0 && c0 < c1) atomicAdd(out + r, 0.5f * t);
    }
}
// many short rows: one warp per row
__global__ void __launch_bounds__(256) feat_mse_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t rows,
                                                            int64_t cols) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
        float s = 0.f;
        for (int64_t j = lane; j < cols; j += 32) { const float d = a[r * cols + j] - b[r * cols + j]; s = fmaf(d, d, s); }
        s = warp_sum(s);
        if (lane == 0) out[r] = 0.5f * s;
    }
}
// da = g[r] * (a - b), db = -da (either may be null)
__global__ void __launch_bounds__(256) feat_mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ g,
                                                           float* __restrict__ da, float* __restrict__ db, int64_t rows, int64_t cols) {
    pdl_sync();
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float d = g[i / cols] * (a[i] - b[i]);
        if (da) da[i] = d;
        if (db) db[i] = -d;
    }
}
// out = -log(s * p + c): (s, c) = (1, 1e-3) for the original images, (-1, 1 + 1e-3) for reconstructed / sampled    networks.py:276-278
__global__ void __launch_bounds__(256) neglog_fwd_kernel(const float* __restrict__ p, float* __restrict__ out, int64_t n, float s, float c) {
    pdl_sync();
    // clamped at 100 like torch's binary_cross_entropy (log >= -100): only reachable when the argument underflows to 0
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = fminf(-logf(fmaf(s, p[i], c)), 100.f);
}
__global__ void __launch_bounds__(256) neglog_bwd_kernel(const float* __restrict__ p, const float* __restrict__ g, float* __restrict__ dp, int64_t n, float s,
                                                         float c) {
    pdl_sync();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dp[i] = -g[i] * s / fmaf(s, p[i], c);
}
// out[0] = scale * sum smooth_l1(a - b), beta = 1 (single block, deterministic)                                   networks.py:279
__global__ void __launch_bounds__(256) smooth_l1_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n,
                                                            float scale) {
    pdl_sync();
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float d = fabsf(a[i] - b[i]);
        s += (double)(d < 1.f ? 0.5f * d * d : d - 0.5f);
    }
    s = warp_sum(s);
    __shared__ double sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < 8; ++i) t += sh[i];
        out[0] = (float)(t * (double)scale);
    }
}
__global__ void __launch_bounds__(256) smooth_l1_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ g,
                                                            float* __restrict__ da, float* __restrict__ db, int64_t n, float scale) {
    pdl_sync();
    const float gs = (g ? *g : 1.f) * scale;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float d = a[i] - b[i];
        const float v = gs * fminf(fmaxf(d, -1.f), 1.f);
        if (da) da[i] = v;
        if (db) db[i] = -v;
    }
}
// kl[r] = -0.5 * sum_j (-exp(lv) - mu^2 + lv + 1)   (one warp per row)                                           networks.py:270
__global__ void __launch_bounds__(256) kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, int64_t ld, float* __restrict__ kl,
                                                     int64_t rows, int zdim) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
        float acc = 0.f;
        for (int j = lane; j < zdim; j += 32) {
            const float m = mu[r * ld + j], l = lv[r * ld + j];
            acc += -expf(l) - m * m + l + 1.f;
        }
        acc = warp_sum(acc);
        if (lane == 0) kl[r] = -0.5f * acc;
    }
}

// out[b] = -log(s[b, label[b]]) (the pick of F.cross_entropy after the softmax); bwd: ds[b, j] = -g[b] / s[b, label[b]] at j = label[b]
__global__ void __launch_bounds__(256) nll_pick_fwd_kernel(const float* __restrict__ s, const long long* __restrict__ label, float* __restrict__ out, int64_t rows,
                                                           int cols) {
    pdl_sync();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) out[r] = -logf(s[r * cols + (int)label[r]]);
}
__global__ void __launch_bounds__(256) nll_pick_bwd_kernel(const float* __restrict__ s, const long long* __restrict__ label, const float* __restrict__ g,
                                                           float* __restrict__ ds, int64_t rows, int cols) {
    pdl_sync();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols;
    const int j = (int)(i - r * cols);
    ds[i] = j == (int)label[r] ? -g[r] / s[i] : 0.f;
}

inline unsigned grid_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace
}  // namespace vp

using namespace vp;

extern "C" int vp_feature_mse_fwd(const float* a, const float* b, float* out, int64_t rows, int64_t cols, void* stream) {
    VP_CHECK_ARG(a && b && out && rows > 0 && cols > 0, "vp_feature_mse_fwd: bad arguments");
    if (cols < 4096 || rows > 65535) {
        launch_k(feat_mse_rows_kernel, dim3(grid_for(rows * 32)), dim3(256), 0, (cudaStream_t)stream, a, b, out, rows, cols);
        VP_CHECK_LAUNCH("vp_feature_mse_fwd");
        return VP_OK;
    }
    // long rows (a batch of 64 x 65 536 discriminator features): ~8192 elements per CTA, at most ~8 CTAs per SM in total
    int64_t splits = (cols + 8191) / 8192;
    const int64_t cap = (8 * (int64_t)num_sms() + rows - 1) / rows;
    splits = splits > cap ? cap : splits;
    splits = splits < 1 ? 1 : splits;
    zero_async(out, sizeof(float) * (size_t)rows, (cudaStream_t)stream);
    launch_k(feat_mse_fwd_kernel, dim3((unsigned)splits, (unsigned)rows), dim3(256), 0, (cudaStream_t)stream, a, b, out, rows, cols);
    VP_CHECK_LAUNCH("vp_feature_mse_fwd");
    return VP_OK;
}
extern "C" int vp_feature_mse_bwd(const float* a, const float* b, const float* g, float* da, float* db, int64_t rows, int64_t cols, void* stream) {
    VP_CHECK_ARG(a && b && g && (da || db) && rows > 0 && cols > 0, "vp_feature_mse_bwd: bad arguments");
    launch_k(feat_mse_bwd_kernel, dim3(grid_for(rows * cols)), dim3(256), 0, (cudaStream_t)stream, a, b, g, da, db, rows, cols);
    VP_CHECK_LAUNCH("vp_feature_mse_bwd");
    return VP_OK;
}
extern "C" int vp_neglog_fwd(const float* p, float* out, int64_t n, float sign, float offset, void* stream) {
    VP_CHECK_ARG(p && out && n > 0, "vp_neglog_fwd: bad arguments");
    launch_k(neglog_fwd_kernel, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, p, out, n, sign, offset);
    VP_CHECK_LAUNCH("vp_neglog_fwd");
    return VP_OK;
}
extern "C" int vp_neglog_bwd(const float* p, const float* g, float* dp, int64_t n, float sign, float offset, void* stream) {
    VP_CHECK_ARG(p && g && dp && n > 0, "vp_neglog_bwd: bad arguments");
    launch_k(neglog_bwd_kernel, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, p, g, dp, n, sign, offset);
    VP_CHECK_LAUNCH("vp_neglog_bwd");
    return VP_OK;
}
extern "C" int vp_smooth_l1_sum_fwd(const float* a, const float* b, float* out, int64_t n, float scale, void* stream) {
    VP_CHECK_ARG(a && b && out && n > 0, "vp_smooth_l1_sum_fwd: bad arguments");
    launch_k(smooth_l1_fwd_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, a, b, out, n, scale);
    VP_CHECK_LAUNCH("vp_smooth_l1_sum_fwd");
    return VP_OK;
}
extern "C" int vp_smooth_l1_sum_bwd(const float* a, const float* b, const float* g, float* da, float* db, int64_t n, float scale, void* stream) {
    VP_CHECK_ARG(a && b && (da || db) && n > 0, "vp_smooth_l1_sum_bwd: bad arguments");
    launch_k(smooth_l1_bwd_kernel, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, a, b, g, da, db, n, scale);
    VP_CHECK_LAUNCH("vp_smooth_l1_sum_bwd");
    return VP_OK;
}
extern "C" int vp_kl_fwd(const float* mu, const float* logvar, int64_t ld, float* kl, int64_t rows, int zdim, void* stream) {
    VP_CHECK_ARG(mu && logvar && kl && rows > 0 && zdim > 0 && ld >= zdim, "vp_kl_fwd: bad arguments");
    launch_k(kl_fwd_kernel, dim3(grid_for(rows * 32)), dim3(256), 0, (cudaStream_t)stream, mu, logvar, ld, kl, rows, zdim);
    VP_CHECK_LAUNCH("vp_kl_fwd");
    return VP_OK;
}

extern "C" int vp_nll_pick_fwd(const float* s, const int64_t* label, float* out, int64_t rows, int cols, void* stream) {
    VP_CHECK_ARG(s && label && out && rows > 0 && cols > 0, "vp_nll_pick_fwd: bad arguments");
    launch_k(nll_pick_fwd_kernel, dim3(grid_for(rows)), dim3(256), 0, (cudaStream_t)stream, s, (const long long*)label, out, rows, cols);
    VP_CHECK_LAUNCH("vp_nll_pick_fwd");
    return VP_OK;
}
extern "C" int vp_nll_pick_bwd(const float* s, const int64_t* label, const float* g, float* ds, int64_t rows, int cols, void* stream) {
    VP_CHECK_ARG(s && label && g && ds && rows > 0 && cols > 0, "vp_nll_pick_bwd: bad arguments");
    launch_k(nll_pick_bwd_kernel, dim3(grid_for(rows * cols)), dim3(256), 0, (cudaStream_t)stream, s, (const long long*)label, g, ds, rows, cols);
    VP_CHECK_LAUNCH("vp_nll_pick_bwd");
    return VP_OK;
}
