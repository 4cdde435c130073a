// Element-wise / small-reduction operators of models/blocks.py and models/network_Style_GAN.py around the contractions
// (channels-last activations, T = fp32 check mode or bf16): channel concat / slice, AddCoords, bilinear x2 up-sampling,
// adaptive average pooling, the SCSE gate, the label-gated blend of myConv2d, softmax over rows, a batched matmul for the
// attention block, dice on probabilities and the depthwise edge filter.  All HBM-bound: coalesced along the channel axis,
// one pass each, launched with programmatic dependent launch like every other kernel of the library.
#include "common.cuh"
namespace vp { void* splitk_workspace(size_t bytes); }     // tapgemm_tc.cu: the registered scratch buffer, or null when too small

namespace vp {
namespace {

inline unsigned grid_for(int64_t n, int per = 1) {
    int64_t b = (n + 256LL * per - 1) / (256LL * per);
    const int64_t cap = (int64_t)num_sms() * 16;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

// ---- dst[r, dst_off : dst_off + nc] = src[r, src_off : src_off + nc]  (concat = two calls, slice = one) -----------------
template <typename T>
__global__ void __launch_bounds__(256) copy_channels_kernel(const T* __restrict__ src, int src_c, int src_off, T* __restrict__ dst, int dst_c,
                                                            int dst_off, int nc, int64_t rows, int accumulate) {
    pdl_sync();
    const int64_t total = rows * nc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / nc;
        const int c = (int)(i - r * nc);
        const float v = Cvt<T>::ld(src + r * src_c + src_off + c);
        T* d = dst + r * dst_c + dst_off + c;
        Cvt<T>::st(d, accumulate ? Cvt<T>::ld(d) + v : v);
    }
}

// 2-byte elements, everything a multiple of 8 channels, no accumulation (the slice after a padded data gradient, concatenations
// of 32 / 64-channel maps): one 16-byte chunk per thread
__global__ void __launch_bounds__(256) copy_channels16_kernel(const uint4* __restrict__ src, int src_c8, int src_off8, uint4* __restrict__ dst, int dst_c8,
                                                              int dst_off8, int nc8, int64_t rows) {
    pdl_sync();
    const int64_t total = rows * nc8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / nc8;
        const int c = (int)(i - r * nc8);
        dst[r * dst_c8 + dst_off8 + c] = src[r * src_c8 + src_off8 + c];
    }
}

// ---- AddCoords (models/blocks.py:97-112): out[..., :c] = x, out[..., c] = column index, out[..., c+1] = row index ---------
template <typename T>
__global__ void __launch_bounds__(256) add_coords_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t n, int h, int w, int c, int normalize) {
    pdl_sync();
    const int co = c + 2;
    const int64_t total = n * h * w * co;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = i / co;
        const int ch = (int)(i - pix * co);
        float v;
        if (ch < c) {
            v = Cvt<T>::ld(x + pix * c + ch);
        } else {
            const int xx = (int)(pix % w), yy = (int)((pix / w) % h);
            if (ch == c) v = normalize ? ((float)xx / (float)w - 0.5f) / 0.5f : (float)xx;
            else v = normalize ? ((float)yy / (float)h - 0.5f) / 0.5f : (float)yy;
        }
        Cvt<T>::st(out + i, v);
    }
}

// ---- bilinear x2, align_corners = False (F.interpolate(scale_factor=2, mode='bilinear'), models/blocks.py:145) --------------
// source index of output o: s = max(0, (o + .5)/2 - .5); i0 = floor(s), i1 = min(i0 + 1, n - 1), l1 = s - i0, l0 = 1 - l1
__device__ __forceinline__ void up2_src(int o, int n, int& i0, int& i1, float& l0, float& l1) {
    float s = ((float)o + 0.5f) * 0.5f - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    i1 = i0 + 1 < n ? i0 + 1 : n - 1;
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}
template <typename T>
__global__ void __launch_bounds__(256) up2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, int h, int w, int c) {
    pdl_sync();
    const int ho = 2 * h, wo = 2 * w;
    const int64_t total = n * ho * wo * c;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        int64_t p = i / c;
        const int ox = (int)(p % wo); p /= wo;
        const int oy = (int)(p % ho);
        const int64_t img = p / ho;
        int y0, y1, x0, x1;
        float ly0, ly1, lx0, lx1;
        up2_src(oy, h, y0, y1, ly0, ly1);
        up2_src(ox, w, x0, x1, lx0, lx1);
        const T* b = x + img * h * w * c + ch;
        const float v = ly0 * (lx0 * Cvt<T>::ld(b + ((int64_t)y0 * w + x0) * c) + lx1 * Cvt<T>::ld(b + ((int64_t)y0 * w + x1) * c)) +
                        ly1 * (lx0 * Cvt<T>::ld(b + ((int64_t)y1 * w + x0) * c) + lx1 * Cvt<T>::ld(b + ((int64_t)y1 * w + x1) * c));
        Cvt<T>::st(y + i, v);
    }
}
// exact adjoint in gather form: input pixel (iy, ix) collects from the <= 4 x 4 output pixels whose stencil touches it
__device__ __forceinline__ float up2_weight(int o, int n, int i) {
    int i0, i1;
    float l0, l1;
    up2_src(o, n, i0, i1, l0, l1);
    return (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
}
template <typename T>
__global__ void __launch_bounds__(256) up2_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int64_t n, int h, int w, int c) {
    pdl_sync();
    const int ho = 2 * h, wo = 2 * w;
    const int64_t total = n * h * w * c;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        int64_t p = i / c;
        const int ix = (int)(p % w); p /= w;
        const int iy = (int)(p % h);
        const int64_t img = p / h;
        const T* b = dy + img * ho * wo * c + ch;
        float acc = 0.f;
        for (int oy = 2 * iy - 1; oy <= 2 * iy + 2; ++oy) {
            if (oy < 0 || oy >= ho) continue;
            const float wy = up2_weight(oy, h, iy);
            if (wy == 0.f) continue;
            for (int ox = 2 * ix - 1; ox <= 2 * ix + 2; ++ox) {
                if (ox < 0 || ox >= wo) continue;
                const float wx = up2_weight(ox, w, ix);
                if (wx != 0.f) acc = fmaf(wy * wx, Cvt<T>::ld(b + ((int64_t)oy * wo + ox) * c), acc);
            }
        }
        Cvt<T>::st(dx + i, acc);
    }
}

// ---- adaptive average pooling (nn.AdaptiveAvgPool2d): bin i covers [floor(i*H/oh), ceil((i+1)*H/oh)) -----------------------
template <typename T>
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, int h, int w, int c, int oh, int ow) {
    pdl_sync();
    // one block per (image, bin); threads stride over channels (coalesced), rows of the bin are walked sequentially
    const int64_t bin = blockIdx.x;
    const int bx = (int)(bin % ow), by = (int)((bin / ow) % oh);
    const int64_t img = bin / ((int64_t)ow * oh);
    const int y0 = (by * h) / oh, y1 = ((by + 1) * h + oh - 1) / oh, x0 = (bx * w) / ow, x1 = ((bx + 1) * w + ow - 1) / ow;
    const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float acc = 0.f;
        for (int yy = y0; yy < y1; ++yy)
            for (int xx = x0; xx < x1; ++xx) acc += Cvt<T>::ld(x + ((img * h + yy) * w + xx) * c + ch);
        Cvt<T>::st(y + bin * c + ch, acc * inv);
    }
}
// large bins (global pooling of a 256 x 256 map = 65 536 pixels per bin): the rows of a bin are split over gridDim.y CTAs, threads
// cover (pixel lane, channel), partial sums are added into an fp32 scratch [bins][c]; avgpool_finish_kernel scales and converts.
template <typename T>
__global__ void __launch_bounds__(256) avgpool_split_kernel(const T* __restrict__ x, float* __restrict__ acc, int h, int w, int c, int oh, int ow) {
    pdl_sync();
    const int64_t bin = blockIdx.x;
    const int bx = (int)(bin % ow), by = (int)((bin / ow) % oh);
    const int64_t img = bin / ((int64_t)ow * oh);
    const int y0 = (by * h) / oh, y1 = ((by + 1) * h + oh - 1) / oh, x0 = (bx * w) / ow, x1 = ((bx + 1) * w + ow - 1) / ow;
    const int rows = y1 - y0, rps = (rows + gridDim.y - 1) / gridDim.y;
    const int ya = y0 + blockIdx.y * rps, yb = min(ya + rps, y1);
    const int lanes = c <= 256 ? 256 / c : 1, wb = x1 - x0;
    const int lane = threadIdx.x / c;
    if (c <= 256 && lane >= lanes) return;
    for (int ch = c <= 256 ? (int)(threadIdx.x % c) : (int)threadIdx.x; ch < c; ch += 256) {
        float a = 0.f;
        const int64_t npix = (int64_t)(yb > ya ? yb - ya : 0) * wb;
        for (int64_t p = (c <= 256 ? lane : 0); p < npix; p += lanes) {
            const int yy = ya + (int)(p / wb), xx = x0 + (int)(p % wb);
            a += Cvt<T>::ld(x + ((img * h + yy) * w + xx) * c + ch);
        }
        if (npix > 0) atomicAdd(acc + bin * c + ch, a);
        if (c <= 256) break;
    }
}
template <typename T>
__global__ void __launch_bounds__(256) avgpool_finish_kernel(const float* __restrict__ acc, T* __restrict__ y, int64_t total, int h, int w, int c, int oh, int ow) {
    pdl_sync();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t bin = i / c;
        const int bx = (int)(bin % ow), by = (int)((bin / ow) % oh);
        const int y0 = (by * h) / oh, y1 = ((by + 1) * h + oh - 1) / oh, x0 = (bx * w) / ow, x1 = ((bx + 1) * w + ow - 1) / ow;
        Cvt<T>::st(y + i, acc[i] / (float)((y1 - y0) * (x1 - x0)));
    }
}
template <typename T>
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int64_t n, int h, int w, int c, int oh, int ow) {
    pdl_sync();
    const int64_t total = n * h * w * c;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        int64_t p = i / c;
        const int xx = (int)(p % w); p /= w;
        const int yy = (int)(p % h);
        const int64_t img = p / h;
        float acc = 0.f;
        // bins may overlap when h % oh != 0: a pixel can belong to two bins per axis
        for (int by = (yy * oh) / h - 1; by <= (yy * oh) / h + 1; ++by) {
            if (by < 0 || by >= oh) continue;
            const int y0 = (by * h) / oh, y1 = ((by + 1) * h + oh - 1) / oh;
            if (yy < y0 || yy >= y1) continue;
            for (int bx = (xx * ow) / w - 1; bx <= (xx * ow) / w + 1; ++bx) {
                if (bx < 0 || bx >= ow) continue;
                const int x0 = (bx * w) / ow, x1 = ((bx + 1) * w + ow - 1) / ow;
                if (xx < x0 || xx >= x1) continue;
                acc += Cvt<T>::ld(dy + ((img * oh + by) * ow + bx) * c + ch) / (float)((y1 - y0) * (x1 - x0));
            }
        }
        Cvt<T>::st(dx + i, acc);
    }
}

// ---- SCSE gate (models/blocks.py:64-65): y = x * cse[n, c] + x * sse[n, pixel] -----------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) scse_fwd_kernel(const T* __restrict__ x, const T* __restrict__ cse, const T* __restrict__ sse, T* __restrict__ y,
                                                       int64_t n, int64_t hw, int c) {
    pdl_sync();
    const int64_t total = n * hw * c;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        const int64_t pix = i / c, img = pix / hw;
        Cvt<T>::st(y + i, Cvt<T>::ld(x + i) * (Cvt<T>::ld(cse + img * c + ch) + Cvt<T>::ld(sse + pix)));
    }
}
// dx = dy * (cse + sse);  dsse[pixel] = sum_c dy*x (one warp per pixel);  dcse[n, c] = sum_pixels dy*x (fp32 atomics into a
// zeroed [n, c] buffer, converted by the caller)
template <typename T>
__global__ void __launch_bounds__(256) scse_bwd_kernel(const T* __restrict__ x, const T* __restrict__ cse, const T* __restrict__ sse,
                                                       const T* __restrict__ dy, T* __restrict__ dx, float* __restrict__ dcse, T* __restrict__ dsse,
                                                       int64_t n, int64_t hw, int c) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t pix = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < n * hw; pix += warps) {
        const int64_t img = pix / hw;
        const float s = Cvt<T>::ld(sse + pix);
        float acc = 0.f;
        for (int ch = lane; ch < c; ch += 32) {
            const float g = Cvt<T>::ld(dy + pix * c + ch), xv = Cvt<T>::ld(x + pix * c + ch);
            Cvt<T>::st(dx + pix * c + ch, g * (Cvt<T>::ld(cse + img * c + ch) + s));
            const float gx = g * xv;
            acc += gx;
            atomicAdd(dcse + img * c + ch, gx);
        }
        acc = warp_sum(acc);
        if (lane == 0) Cvt<T>::st(dsse + pix, acc);
    }
}

// ---- myConv2d blend (models/network_Style_GAN.py:78-79): y = a1 * (1 - label[n]) + a2 * label[n] ----------------------------
template <typename T>
__global__ void __launch_bounds__(256) blend_fwd_kernel(const T* __restrict__ a1, const T* __restrict__ a2, const float* __restrict__ label,
                                                        T* __restrict__ y, int64_t n, int64_t per) {
    pdl_sync();
    const int64_t total = n * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float l = label[i / per];
        Cvt<T>::st(y + i, Cvt<T>::ld(a1 + i) * (1.f - l) + Cvt<T>::ld(a2 + i) * l);
    }
}
template <typename T>
__global__ void __launch_bounds__(256) blend_bwd_kernel(const T* __restrict__ dy, const float* __restrict__ label, T* __restrict__ d1, T* __restrict__ d2,
                                                        int64_t n, int64_t per) {
    pdl_sync();
    const int64_t total = n * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float l = label[i / per], g = Cvt<T>::ld(dy + i);
        Cvt<T>::st(d1 + i, g * (1.f - l));
        Cvt<T>::st(d2 + i, g * l);
    }
}

// ---- row softmax (nn.Softmax(dim=-1) of the attention block, models/blocks.py:73,87; also the Style discriminator head) ----
// one warp per row; backward: dx = y * (dy - sum(dy * y))
template <typename T>
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int cols) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
        const T* xr = x + r * cols;
        float m = -INFINITY;
        for (int j = lane; j < cols; j += 32) m = fmaxf(m, Cvt<T>::ld(xr + j));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int j = lane; j < cols; j += 32) s += expf(Cvt<T>::ld(xr + j) - m);
        s = warp_sum(s);
        const float inv = 1.f / s;
        for (int j = lane; j < cols; j += 32) Cvt<T>::st(y + r * cols + j, expf(Cvt<T>::ld(xr + j) - m) * inv);
    }
}
template <typename T>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, int64_t rows, int cols) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
        float s = 0.f;
        for (int j = lane; j < cols; j += 32) s = fmaf(Cvt<T>::ld(dy + r * cols + j), Cvt<T>::ld(y + r * cols + j), s);
        s = warp_sum(s);
        for (int j = lane; j < cols; j += 32) {
            const float yv = Cvt<T>::ld(y + r * cols + j);
            Cvt<T>::st(dx + r * cols + j, yv * (Cvt<T>::ld(dy + r * cols + j) - s));
        }
    }
}

// ---- batched matmul for the attention block (torch.bmm, models/blocks.py:86,90): C[b] = op(A[b]) . op(B[b]) ----------------
// element (i, k) of op(A) at A[b*sa_b + i*sa_i + k*sa_k], element (k, j) of op(B) at B[b*sb_b + k*sb_k + j*sb_j]; fp32 accumulation.
// 32 x 32 output tile per block, 32-deep k panels through shared memory.  The attention maps here are (h*w) x (h*w) with
// c/8 .. c-deep reductions: a few MFLOP per image, far from the step's contractions.
template <typename T>
__global__ void __launch_bounds__(256) bmm_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, int M, int N, int K,
                                                  int64_t sa_b, int64_t sa_i, int64_t sa_k, int64_t sb_b, int64_t sb_k, int64_t sb_j, int accumulate) {
    pdl_sync();
    __shared__ float sA[32][33], sB[32][33];
    const int b = blockIdx.z, i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 8 rows of threads: each thread owns 4 output rows
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int r = ty; r < 32; r += 8) {
            const int i = i0 + r, k = k0 + tx;
            sA[r][tx] = (i < M && k < K) ? Cvt<T>::ld(A + b * sa_b + i * sa_i + k * sa_k) : 0.f;
            const int kk = k0 + r, j = j0 + tx;
            sB[r][tx] = (kk < K && j < N) ? Cvt<T>::ld(B + b * sb_b + kk * sb_k + j * sb_j) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const float bv = sB[k][tx];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fmaf(sA[ty + 8 * q][k], bv, acc[q]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = i0 + ty + 8 * q, j = j0 + tx;
        if (i < M && j < N) {
            T* c = C + ((int64_t)b * M + i) * N + j;
            Cvt<T>::st(c, accumulate ? Cvt<T>::ld(c) + acc[q] : acc[q]);
        }
    }
}

// ---- out = gamma[0] * a + x (attention residual, models/blocks.py:93) and its gradients ---------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) scale_add_kernel(const float* __restrict__ gamma, const T* __restrict__ a, const T* __restrict__ x, T* __restrict__ y, int64_t n) {
    pdl_sync();
    const float g = *gamma;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        Cvt<T>::st(y + i, fmaf(g, Cvt<T>::ld(a + i), Cvt<T>::ld(x + i)));
}
// acc[0] += sum(a * b)  (double atomics; acc zeroed by the caller)
template <typename T>
__global__ void __launch_bounds__(256) dot_kernel(const T* __restrict__ a, const T* __restrict__ b, double* acc, int64_t n) {
    pdl_sync();
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        s += (double)(Cvt<T>::ld(a + i) * Cvt<T>::ld(b + i));
    s = warp_sum(s);
    __shared__ double sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < 8; ++i) t += sh[i];
        atomicAdd(acc, t);
    }
}

// ---- dice on probabilities (tools/ops.py:12-19): 1 - mean_b (2 sum(p t) + 1) / (sum p + sum t + 1) ------------------------
// acc[row][3] = (sum p*t, sum p, sum t) in double; the last block to finish writes the scalar
__global__ void __launch_bounds__(256) dice_fwd_kernel(const float* __restrict__ p, const float* __restrict__ t, int64_t rows, int64_t per, float smooth,
                                                       double* acc, unsigned int* counter, float* loss) {
    pdl_sync();
    const int64_t row = blockIdx.y;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
        const float pv = p[row * per + i], tv = t[row * per + i];
        s0 = fmaf(pv, tv, s0); s1 += pv; s2 += tv;
    }
    double d0 = warp_sum((double)s0), d1 = warp_sum((double)s1), d2 = warp_sum((double)s2);
    __shared__ double sh[3][8];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wi] = d0; sh[1][wi] = d1; sh[2][wi] = d2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0, a1 = 0, a2 = 0;
        for (int i = 0; i < 8; ++i) { a0 += sh[0][i]; a1 += sh[1][i]; a2 += sh[2][i]; }
        atomicAdd(acc + row * 3, a0); atomicAdd(acc + row * 3 + 1, a1); atomicAdd(acc + row * 3 + 2, a2);
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double score = 0;
        for (int64_t r = 0; r < rows; ++r) {
            const volatile double* a = acc + r * 3;
            score += (2.0 * a[0] + (double)smooth) / (a[1] + a[2] + (double)smooth);
        }
        *loss = (float)(1.0 - score / (double)rows);
        *counter = 0;
    }
}
// d loss / d p = -(1/rows) * (2 t D - N) / D^2,  N = 2 sum(p t) + s, D = sum p + sum t + s
__global__ void __launch_bounds__(256) dice_bwd_kernel(const float* __restrict__ t, int64_t rows, int64_t per, float smooth, const double* __restrict__ acc,
                                                       const float* __restrict__ gscale, float* __restrict__ dp) {
    pdl_sync();
    const int64_t row = blockIdx.y;
    const double N = 2.0 * acc[row * 3] + (double)smooth, D = acc[row * 3 + 1] + acc[row * 3 + 2] + (double)smooth;
    const float g = (gscale ? *gscale : 1.f) / (float)rows;
    const float k1 = (float)(2.0 / D), k0 = (float)(N / (D * D));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x)
        dp[row * per + i] = -g * (k1 * t[row * per + i] - k0);
}

// ---- |depthwise 3x3 edge filter| (tools/ops.py:187-211): e = |x - mean of the 8 neighbours| with zero padding -------------
// kernel [[-1,-1,-1],[-1,8,-1],[-1,-1,-1]] / 8 on single-channel fp32 maps; backward is the same (symmetric) filter applied to
// dy * sign(pre-abs response)
__global__ void __launch_bounds__(256) edge_fwd_kernel(const float* __restrict__ x, float* __restrict__ e, float* __restrict__ sgn, int64_t n, int h, int w) {
    pdl_sync();
    const int64_t total = n * h * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xx = (int)(i % w), yy = (int)((i / w) % h);
        const float* b = x + (i - (int64_t)yy * w - xx);
        float acc = 0.f;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int y2 = yy + dy, x2 = xx + dx;
                if (y2 < 0 || y2 >= h || x2 < 0 || x2 >= w) continue;
                acc += ((dy | dx) == 0 ? 8.f : -1.f) * b[(int64_t)y2 * w + x2];
            }
        acc *= 0.125f;
        e[i] = fabsf(acc);
        if (sgn) sgn[i] = acc > 0.f ? 1.f : (acc < 0.f ? -1.f : 0.f);
    }
}
__global__ void __launch_bounds__(256) edge_bwd_kernel(const float* __restrict__ de, const float* __restrict__ sgn, float* __restrict__ dx_, int64_t n, int h, int w) {
    pdl_sync();
    const int64_t total = n * h * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xx = (int)(i % w), yy = (int)((i / w) % h);
        const int64_t base = i - (int64_t)yy * w - xx;
        float acc = 0.f;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int y2 = yy + dy, x2 = xx + dx;
                if (y2 < 0 || y2 >= h || x2 < 0 || x2 >= w) continue;
                const int64_t j = base + (int64_t)y2 * w + x2;
                acc += ((dy | dx) == 0 ? 8.f : -1.f) * de[j] * sgn[j];
            }
        dx_[i] = acc * 0.125f;
    }
}

}  // namespace
}  // namespace vp

using namespace vp;

#define VP_DISPATCH_T(dtype, CALL_F32, CALL_BF16) \
    do {                                           \
        if ((dtype) == VP_F32) { CALL_F32; }       \
        else { CALL_BF16; }                        \
    } while (0)

extern "C" int vp_copy_channels(const void* src, int src_c, int src_off, void* dst, int dst_c, int dst_off, int nc, int64_t rows, int dtype,
                                int accumulate, void* stream) {
    VP_CHECK_ARG(src && dst && nc > 0 && rows >= 0 && src_off >= 0 && dst_off >= 0 && src_off + nc <= src_c && dst_off + nc <= dst_c,
                 "vp_copy_channels: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_copy_channels: bad dtype %d", dtype);
    if (rows == 0) return VP_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == VP_BF16 && !accumulate && ((src_c | src_off | dst_c | dst_off | nc) & 7) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
        launch_k(copy_channels16_kernel, dim3(grid_for(rows * (nc / 8))), dim3(256), 0, s, (const uint4*)src, src_c / 8, src_off / 8, (uint4*)dst, dst_c / 8,
                 dst_off / 8, nc / 8, rows);
        VP_CHECK_LAUNCH("vp_copy_channels");
        return VP_OK;
    }
    VP_DISPATCH_T(dtype, launch_k(copy_channels_kernel<float>, dim3(grid_for(rows * nc)), dim3(256), 0, s, (const float*)src, src_c, src_off, (float*)dst, dst_c, dst_off, nc, rows, accumulate),
                  launch_k(copy_channels_kernel<bf16>, dim3(grid_for(rows * nc)), dim3(256), 0, s, (const bf16*)src, src_c, src_off, (bf16*)dst, dst_c, dst_off, nc, rows, accumulate));
    VP_CHECK_LAUNCH("vp_copy_channels");
    return VP_OK;
}

extern "C" int vp_add_coords(const void* x, void* out, int dtype, int64_t n, int h, int w, int c, int normalize, void* stream) {
    VP_CHECK_ARG(x && out && n > 0 && h > 0 && w > 0 && c > 0, "vp_add_coords: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_add_coords: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(n * h * w * (c + 2));
    VP_DISPATCH_T(dtype, launch_k(add_coords_kernel<float>, dim3(g), dim3(256), 0, s, (const float*)x, (float*)out, n, h, w, c, normalize),
                  launch_k(add_coords_kernel<bf16>, dim3(g), dim3(256), 0, s, (const bf16*)x, (bf16*)out, n, h, w, c, normalize));
    VP_CHECK_LAUNCH("vp_add_coords");
    return VP_OK;
}

extern "C" int vp_upsample2x_fwd(const void* x, void* y, int dtype, int64_t n, int h, int w, int c, void* stream) {
    VP_CHECK_ARG(x && y && n > 0 && h > 0 && w > 0 && c > 0, "vp_upsample2x_fwd: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_upsample2x_fwd: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(n * h * w * c * 4);
    VP_DISPATCH_T(dtype, launch_k(up2_fwd_kernel<float>, dim3(g), dim3(256), 0, s, (const float*)x, (float*)y, n, h, w, c),
                  launch_k(up2_fwd_kernel<bf16>, dim3(g), dim3(256), 0, s, (const bf16*)x, (bf16*)y, n, h, w, c));
    VP_CHECK_LAUNCH("vp_upsample2x_fwd");
    return VP_OK;
}

extern "C" int vp_upsample2x_bwd(const void* dy, void* dx, int dtype, int64_t n, int h, int w, int c, void* stream) {
    VP_CHECK_ARG(dy && dx && n > 0 && h > 0 && w > 0 && c > 0, "vp_upsample2x_bwd: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_upsample2x_bwd: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(n * h * w * c);
    VP_DISPATCH_T(dtype, launch_k(up2_bwd_kernel<float>, dim3(g), dim3(256), 0, s, (const float*)dy, (float*)dx, n, h, w, c),
                  launch_k(up2_bwd_kernel<bf16>, dim3(g), dim3(256), 0, s, (const bf16*)dy, (bf16*)dx, n, h, w, c));
    VP_CHECK_LAUNCH("vp_upsample2x_bwd");
    return VP_OK;
}

extern "C" int vp_avgpool_fwd(const void* x, void* y, int dtype, int64_t n, int h, int w, int c, int oh, int ow, void* stream) {
    VP_CHECK_ARG(x && y && n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0 && oh <= h && ow <= w, "vp_avgpool_fwd: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_avgpool_fwd: bad dtype %d", dtype);
    VP_CHECK_ARG(n * oh * ow < 0x7fffffff, "vp_avgpool_fwd: too many bins");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = (unsigned)(n * oh * ow);
    const int64_t bin_pix = (int64_t)((h + oh - 1) / oh) * ((w + ow - 1) / ow);
    if (bin_pix >= 2048 && (int64_t)g * 4 < 148 * 8) {
        // few, large bins: split each bin's rows over enough CTAs to fill the GPU (fp32 scratch from the library workspace)
        const size_t bytes = sizeof(float) * (size_t)g * c;
        float* acc = (float*)vp::splitk_workspace(bytes);
        if (acc) {
            int splits = (148 * 8 + (int)g - 1) / (int)g;
            const int rows = (h + oh - 1) / oh;
            splits = splits > rows ? rows : splits;
            zero_async(acc, bytes, s);
            VP_DISPATCH_T(dtype, launch_k(avgpool_split_kernel<float>, dim3(g, (unsigned)splits), dim3(256), 0, s, (const float*)x, acc, h, w, c, oh, ow),
                          launch_k(avgpool_split_kernel<bf16>, dim3(g, (unsigned)splits), dim3(256), 0, s, (const bf16*)x, acc, h, w, c, oh, ow));
            VP_CHECK_LAUNCH("vp_avgpool_fwd");
            const int64_t total = (int64_t)g * c;
            VP_DISPATCH_T(dtype, launch_k(avgpool_finish_kernel<float>, dim3(grid_for(total)), dim3(256), 0, s, (const float*)acc, (float*)y, total, h, w, c, oh, ow),
                          launch_k(avgpool_finish_kernel<bf16>, dim3(grid_for(total)), dim3(256), 0, s, (const float*)acc, (bf16*)y, total, h, w, c, oh, ow));
            VP_CHECK_LAUNCH("vp_avgpool_fwd");
            return VP_OK;
        }
    }
    VP_DISPATCH_T(dtype, launch_k(avgpool_fwd_kernel<float>, dim3(g), dim3(256), 0, s, (const float*)x, (float*)y, n, h, w, c, oh, ow),
                  launch_k(avgpool_fwd_kernel<bf16>, dim3(g), dim3(256), 0, s, (const bf16*)x, (bf16*)y, n, h, w, c, oh, ow));
    VP_CHECK_LAUNCH("vp_avgpool_fwd");
    return VP_OK;
}

extern "C" int vp_avgpool_bwd(const void* dy, void* dx, int dtype, int64_t n, int h, int w, int c, int oh, int ow, void* stream) {
    VP_CHECK_ARG(dy && dx && n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0 && oh <= h && ow <= w, "vp_avgpool_bwd: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_avgpool_bwd: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(n * h * w * c);
    VP_DISPATCH_T(dtype, launch_k(avgpool_bwd_kernel<float>, dim3(g), dim3(256), 0, s, (const float*)dy, (float*)dx, n, h, w, c, oh, ow),
                  launch_k(avgpool_bwd_kernel<bf16>, dim3(g), dim3(256), 0, s, (const bf16*)dy, (bf16*)dx, n, h, w, c, oh, ow));
    VP_CHECK_LAUNCH("vp_avgpool_bwd");
    return VP_OK;
}

extern "C" int vp_scse_fwd(const void* x, const void* cse, const void* sse, void* y, int dtype, int64_t n, int64_t hw, int c, void* stream) {
    VP_CHECK_ARG(x && cse && sse && y && n > 0 && hw > 0 && c > 0, "vp_scse_fwd: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_scse_fwd: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(n * hw * c);
    VP_DISPATCH_T(dtype, launch_k(scse_fwd_kernel<float>, dim3(g), dim3(256), 0, s, (const float*)x, (const float*)cse, (const float*)sse, (float*)y, n, hw, c),
                  launch_k(scse_fwd_kernel<bf16>, dim3(g), dim3(256), 0, s, (const bf16*)x, (const bf16*)cse, (const bf16*)sse, (bf16*)y, n, hw, c));
    VP_CHECK_LAUNCH("vp_scse_fwd");
    return VP_OK;
}

extern "C" int vp_scse_bwd(const void* x, const void* cse, const void* sse, const void* dy, void* dx, float* dcse_f32, void* dsse, int dtype,
                           int64_t n, int64_t hw, int c, void* stream) {
    VP_CHECK_ARG(x && cse && sse && dy && dx && dcse_f32 && dsse && n > 0 && hw > 0 && c > 0, "vp_scse_bwd: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_scse_bwd: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    zero_async(dcse_f32, sizeof(float) * (size_t)(n * c), s);
    const unsigned g = grid_for(n * hw * 32);
    VP_DISPATCH_T(dtype, launch_k(scse_bwd_kernel<float>, dim3(g), dim3(256), 0, s, (const float*)x, (const float*)cse, (const float*)sse, (const float*)dy, (float*)dx, dcse_f32, (float*)dsse, n, hw, c),
                  launch_k(scse_bwd_kernel<bf16>, dim3(g), dim3(256), 0, s, (const bf16*)x, (const bf16*)cse, (const bf16*)sse, (const bf16*)dy, (bf16*)dx, dcse_f32, (bf16*)dsse, n, hw, c));
    VP_CHECK_LAUNCH("vp_scse_bwd");
    return VP_OK;
}

extern "C" int vp_blend_fwd(const void* a1, const void* a2, const float* label, void* y, int dtype, int64_t n, int64_t per, void* stream) {
    VP_CHECK_ARG(a1 && a2 && label && y && n > 0 && per > 0, "vp_blend_fwd: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_blend_fwd: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(n * per);
    VP_DISPATCH_T(dtype, launch_k(blend_fwd_kernel<float>, dim3(g), dim3(256), 0, s, (const float*)a1, (const float*)a2, label, (float*)y, n, per),
                  launch_k(blend_fwd_kernel<bf16>, dim3(g), dim3(256), 0, s, (const bf16*)a1, (const bf16*)a2, label, (bf16*)y, n, per));
    VP_CHECK_LAUNCH("vp_blend_fwd");
    return VP_OK;
}

extern "C" int vp_blend_bwd(const void* dy, const float* label, void* d1, void* d2, int dtype, int64_t n, int64_t per, void* stream) {
    VP_CHECK_ARG(dy && label && d1 && d2 && n > 0 && per > 0, "vp_blend_bwd: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_blend_bwd: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(n * per);
    VP_DISPATCH_T(dtype, launch_k(blend_bwd_kernel<float>, dim3(g), dim3(256), 0, s, (const float*)dy, label, (float*)d1, (float*)d2, n, per),
                  launch_k(blend_bwd_kernel<bf16>, dim3(g), dim3(256), 0, s, (const bf16*)dy, label, (bf16*)d1, (bf16*)d2, n, per));
    VP_CHECK_LAUNCH("vp_blend_bwd");
    return VP_OK;
}

extern "C" int vp_softmax_fwd(const void* x, void* y, int dtype, int64_t rows, int cols, void* stream) {
    VP_CHECK_ARG(x && y && rows > 0 && cols > 0, "vp_softmax_fwd: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_softmax_fwd: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(rows * 32);
    VP_DISPATCH_T(dtype, launch_k(softmax_fwd_kernel<float>, dim3(g), dim3(256), 0, s, (const float*)x, (float*)y, rows, cols),
                  launch_k(softmax_fwd_kernel<bf16>, dim3(g), dim3(256), 0, s, (const bf16*)x, (bf16*)y, rows, cols));
    VP_CHECK_LAUNCH("vp_softmax_fwd");
    return VP_OK;
}

extern "C" int vp_softmax_bwd(const void* y, const void* dy, void* dx, int dtype, int64_t rows, int cols, void* stream) {
    VP_CHECK_ARG(y && dy && dx && rows > 0 && cols > 0, "vp_softmax_bwd: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_softmax_bwd: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(rows * 32);
    VP_DISPATCH_T(dtype, launch_k(softmax_bwd_kernel<float>, dim3(g), dim3(256), 0, s, (const float*)y, (const float*)dy, (float*)dx, rows, cols),
                  launch_k(softmax_bwd_kernel<bf16>, dim3(g), dim3(256), 0, s, (const bf16*)y, (const bf16*)dy, (bf16*)dx, rows, cols));
    VP_CHECK_LAUNCH("vp_softmax_bwd");
    return VP_OK;
}

extern "C" int vp_bmm(const void* a, const void* b, void* c, int dtype, int batch, int m, int n, int k, int64_t sa_b, int64_t sa_i, int64_t sa_k,
                      int64_t sb_b, int64_t sb_k, int64_t sb_j, int accumulate, void* stream) {
    VP_CHECK_ARG(a && b && c && batch > 0 && m > 0 && n > 0 && k > 0 && batch <= 65535, "vp_bmm: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_bmm: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((n + 31) / 32, (m + 31) / 32, batch);
    VP_DISPATCH_T(dtype, launch_k(bmm_kernel<float>, grid, dim3(256), 0, s, (const float*)a, (const float*)b, (float*)c, m, n, k, sa_b, sa_i, sa_k, sb_b, sb_k, sb_j, accumulate),
                  launch_k(bmm_kernel<bf16>, grid, dim3(256), 0, s, (const bf16*)a, (const bf16*)b, (bf16*)c, m, n, k, sa_b, sa_i, sa_k, sb_b, sb_k, sb_j, accumulate));
    VP_CHECK_LAUNCH("vp_bmm");
    return VP_OK;
}

extern "C" int vp_scale_add(const float* gamma, const void* a, const void* x, void* y, int dtype, int64_t n, void* stream) {
    VP_CHECK_ARG(gamma && a && x && y && n > 0, "vp_scale_add: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_scale_add: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    VP_DISPATCH_T(dtype, launch_k(scale_add_kernel<float>, dim3(grid_for(n)), dim3(256), 0, s, gamma, (const float*)a, (const float*)x, (float*)y, n),
                  launch_k(scale_add_kernel<bf16>, dim3(grid_for(n)), dim3(256), 0, s, gamma, (const bf16*)a, (const bf16*)x, (bf16*)y, n));
    VP_CHECK_LAUNCH("vp_scale_add");
    return VP_OK;
}

extern "C" int vp_dot(const void* a, const void* b, double* acc, int dtype, int64_t n, void* stream) {
    VP_CHECK_ARG(a && b && acc && n > 0, "vp_dot: bad arguments");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_dot: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    zero_async(acc, sizeof(double), s);
    VP_DISPATCH_T(dtype, launch_k(dot_kernel<float>, dim3(grid_for(n, 4)), dim3(256), 0, s, (const float*)a, (const float*)b, acc, n),
                  launch_k(dot_kernel<bf16>, dim3(grid_for(n, 4)), dim3(256), 0, s, (const bf16*)a, (const bf16*)b, acc, n));
    VP_CHECK_LAUNCH("vp_dot");
    return VP_OK;
}

extern "C" int vp_dice_fwd(const float* p, const float* t, int64_t rows, int64_t per, float smooth, double* acc, unsigned int* counter, float* loss,
                           void* stream) {
    VP_CHECK_ARG(p && t && acc && counter && loss && rows > 0 && per > 0 && rows <= 65535, "vp_dice_fwd: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    zero_async(acc, sizeof(double) * 3 * (size_t)rows, s);
    int64_t bx = (per + 1023) / 1024;
    if (bx > 64) bx = 64;
    launch_k(dice_fwd_kernel, dim3((unsigned)bx, (unsigned)rows), dim3(256), 0, s, p, t, rows, per, smooth, acc, counter, loss);
    VP_CHECK_LAUNCH("vp_dice_fwd");
    return VP_OK;
}

extern "C" int vp_dice_bwd(const float* t, int64_t rows, int64_t per, float smooth, const double* acc, const float* gscale, float* dp, void* stream) {
    VP_CHECK_ARG(t && acc && dp && rows > 0 && per > 0 && rows <= 65535, "vp_dice_bwd: bad arguments");
    int64_t bx = (per + 1023) / 1024;
    if (bx > 64) bx = 64;
    launch_k(dice_bwd_kernel, dim3((unsigned)bx, (unsigned)rows), dim3(256), 0, (cudaStream_t)stream, t, rows, per, smooth, acc, gscale, dp);
    VP_CHECK_LAUNCH("vp_dice_bwd");
    return VP_OK;
}

extern "C" int vp_edge_fwd(const float* x, float* e, float* sign_or_null, int64_t n, int h, int w, void* stream) {
    VP_CHECK_ARG(x && e && n > 0 && h > 0 && w > 0, "vp_edge_fwd: bad arguments");
    launch_k(edge_fwd_kernel, dim3(grid_for(n * h * w)), dim3(256), 0, (cudaStream_t)stream, x, e, sign_or_null, n, h, w);
    VP_CHECK_LAUNCH("vp_edge_fwd");
    return VP_OK;
}

extern "C" int vp_edge_bwd(const float* de, const float* sign, float* dx, int64_t n, int h, int w, void* stream) {
    VP_CHECK_ARG(de && sign && dx && n > 0 && h > 0 && w > 0, "vp_edge_bwd: bad arguments");
    launch_k(edge_bwd_kernel, dim3(grid_for(n * h * w)), dim3(256), 0, (cudaStream_t)stream, de, sign, dx, n, h, w);
    VP_CHECK_LAUNCH("vp_edge_bwd");
    return VP_OK;
}
