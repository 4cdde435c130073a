// CUDA-core engine of the tap GEMM: fp32 check mode, and odd shapes (Cin=1/3, Cout=1/3, K % 64 != 0)
// in bf16 mode.  Classic 64x64x16 shared-memory tiling, 4x4 micro-tile per thread, fp32 accumulate.
#include <cstring>

#include "common.cuh"

namespace vp {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename T, typename TD>
__global__ void __launch_bounds__(NT) tapgemm_simt_kernel(const TapGemm p) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const T* __restrict__ A = (const T*)p.A;
    const T* __restrict__ W = (const T*)p.Wp;
    const int64_t M = (int64_t)p.n * p.gh * p.gw;
    const int KK = p.taps.ntaps * p.K;  // flattened reduction (tap, k)
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // loader coordinates: row lr (0..63), 4 consecutive kk starting at lk
    const int lr = tid >> 2;
    const int lk = (tid & 3) * 4;
    const int64_t am = m0 + lr;
    int an = 0, agy = 0, agx = 0;
    const bool arow_ok = am < M;
    if (arow_ok) {
        agx = (int)(am % p.gw);
        int64_t t = am / p.gw;
        agy = (int)(t % p.gh);
        an = (int)(t / p.gh);
    }
    const int bn_ = n0 + lr;
    const bool brow_ok = bn_ < p.N;

    const int tr = (tid >> 4) * 4;  // micro-tile rows
    const int tc = (tid & 15) * 4;  // micro-tile cols
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < KK; k0 += BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kk = k0 + lk + j;
            float av = 0.f, bv = 0.f;
            if (kk < KK) {
                const int t = kk / p.K;
                const int k = kk - t * p.K;
                if (arow_ok) {
                    const int iy = agy * p.as + p.taps.ty[t];
                    const int ix = agx * p.as + p.taps.tx[t];
                    if (iy >= 0 && iy < p.ha && ix >= 0 && ix < p.wa)
                        av = Cvt<T>::ld(A + (((int64_t)an * p.ha + iy) * p.wa + ix) * p.K + k);
                }
                if (brow_ok) bv = Cvt<T>::ld(W + ((int64_t)p.taps.widx[t] * p.N + bn_) * p.K + k);
            }
            As[lk + j][lr] = av;
            Bs[lk + j][lr] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][tr + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tc + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    TD* __restrict__ D = (TD*)p.D;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + tr + i;
        if (m >= M) continue;
        const int gx = (int)(m % p.gw);
        const int64_t t = m / p.gw;
        const int gy = (int)(t % p.gh);
        const int n = (int)(t / p.gh);
        const int oy = gy * p.ds + p.doy, ox = gx * p.ds + p.dox;
        if (oy >= p.hd || ox >= p.wd) continue;
        TD* drow = D + (((int64_t)n * p.hd + oy) * p.wd + ox) * p.N;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tc + j;
            if (c < p.N) {
                float v = acc[i][j];
                if (p.bias) v += p.bias[c];
                Cvt<TD>::st(drow + c, act_fwd(v, p.act, p.slope));
            }
        }
    }
}

// dWp[widx_t][gc][ac] += sum_m G[m][gc] * A[pix(m,t)][ac]; block = (gc tile, ac tile, tap*split)
template <typename T, typename TAcc>
__global__ void __launch_bounds__(NT) tapwgrad_simt_kernel(const TapWgrad p, int nsplit) {
    __shared__ float Gs[BK][BM + 4];
    __shared__ float As[BK][BN + 4];
    const T* __restrict__ G = (const T*)p.G;
    const T* __restrict__ A = (const T*)p.A;
    const int64_t M = (int64_t)p.n * p.gh * p.gw;
    const int tap = blockIdx.z / nsplit;
    const int split = blockIdx.z - tap * nsplit;
    const int ty = p.taps.ty[tap], tx = p.taps.tx[tap];
    const int64_t chunk = ((M + nsplit - 1) / nsplit + BK - 1) / BK * BK;
    const int64_t mbeg = split * chunk;
    const int64_t mend = (mbeg + chunk < M) ? mbeg + chunk : M;
    const int gc0 = blockIdx.x * BM;
    const int ac0 = blockIdx.y * BN;
    const int tid = threadIdx.x;
    // loader: row of the m-chunk lm (0..15), 4 consecutive channels starting at lc
    const int lm = tid >> 4;
    const int lc = (tid & 15) * 4;
    const int tr = (tid >> 4) * 4;
    const int tc = (tid & 15) * 4;
    TAcc acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = (TAcc)0;

    for (int64_t mb = mbeg; mb < mend; mb += BK) {
        const int64_t m = mb + lm;
        const bool ok = m < mend;
        int n = 0, gy = 0, gx = 0;
        if (ok) {
            gx = (int)(m % p.gw);
            int64_t t = m / p.gw;
            gy = (int)(t % p.gh);
            n = (int)(t / p.gh);
        }
        const int iy = gy * p.as + ty, ix = gx * p.as + tx;
        const bool aok = ok && iy >= 0 && iy < p.ha && ix >= 0 && ix < p.wa;
        const T* grow = G + m * p.GC;
        const T* arow = A + (((int64_t)n * p.ha + iy) * p.wa + ix) * p.AC;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gc = gc0 + lc + j, ac = ac0 + lc + j;
            Gs[lm][lc + j] = (ok && gc < p.GC) ? Cvt<T>::ld(grow + gc) : 0.f;
            As[lm][lc + j] = (aok && ac < p.AC) ? Cvt<T>::ld(arow + ac) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Gs[k][tr + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = As[k][tc + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += (TAcc)a[i] * (TAcc)b[j];
        }
        __syncthreads();
    }
    float* out = p.dWp + (int64_t)p.taps.widx[tap] * p.GC * p.AC;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gc = gc0 + tr + i;
        if (gc >= p.GC) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ac = ac0 + tc + j;
            if (ac < p.AC) {
                if (nsplit == 1) out[(int64_t)gc * p.AC + ac] = (float)acc[i][j];
                else atomicAdd(out + (int64_t)gc * p.AC + ac, (float)acc[i][j]);
            }
        }
    }
}

// ---- "thin" wgrad: one side of the outer product has <= 4 channels (first layer: x has 1 or 3 channels;
// last layer: dY has 1 or 3).  The wide side (64.. channels) is streamed once from HBM, one channel per thread;
// the thin side's halo band is staged in shared memory and read as warp-wide broadcasts.
//   wide pixel (y,x), tap t  <->  thin pixel (y*s + dir*ty_t, x*s + dir*tx_t)
//   out[widx_t*so_t + c*so_w + g*so_th] += sum wide[n,y,x,c] * thin[n, ., ., g]
struct ThinWgrad {
    const void* wide;   // [n, hw, ww, Cw]
    const void* thin;   // [n, ht, wt, Ct]
    float* out;
    int n, hw, ww, Cw, ht, wt, Ct;
    int s, dir;
    int so_t, so_w, so_th;
    int tymin, tymax, txmin, txmax;   // of dir*ty, dir*tx
    TapList taps;
};

constexpr int THIN_ROWS = 4;       // wide rows per band
constexpr int THIN_MAXT = 25;

template <typename T, int CT>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const ThinWgrad p) {
    extern __shared__ float sh[];
    const T* __restrict__ wide = (const T*)p.wide;
    const T* __restrict__ thin = (const T*)p.thin;
    const int c = blockIdx.y * 64 + (threadIdx.x & 63);
    const int rr = threadIdx.x >> 6;                      // 0..3: row inside the band
    const int nt = p.taps.ntaps;
    const int band_rows = (THIN_ROWS - 1) * p.s + (p.tymax - p.tymin) + 1;
    const int band_cols = (p.ww - 1) * p.s + (p.txmax - p.txmin) + 1;
    const int bands_per_img = (p.hw + THIN_ROWS - 1) / THIN_ROWS;
    const int nbands = p.n * bands_per_img;
    float acc[THIN_MAXT][CT];
#pragma unroll
    for (int t = 0; t < THIN_MAXT; ++t)
#pragma unroll
        for (int g = 0; g < CT; ++g) acc[t][g] = 0.f;

    for (int band = blockIdx.x; band < nbands; band += gridDim.x) {
        const int img = band / bands_per_img;
        const int y0 = (band - img * bands_per_img) * THIN_ROWS;
        __syncthreads();
        // stage the thin halo band (zero outside the tensor)
        for (int i = threadIdx.x; i < band_rows * band_cols * CT; i += blockDim.x) {
            const int g = i % CT;
            const int col = (i / CT) % band_cols;
            const int row = i / (CT * band_cols);
            const int ty = y0 * p.s + p.tymin + row, tx = p.txmin + col;
            float v = 0.f;
            if (ty >= 0 && ty < p.ht && tx >= 0 && tx < p.wt)
                v = Cvt<T>::ld(thin + (((int64_t)img * p.ht + ty) * p.wt + tx) * p.Ct + g);
            sh[i] = v;
        }
        __syncthreads();
        const int y = y0 + rr;
        if (y < p.hw && c < p.Cw) {
            const T* wrow = wide + (((int64_t)img * p.hw + y) * p.ww) * p.Cw + c;
            for (int x = 0; x < p.ww; ++x) {
                const float w = Cvt<T>::ld(wrow + (int64_t)x * p.Cw);
#pragma unroll
                for (int t = 0; t < THIN_MAXT; ++t) {
                    if (t < nt) {
                        const int row = rr * p.s + p.dir * p.taps.ty[t] - p.tymin;
                        const int col = x * p.s + p.dir * p.taps.tx[t] - p.txmin;
                        const float* th = sh + (row * band_cols + col) * CT;
#pragma unroll
                        for (int g = 0; g < CT; ++g) acc[t][g] = fmaf(w, th[g], acc[t][g]);
                    }
                }
            }
        }
    }
    // combine the 4 row lanes, then one atomic per (tap, channel, thin channel) per block
    __syncthreads();
    float* red = sh;  // [4][64] per (t,g) pass
    for (int t = 0; t < nt; ++t)
        for (int g = 0; g < CT; ++g) {
            red[rr * 64 + (threadIdx.x & 63)] = acc[t][g];
            __syncthreads();
            if (rr == 0 && c < p.Cw) {
                const float v = red[threadIdx.x] + red[64 + threadIdx.x] + red[128 + threadIdx.x] + red[192 + threadIdx.x];
                atomicAdd(p.out + (int64_t)p.taps.widx[t] * p.so_t + (int64_t)c * p.so_w + (int64_t)g * p.so_th, v);
            }
            __syncthreads();
        }
}

// Dense KHxKW window specialisation (the 5x5 layers of models/networks.py): tap (ky,kx) of wide pixel x reads
// the staged band at row rr*s + ky', column x*s + kx' (ky' = ky for dir=+1, KH-1-ky for dir=-1), so the five
// row pointers live in registers and the kx offsets are immediates: one LDS (warp broadcast) + one FMA per tap.
template <typename T, int CT, int DIR, int KH, int KW>
__global__ void __launch_bounds__(256) thin_wgrad_dense_kernel(const ThinWgrad p) {
    extern __shared__ float sh[];
    const T* __restrict__ wide = (const T*)p.wide;
    const T* __restrict__ thin = (const T*)p.thin;
    const int c = blockIdx.y * 64 + (threadIdx.x & 63);
    const int rr = threadIdx.x >> 6;
    const int band_rows = (THIN_ROWS - 1) * p.s + KH;
    const int band_cols = (p.ww - 1) * p.s + KW;
    const int bands_per_img = (p.hw + THIN_ROWS - 1) / THIN_ROWS;
    const int nbands = p.n * bands_per_img;
    float acc[KH][KW][CT];
#pragma unroll
    for (int a = 0; a < KH; ++a)
#pragma unroll
        for (int b = 0; b < KW; ++b)
#pragma unroll
            for (int g = 0; g < CT; ++g) acc[a][b][g] = 0.f;
    const float* rowp[KH];
#pragma unroll
    for (int ky = 0; ky < KH; ++ky) rowp[ky] = sh + (rr * p.s + (DIR > 0 ? ky : KH - 1 - ky)) * band_cols * CT;
    const int xstep = p.s * CT;

    for (int band = blockIdx.x; band < nbands; band += gridDim.x) {
        const int img = band / bands_per_img;
        const int y0 = (band - img * bands_per_img) * THIN_ROWS;
        __syncthreads();
        for (int i = threadIdx.x; i < band_rows * band_cols * CT; i += blockDim.x) {
            const int g = i % CT;
            const int col = (i / CT) % band_cols;
            const int row = i / (CT * band_cols);
            const int ty = y0 * p.s + p.tymin + row, tx = p.txmin + col;
            float v = 0.f;
            if (ty >= 0 && ty < p.ht && tx >= 0 && tx < p.wt)
                v = Cvt<T>::ld(thin + (((int64_t)img * p.ht + ty) * p.wt + tx) * p.Ct + g);
            sh[i] = v;
        }
        __syncthreads();
        const int y = y0 + rr;
        if (y < p.hw && c < p.Cw) {
            const T* wrow = wide + (((int64_t)img * p.hw + y) * p.ww) * p.Cw + c;
#pragma unroll 4
            for (int x = 0; x < p.ww; ++x) {
                const float w = Cvt<T>::ld(wrow + (int64_t)x * p.Cw);
#pragma unroll
                for (int ky = 0; ky < KH; ++ky) {
                    const float* b = rowp[ky] + x * xstep;
#pragma unroll
                    for (int kx = 0; kx < KW; ++kx)
#pragma unroll
                        for (int g = 0; g < CT; ++g)
                            acc[ky][kx][g] = fmaf(w, b[(DIR > 0 ? kx : KW - 1 - kx) * CT + g], acc[ky][kx][g]);
                }
            }
        }
    }
    __syncthreads();
    float* red = sh;
#pragma unroll
    for (int ky = 0; ky < KH; ++ky)
#pragma unroll
        for (int kx = 0; kx < KW; ++kx)
#pragma unroll
            for (int g = 0; g < CT; ++g) {
                red[rr * 64 + (threadIdx.x & 63)] = acc[ky][kx][g];
                __syncthreads();
                if (rr == 0 && c < p.Cw) {
                    const float v = red[threadIdx.x] + red[64 + threadIdx.x] + red[128 + threadIdx.x] + red[192 + threadIdx.x];
                    atomicAdd(p.out + (int64_t)p.taps.widx[ky * KW + kx] * p.so_t + (int64_t)c * p.so_w + (int64_t)g * p.so_th, v);
                }
                __syncthreads();
            }
}

bool taps_dense(const TapList& t, int kh, int kw) {
    if (t.ntaps != kh * kw) return false;
    for (int i = 0; i < t.ntaps; ++i)
        if (t.ty[i] != t.ty[0] + i / kw || t.tx[i] != t.tx[0] + i % kw) return false;
    return true;
}

template <typename T>
int launch_thin(const ThinWgrad& tp, cudaStream_t s) {
    const int band_rows = (THIN_ROWS - 1) * tp.s + (tp.tymax - tp.tymin) + 1;
    const int band_cols = (tp.ww - 1) * tp.s + (tp.txmax - tp.txmin) + 1;
    size_t smem = sizeof(float) * (size_t)band_rows * band_cols * tp.Ct;
    if (smem < 256 * sizeof(float)) smem = 256 * sizeof(float);
    if (smem > 48 * 1024) return VP_EUNSUPPORTED;
    const int nbands = tp.n * ((tp.hw + THIN_ROWS - 1) / THIN_ROWS);
    dim3 grid((unsigned)(nbands < 148 * 8 ? nbands : 148 * 8), (unsigned)((tp.Cw + 63) / 64));
    if (taps_dense(tp.taps, 5, 5) && tp.Ct <= 3) {
#define VP_THIN_DENSE(CT, DIR) thin_wgrad_dense_kernel<T, CT, DIR, 5, 5><<<grid, 256, smem, s>>>(tp)
        if (tp.dir > 0) { if (tp.Ct == 1) VP_THIN_DENSE(1, 1); else if (tp.Ct == 2) VP_THIN_DENSE(2, 1); else VP_THIN_DENSE(3, 1); }
        else { if (tp.Ct == 1) VP_THIN_DENSE(1, -1); else if (tp.Ct == 2) VP_THIN_DENSE(2, -1); else VP_THIN_DENSE(3, -1); }
#undef VP_THIN_DENSE
        VP_CHECK_LAUNCH("thin_wgrad_dense");
        return VP_OK;
    }
    switch (tp.Ct) {
        case 1: thin_wgrad_kernel<T, 1><<<grid, 256, smem, s>>>(tp); break;
        case 2: thin_wgrad_kernel<T, 2><<<grid, 256, smem, s>>>(tp); break;
        case 3: thin_wgrad_kernel<T, 3><<<grid, 256, smem, s>>>(tp); break;
        default: thin_wgrad_kernel<T, 4><<<grid, 256, smem, s>>>(tp); break;
    }
    VP_CHECK_LAUNCH("thin_wgrad");
    return VP_OK;
}

// returns VP_EUNSUPPORTED when the problem is not of the thin form
int try_thin_wgrad(const TapWgrad& p, cudaStream_t s) {
    if (p.taps.ntaps > THIN_MAXT) return VP_EUNSUPPORTED;
    ThinWgrad tp;
    memset(&tp, 0, sizeof(tp));
    tp.n = p.n; tp.out = p.dWp; tp.taps = p.taps;
    int dir;
    if (p.AC <= 4 && p.GC >= 16) {          // first-layer form: wide = G (grid side), thin = A
        tp.wide = p.G; tp.hw = p.gh; tp.ww = p.gw; tp.Cw = p.GC;
        tp.thin = p.A; tp.ht = p.ha; tp.wt = p.wa; tp.Ct = p.AC;
        tp.s = p.as; dir = 1;
        tp.so_t = p.GC * p.AC; tp.so_w = p.AC; tp.so_th = 1;
    } else if (p.GC <= 4 && p.AC >= 16 && p.as == 1) {   // last-layer form: wide = A, thin = G
        tp.wide = p.A; tp.hw = p.ha; tp.ww = p.wa; tp.Cw = p.AC;
        tp.thin = p.G; tp.ht = p.gh; tp.wt = p.gw; tp.Ct = p.GC;
        tp.s = 1; dir = -1;
        tp.so_t = p.GC * p.AC; tp.so_w = 1; tp.so_th = p.AC;
    } else {
        return VP_EUNSUPPORTED;
    }
    tp.dir = dir;
    tp.tymin = tp.txmin = 1 << 20; tp.tymax = tp.txmax = -(1 << 20);
    for (int t = 0; t < p.taps.ntaps; ++t) {
        const int a = dir * p.taps.ty[t], b = dir * p.taps.tx[t];
        tp.tymin = a < tp.tymin ? a : tp.tymin; tp.tymax = a > tp.tymax ? a : tp.tymax;
        tp.txmin = b < tp.txmin ? b : tp.txmin; tp.txmax = b > tp.txmax ? b : tp.txmax;
    }
    return launch_thin<bf16>(tp, s);
}

// ---- thin-K tap GEMM: the input has <= 4 channels (first-layer forward, last-layer data gradient) ------
// One thread = one output pixel x 8 output channels (a 16-byte store); the weights [tap][k][N] sit in shared
// memory as fp32, the few input scalars per tap are warp-broadcast loads.  Purely store-bound.
constexpr int THIN_PX = 4;   // output pixels (consecutive gx) per thread: each weight fetch from smem feeds 4 pixels

template <typename T, typename TD, int CT>
__global__ void __launch_bounds__(256) thin_fwd_kernel(const TapGemm p) {
    extern __shared__ float wsm[];   // [ntaps][CT][N]
    const T* __restrict__ A = (const T*)p.A;
    const T* __restrict__ W = (const T*)p.Wp;
    const int nt = p.taps.ntaps;
    for (int i = threadIdx.x; i < nt * CT * p.N; i += blockDim.x) {
        const int n = i % p.N;
        const int k = (i / p.N) % CT;
        const int t = i / (p.N * CT);
        wsm[i] = Cvt<T>::ld(W + ((int64_t)p.taps.widx[t] * p.N + n) * p.K + k);
    }
    __syncthreads();
    const int groups = p.N >> 3;
    const int xq = (p.gw + THIN_PX - 1) / THIN_PX;
    const int64_t total = (int64_t)p.n * p.gh * xq * groups;
    TD* __restrict__ D = (TD*)p.D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int g8 = (int)(i % groups) * 8;
        int64_t m = i / groups;
        const int gx0 = (int)(m % xq) * THIN_PX; m /= xq;
        const int gy = (int)(m % p.gh);
        const int n = (int)(m / p.gh);
        const int oy = gy * p.ds + p.doy;
        if (oy >= p.hd) continue;
        float acc[THIN_PX][8];
#pragma unroll
        for (int q = 0; q < THIN_PX; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[q][j] = p.bias ? p.bias[g8 + j] : 0.f;
        for (int t = 0; t < nt; ++t) {
            const int iy = gy * p.as + p.taps.ty[t];
            if (iy < 0 || iy >= p.ha) continue;
            const T* arow = A + ((int64_t)n * p.ha + iy) * p.wa * p.K;
#pragma unroll
            for (int k = 0; k < CT; ++k) {
                float a[THIN_PX];
#pragma unroll
                for (int q = 0; q < THIN_PX; ++q) {
                    const int ix = (gx0 + q) * p.as + p.taps.tx[t];
                    a[q] = (ix >= 0 && ix < p.wa) ? Cvt<T>::ld(arow + (int64_t)ix * p.K + k) : 0.f;
                }
                const float4 w0 = *reinterpret_cast<const float4*>(wsm + (t * CT + k) * p.N + g8);
                const float4 w1 = *reinterpret_cast<const float4*>(wsm + (t * CT + k) * p.N + g8 + 4);
#pragma unroll
                for (int q = 0; q < THIN_PX; ++q) {
                    acc[q][0] = fmaf(a[q], w0.x, acc[q][0]); acc[q][1] = fmaf(a[q], w0.y, acc[q][1]);
                    acc[q][2] = fmaf(a[q], w0.z, acc[q][2]); acc[q][3] = fmaf(a[q], w0.w, acc[q][3]);
                    acc[q][4] = fmaf(a[q], w1.x, acc[q][4]); acc[q][5] = fmaf(a[q], w1.y, acc[q][5]);
                    acc[q][6] = fmaf(a[q], w1.z, acc[q][6]); acc[q][7] = fmaf(a[q], w1.w, acc[q][7]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < THIN_PX; ++q) {
            const int gx = gx0 + q;
            const int ox = gx * p.ds + p.dox;
            if (gx >= p.gw || ox >= p.wd) continue;
            TD* out = D + (((int64_t)n * p.hd + oy) * p.wd + ox) * p.N + g8;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = act_fwd(acc[q][j], p.act, p.slope);
            if (sizeof(TD) == 2) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(out) = pk;
            } else {
                float* o = reinterpret_cast<float*>(out);
                *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
            }
        }
    }
}

int try_thin_fwd(const TapGemm& p, cudaStream_t s) {
    if (p.K > 4 || p.K < 1 || (p.N & 7) != 0 || p.N > 256 || ((uintptr_t)p.D & 15)) return VP_EUNSUPPORTED;
    const size_t smem = sizeof(float) * (size_t)p.taps.ntaps * p.K * p.N;
    if (smem > 48 * 1024 || smem == 0) return VP_EUNSUPPORTED;
    const int64_t total = (int64_t)p.n * p.gh * ((p.gw + THIN_PX - 1) / THIN_PX) * (p.N >> 3);
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const bool f32 = p.out_dtype == VP_F32;
#define VP_THIN_FWD(CT)                                                                              \
    if (f32) thin_fwd_kernel<bf16, float, CT><<<(unsigned)blocks, 256, smem, s>>>(p);                 \
    else thin_fwd_kernel<bf16, bf16, CT><<<(unsigned)blocks, 256, smem, s>>>(p)
    switch (p.K) {
        case 1: VP_THIN_FWD(1); break;
        case 2: VP_THIN_FWD(2); break;
        case 3: VP_THIN_FWD(3); break;
        default: VP_THIN_FWD(4); break;
    }
#undef VP_THIN_FWD
    VP_CHECK_LAUNCH("thin_fwd");
    return VP_OK;
}

}  // namespace

int launch_tapgemm_simt(const TapGemm& p, int dtype, cudaStream_t s) {
    const int64_t M = (int64_t)p.n * p.gh * p.gw;
    if (M == 0 || p.N == 0) return VP_OK;
    if (dtype == VP_BF16) count_simt_bf16();
    if (dtype == VP_BF16) {
        const int rc = try_thin_fwd(p, s);
        if (rc != VP_EUNSUPPORTED) return rc;
    }
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN));
    if (dtype == VP_F32 && p.out_dtype == VP_F32) tapgemm_simt_kernel<float, float><<<grid, NT, 0, s>>>(p);
    else if (dtype == VP_F32) tapgemm_simt_kernel<float, bf16><<<grid, NT, 0, s>>>(p);
    else if (p.out_dtype == VP_F32) tapgemm_simt_kernel<bf16, float><<<grid, NT, 0, s>>>(p);
    else tapgemm_simt_kernel<bf16, bf16><<<grid, NT, 0, s>>>(p);
    VP_CHECK_LAUNCH("tapgemm_simt");
    return VP_OK;
}

int launch_tapwgrad_simt(const TapWgrad& p, int dtype, cudaStream_t s) {
    const int64_t M = (int64_t)p.n * p.gh * p.gw;
    if (!p.accumulate) zero_async(p.dWp, sizeof(float) * (size_t)p.taps.ntaps * p.GC * p.AC, s);
    if (M == 0 || p.GC == 0 || p.AC == 0) return VP_OK;
    if (dtype == VP_BF16) count_simt_bf16();
    if (dtype == VP_BF16) {   // (the fp32 check mode keeps the double-accumulating generic kernel)
        const int rc = try_thin_wgrad(p, s);
        if (rc != VP_EUNSUPPORTED) return rc;
    }
    const int tiles = ((p.GC + BM - 1) / BM) * ((p.AC + BN - 1) / BN) * p.taps.ntaps;
    // split the pixel reduction so that a few waves of CTAs are in flight (148 SMs)
    // fp32 check mode: one double-precision accumulation chain per output (deterministic, single rounding);
    // the gradient of a conv that feeds BatchNorm is a heavily cancelling sum, fp32 chains lose ~1e-3 there.
    int nsplit = 1;
    const int64_t max_split = (M + 255) / 256;
    if (dtype != VP_F32)
        while ((int64_t)tiles * nsplit < 148 * 4 && nsplit * 2 <= max_split) nsplit *= 2;
    dim3 grid((p.GC + BM - 1) / BM, (p.AC + BN - 1) / BN, p.taps.ntaps * nsplit);
    if (dtype == VP_F32) tapwgrad_simt_kernel<float, double><<<grid, NT, 0, s>>>(p, nsplit);
    else tapwgrad_simt_kernel<bf16, float><<<grid, NT, 0, s>>>(p, nsplit);
    VP_CHECK_LAUNCH("tapwgrad_simt");
    return VP_OK;
}

}  // namespace vp
