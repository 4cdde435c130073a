// CUDA-core engine of the tap GEMM: fp32 check mode, and odd shapes (Cin=1/3, Cout=1/3, K % 64 != 0)
// in bf16 mode.  Classic 64x64x16 shared-memory tiling, 4x4 micro-tile per thread, fp32 accumulate.
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

}  // namespace

int launch_tapgemm_simt(const TapGemm& p, int dtype, cudaStream_t s) {
    const int64_t M = (int64_t)p.n * p.gh * p.gw;
    if (M == 0 || p.N == 0) return VP_OK;
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
    if (M == 0 || p.GC == 0 || p.AC == 0) return VP_OK;
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
