// extern "C" entry points of libvaeplay_b200: geometry -> tap-GEMM problems -> engine dispatch.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace vp {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
static std::atomic<uint64_t> g_simt_bf16{0};
void count_simt_bf16() { g_simt_bf16.fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VP_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

namespace {

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

bool check_geom(const VpConvGeom* g, const char* who) {
    if (!g) { set_error("%s: null geometry", who); return false; }
    if (g->n <= 0 || g->hi <= 0 || g->wi <= 0 || g->ci <= 0 || g->ho <= 0 || g->wo <= 0 || g->co <= 0 || g->kh <= 0 ||
        g->kw <= 0 || g->stride <= 0 || g->pad < 0) {
        set_error("%s: non-positive dimension in geometry", who);
        return false;
    }
    if (g->kh * g->kw > kMaxTaps) { set_error("%s: kernel %dx%d exceeds %d taps", who, g->kh, g->kw, kMaxTaps); return false; }
    if (g->kh > 100 || g->pad > 100) { set_error("%s: kernel/pad too large", who); return false; }
    if (!g->transposed) {
        const int ho = (g->hi + 2 * g->pad - g->kh) / g->stride + 1, wo = (g->wi + 2 * g->pad - g->kw) / g->stride + 1;
        if (ho != g->ho || wo != g->wo) {
            set_error("%s: conv output %dx%d inconsistent with input %dx%d k%d s%d p%d (expected %dx%d)", who, g->ho, g->wo,
                      g->hi, g->wi, g->kh, g->stride, g->pad, ho, wo);
            return false;
        }
    } else {
        // ho = (hi-1)*s - 2p + k + output_padding, 0 <= output_padding < s
        const int base_h = (g->hi - 1) * g->stride - 2 * g->pad + g->kh, base_w = (g->wi - 1) * g->stride - 2 * g->pad + g->kw;
        if (g->ho < base_h || g->ho >= base_h + g->stride || g->wo < base_w || g->wo >= base_w + g->stride) {
            set_error("%s: transposed-conv output %dx%d inconsistent with input %dx%d k%d s%d p%d", who, g->ho, g->wo, g->hi,
                      g->wi, g->kh, g->stride, g->pad);
            return false;
        }
    }
    return true;
}

// "gather" form: one phase, grid over the small side, A read with stride s.
//   conv fwd:  grid (ho,wo), A = x;   convT dgrad: grid (hi,wi), A = dy.
void gather_taps(const VpConvGeom& g, TapList& t) {
    t.ntaps = 0;
    for (int ky = 0; ky < g.kh; ++ky)
        for (int kx = 0; kx < g.kw; ++kx) {
            t.ty[t.ntaps] = (int8_t)(ky - g.pad);
            t.tx[t.ntaps] = (int8_t)(kx - g.pad);
            t.widx[t.ntaps] = (int8_t)(ky * g.kw + kx);
            ++t.ntaps;
        }
}

// "scatter" form, phase (py,px) of the large side: Y = s*gy + py receives tap ky = r0 + s*j from
// small-side row gy + c0 - j, with r0 = (py+p) % s, c0 = (py+p-r0)/s.
//   convT fwd: large = y, small = x;   conv dgrad: large = dx, small = dy.
void scatter_taps(const VpConvGeom& g, int py, int px, TapList& t) {
    const int s = g.stride;
    const int r0y = (py + g.pad) % s, c0y = (py + g.pad - r0y) / s;
    const int r0x = (px + g.pad) % s, c0x = (px + g.pad - r0x) / s;
    t.ntaps = 0;
    for (int jy = 0; r0y + s * jy < g.kh; ++jy)
        for (int jx = 0; r0x + s * jx < g.kw; ++jx) {
            t.ty[t.ntaps] = (int8_t)(c0y - jy);
            t.tx[t.ntaps] = (int8_t)(c0x - jx);
            t.widx[t.ntaps] = (int8_t)((r0y + s * jy) * g.kw + (r0x + s * jx));
            ++t.ntaps;
        }
}

int run_tapgemm(const TapGemm& p, int dtype, int engine, cudaStream_t s) {
    if (engine != VP_ENGINE_SIMT && dtype == VP_BF16) {
        const int rc = launch_tapgemm_tc(p, s);
        if (rc != VP_EUNSUPPORTED) return rc;
        if (engine == VP_ENGINE_TC) return rc;
    } else if (engine == VP_ENGINE_TC) {
        set_error("tensor-core engine needs dtype bf16");
        return VP_EUNSUPPORTED;
    }
    return launch_tapgemm_simt(p, dtype, s);
}

// forward == true : small side is the input (A), large side the output (conv fwd / convT dgrad -> gather;
//                                                                       convT fwd / conv dgrad -> scatter)
// wlay: 0 = packed panels Wp[tap][N][K]; 1 = the module's weight in channels-last order [d0][kh][kw][d1] read in place, with
// (N, K) = (d0, d1) -> K-major operand; 2 = same with (N, K) = (d1, d0) -> MN-major operand
struct StatOut { float* parts; int capacity; int* nparts; };
struct BnRed { const void* y; const float *scale, *shift, *mean; };

int conv_like(const VpConvGeom& g, bool gather, const void* A, int ha, int wa, int K, void* D, int hd, int wd, int N,
              const void* wp, const float* bias, int act, float slope, int dtype, int out_dtype, int engine, cudaStream_t s, int wlay = 0,
              const StatOut* st = nullptr, const BnRed* bn = nullptr) {
    TapGemm p;
    memset(&p, 0, sizeof(p));
    if (st) { p.stat_parts = st->parts; p.stat_capacity = st->capacity; p.stat_nparts = st->nparts; }
    if (bn) {
        if (!gather) { set_error("fused BatchNorm-backward reduction: only gather-form data gradients"); return VP_EUNSUPPORTED; }
        p.bn_y = bn->y; p.bn_scale = bn->scale; p.bn_shift = bn->shift; p.bn_mean = bn->mean;
    }
    p.out_dtype = out_dtype;
    const int64_t T = (int64_t)g.kh * g.kw;
    if (wlay == 0) { p.w_st = (int64_t)N * K; p.w_sn = K; p.w_sk = 1; }
    else if (wlay == 1) { p.w_st = K; p.w_sn = T * K; p.w_sk = 1; }
    else { p.w_st = N; p.w_sn = 1; p.w_sk = T * N; }
    p.A = A; p.Wp = wp; p.D = D; p.bias = bias;
    p.n = g.n; p.ha = ha; p.wa = wa; p.K = K; p.hd = hd; p.wd = wd; p.N = N;
    p.act = act; p.slope = slope;
    if (gather) {
        p.gh = hd; p.gw = wd; p.as = g.stride; p.ds = 1; p.doy = 0; p.dox = 0;
        gather_taps(g, p.taps);
        return run_tapgemm(p, dtype, engine, s);
    }
    // scatter form: s*s output-parity phases; the tcgen05 engine takes up to 4 of them in one persistent launch
    TapGemm phases[4];
    int np = 0;
    const bool batchable = g.stride * g.stride <= 4 && engine != VP_ENGINE_SIMT && dtype == VP_BF16;
    for (int py = 0; py < g.stride; ++py)
        for (int px = 0; px < g.stride; ++px) {
            if (py >= hd || px >= wd) continue;
            p.gh = ceil_div(hd - py, g.stride);
            p.gw = ceil_div(wd - px, g.stride);
            p.as = 1; p.ds = g.stride; p.doy = py; p.dox = px;
            scatter_taps(g, py, px, p.taps);
            if (batchable) { phases[np++] = p; continue; }
            const int rc = run_tapgemm(p, dtype, engine, s);
            if (rc) return rc;
        }
    if (batchable && np > 0) {
        const int rc = launch_tapgemm_tc_multi(phases, np, s);
        if (rc != VP_EUNSUPPORTED || engine == VP_ENGINE_TC) return rc;
        for (int i = 0; i < np; ++i) {
            const int r2 = launch_tapgemm_simt(phases[i], dtype, s);
            if (r2) return r2;
        }
    }
    return VP_OK;
}

}  // namespace
}  // namespace vp

using namespace vp;

extern "C" const char* vp_last_error(void) { return g_err; }
extern "C" int vp_abi_version(void) { return VP_ABI_VERSION; }
extern "C" uint64_t vp_launch_count(void) { return g_launches.load(); }
extern "C" uint64_t vp_simt_bf16_count(void) { return g_simt_bf16.load(); }

namespace vp { void set_splitk_workspace(void* ptr, size_t bytes); }
/* Scratch memory for split-K partial sums (skinny contractions such as the fc layers): a device buffer owned by the caller,
 * used stream-ordered by every later call until replaced; without one those contractions run unsplit. */
extern "C" int vp_set_workspace(void* ptr, size_t bytes) {
    vp::set_splitk_workspace(ptr, ptr ? bytes : 0);
    return VP_OK;
}

/* Cap the number of SMs the persistent kernels size their grids for (0 = all): leaves SMs to a collective that runs next to
 * them (data-parallel gradient exchange overlapped with the rest of backward).  Host-side setting read at launch time; a
 * captured CUDA graph keeps the value it was captured with. */
extern "C" int vp_set_sm_limit(int n) {
    vp::set_sm_limit(n);
    return VP_OK;
}

extern "C" int vp_device_arch(void) {
    int dev = 0, maj = 0, min = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("cudaGetDevice failed"); return VP_ECUDA; }
    cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev);
    return maj * 10 + min;
}

extern "C" int vp_conv_fwd(const VpConvGeom* g, const void* x, const void* wp, const float* bias, void* y, int dtype,
                           int out_dtype, int act, float slope, int engine, void* stream) {
    if (!check_geom(g, "vp_conv_fwd")) return VP_EINVAL;
    VP_CHECK_ARG(x && wp && y, "vp_conv_fwd: null pointer");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_conv_fwd: bad dtype %d", dtype);
    VP_CHECK_ARG(out_dtype == VP_F32 || out_dtype == VP_BF16, "vp_conv_fwd: bad out_dtype %d", out_dtype);
    return conv_like(*g, !g->transposed, x, g->hi, g->wi, g->ci, y, g->ho, g->wo, g->co, wp, bias, act, slope, dtype, out_dtype,
                     engine, (cudaStream_t)stream);
}

extern "C" int vp_conv_dgrad(const VpConvGeom* g, const void* dy, const void* wp_t, void* dx, int dtype, int out_dtype,
                             int engine, void* stream) {
    if (!check_geom(g, "vp_conv_dgrad")) return VP_EINVAL;
    VP_CHECK_ARG(dy && wp_t && dx, "vp_conv_dgrad: null pointer");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_conv_dgrad: bad dtype %d", dtype);
    VP_CHECK_ARG(out_dtype == VP_F32 || out_dtype == VP_BF16, "vp_conv_dgrad: bad out_dtype %d", out_dtype);
    return conv_like(*g, g->transposed != 0, dy, g->ho, g->wo, g->co, dx, g->hi, g->wi, g->ci, wp_t, nullptr, VP_ACT_NONE, 0.f,
                     dtype, out_dtype, engine, (cudaStream_t)stream);
}

namespace {
void fill_wgrad(const VpConvGeom* g, const void* x, const void* dy, float* dw, bool channels_last, TapWgrad& p) {
    memset(&p, 0, sizeof(p));
    p.n = g->n; p.as = g->stride; p.dWp = dw;
    gather_taps(*g, p.taps);
    if (!g->transposed) {
        p.G = dy; p.gh = g->ho; p.gw = g->wo; p.GC = g->co;
        p.A = x; p.ha = g->hi; p.wa = g->wi; p.AC = g->ci;
    } else {
        p.G = x; p.gh = g->hi; p.gw = g->wi; p.GC = g->ci;
        p.A = dy; p.ha = g->ho; p.wa = g->wo; p.AC = g->co;
    }
    if (channels_last) { p.o_st = p.AC; p.o_sg = (int64_t)p.taps.ntaps * p.AC; }     // [GC][kh][kw][AC]
    else { p.o_st = (int64_t)p.GC * p.AC; p.o_sg = p.AC; }                            // [taps][GC][AC]
}
}  // namespace

/* ---- the same contractions with the weight read / the gradient written IN PLACE (no packed panels) ----------------
 * w_cl: the bf16 copy of the module's weight in channels-last element order: nn.Conv2d [co][kh][kw][ci], nn.ConvTranspose2d
 * [ci][kh][kw][co], nn.Linear [out][in].  dw_cl: fp32 gradient in the same element order (zeroed by the call). */
extern "C" int vp_conv_fwd_cl(const VpConvGeom* g, const void* x, const void* w_cl, const float* bias, void* y, int out_dtype, int act,
                              float slope, void* stream) {
    if (!check_geom(g, "vp_conv_fwd_cl")) return VP_EINVAL;
    VP_CHECK_ARG(x && w_cl && y, "vp_conv_fwd_cl: null pointer");
    return conv_like(*g, !g->transposed, x, g->hi, g->wi, g->ci, y, g->ho, g->wo, g->co, w_cl, bias, act, slope, VP_BF16, out_dtype,
                     VP_ENGINE_TC, (cudaStream_t)stream, g->transposed ? 2 : 1);
}

/* vp_conv_fwd_cl for a layer followed by BatchNorm (bf16 output, no bias / activation) that also returns the statistics of y:
 * stat_parts[*nparts][2][co] fp32 = per-CTA partial column sums and sums of squares (vp_norm_finalize_parts adds them up).
 * stat_capacity: parts the buffer can hold (2 * #SMs is always enough).  co must be a multiple of 64, <= 512. */
extern "C" int vp_conv_fwd_cl_stats(const VpConvGeom* g, const void* x, const void* w_cl, void* y, float* stat_parts, int stat_capacity,
                                    int* nparts, void* stream) {
    if (!check_geom(g, "vp_conv_fwd_cl_stats")) return VP_EINVAL;
    VP_CHECK_ARG(x && w_cl && y && stat_parts && nparts && stat_capacity > 0, "vp_conv_fwd_cl_stats: null pointer");
    StatOut st{stat_parts, stat_capacity, nparts};
    return conv_like(*g, !g->transposed, x, g->hi, g->wi, g->ci, y, g->ho, g->wo, g->co, w_cl, nullptr, VP_ACT_NONE, 0.f, VP_BF16, VP_BF16,
                     VP_ENGINE_TC, (cudaStream_t)stream, g->transposed ? 2 : 1, &st);
}

extern "C" int vp_conv_dgrad_cl(const VpConvGeom* g, const void* dy, const void* w_cl, void* dx, int out_dtype, void* stream) {
    if (!check_geom(g, "vp_conv_dgrad_cl")) return VP_EINVAL;
    VP_CHECK_ARG(dy && w_cl && dx, "vp_conv_dgrad_cl: null pointer");
    return conv_like(*g, g->transposed != 0, dy, g->ho, g->wo, g->co, dx, g->hi, g->wi, g->ci, w_cl, nullptr, VP_ACT_NONE, 0.f, VP_BF16,
                     out_dtype, VP_ENGINE_TC, (cudaStream_t)stream, g->transposed ? 1 : 2);
}

/* vp_conv_dgrad_cl of the layer that FOLLOWS a conv -> BatchNorm -> ReLU block, fused with the first pass of that block's
 * BatchNorm backward: dx = dL/da of the block is stored as usual, and per-CTA partial sums of d = dx * (y*scale + shift > 0) and of
 * d * (y - mean) are returned in parts[*nparts][2][ci] (vp_norm_bwd_finish_parts -> the `sums` vp_norm_bwd_apply consumes).
 * VP_EUNSUPPORTED when the shape is not served by a kernel with this epilogue (the caller runs the plain vp_conv_dgrad_cl and
 * vp_norm_bwd_reduce instead); nothing has been launched in that case. */
extern "C" int vp_conv_dgrad_cl_bnred(const VpConvGeom* g, const void* dy, const void* w_cl, void* dx, const void* y_prev, const float* scale,
                                      const float* shift, const float* mean, float* parts, int capacity, int* nparts, void* stream) {
    if (!check_geom(g, "vp_conv_dgrad_cl_bnred")) return VP_EINVAL;
    VP_CHECK_ARG(dy && w_cl && dx && y_prev && scale && shift && mean && parts && nparts && capacity > 0, "vp_conv_dgrad_cl_bnred: null pointer");
    StatOut st{parts, capacity, nparts};
    BnRed bn{y_prev, scale, shift, mean};
    return conv_like(*g, g->transposed != 0, dy, g->ho, g->wo, g->co, dx, g->hi, g->wi, g->ci, w_cl, nullptr, VP_ACT_NONE, 0.f, VP_BF16,
                     VP_BF16, VP_ENGINE_TC, (cudaStream_t)stream, g->transposed ? 1 : 2, &st, &bn);
}

extern "C" int vp_conv_wgrad_cl(const VpConvGeom* g, const void* x, const void* dy, float* dw_cl, int accumulate, void* stream) {
    if (!check_geom(g, "vp_conv_wgrad_cl")) return VP_EINVAL;
    VP_CHECK_ARG(x && dy && dw_cl, "vp_conv_wgrad_cl: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    TapWgrad p;
    fill_wgrad(g, x, dy, dw_cl, true, p);
    p.accumulate = accumulate ? 1 : 0;
    const int rc = launch_tapwgrad_win(p, g->kh, g->kw, g->pad, s);
    if (rc != VP_EUNSUPPORTED) return rc;
    return launch_tapwgrad_tc(p, s);
}

extern "C" int vp_conv_wgrad(const VpConvGeom* g, const void* x, const void* dy, float* dwp, int dtype, int engine,
                             void* stream) {
    if (!check_geom(g, "vp_conv_wgrad")) return VP_EINVAL;
    VP_CHECK_ARG(x && dy && dwp, "vp_conv_wgrad: null pointer");
    VP_CHECK_ARG(dtype == VP_F32 || dtype == VP_BF16, "vp_conv_wgrad: bad dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    TapWgrad p;
    fill_wgrad(g, x, dy, dwp, false, p);
    if (engine != VP_ENGINE_SIMT && dtype == VP_BF16) {
        int rc = launch_tapwgrad_win(p, g->kh, g->kw, g->pad, s);
        if (rc == VP_EUNSUPPORTED) rc = launch_tapwgrad_tc(p, s);
        if (rc != VP_EUNSUPPORTED || engine == VP_ENGINE_TC) return rc;
    } else if (engine == VP_ENGINE_TC) {
        set_error("tensor-core engine needs dtype bf16");
        return VP_EUNSUPPORTED;
    }
    return launch_tapwgrad_simt(p, dtype, s);
}
