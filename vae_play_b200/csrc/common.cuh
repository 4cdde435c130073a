// Shared helpers for libvaeplay_b200 (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <utility>

#include "../../include/vaeplay_b200.h"

namespace vp {

typedef __nv_bfloat16 bf16;

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void count_simt_bf16();      // a bf16 contraction that fell back to the CUDA-core engine (vp_simt_bf16_count: tests assert 0)

// ---- programmatic dependent launch ------------------------------------------------------------------------------------
// Every hot-path kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may be scheduled while
// the previous kernel of the stream is still draining, run their prologue (shared-memory carve-up, mbarrier init, TMEM
// allocation, tensor-map prefetch) and then block in pdl_wait() until the previous grid has completed and flushed.  Inside a
// captured CUDA graph the launches become programmatic dependency edges.  RULE: a kernel launched through launch_k() must
// execute pdl_wait() in every thread before its first global-memory access (reads of what the predecessor wrote AND writes
// to what it may still read); pdl_trigger() right after it lets the successor pre-launch in turn (at most one kernel ahead).
// EXCEPTION: the fp32 master weights.  They are written by the optimiser kernels only, which never trigger early: without
// griddepcontrol.launch_dependents the dependent grid is scheduled when the optimiser grid has completed, as in plain stream
// order.  The thin-layer kernels use this to build their weight tile (a strided fp32 gather, ~5 us) ahead of pdl_wait().
// VP_PDL=0 in the environment launches everything fully serialised (A/B measurements, debugging).
bool pdl_enabled();
// Dynamic shared memory to request so that AT MOST `ctas_per_sm` CTAs of a kernel fit on an SM.  A persistent kernel sized as
// "one CTA per SM" must also be unable to double up: CTAs that are scheduled while the previous kernel is still draining
// (programmatic dependent launch) land on whichever SMs free up first, and two of them on one SM (while another SM gets
// none) would double the kernel's run time -- or serialise on the SM's 512 TMEM columns.
inline int smem_for_occupancy(int bytes, int ctas_per_sm) {
    const int floor_bytes = 233472 / (ctas_per_sm + 1) - 1024 + 1;       // 228 KB per SM, 1 KB reserved per CTA
    return bytes > floor_bytes ? bytes : floor_bytes;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef VP_PDL_MODE
#define VP_PDL_MODE 1      // 1: wait + early trigger; 2: wait only (dependents launch when this grid's CTAs exit)
#endif
__device__ __forceinline__ void pdl_trigger() {
#if VP_PDL_MODE == 1
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_trigger(); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at;
    memset(&at, 0, sizeof(at));
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    // Only while the stream is being CAPTURED into a CUDA graph: there a programmatic edge can only come from a kernel node, so a
    // memset / copy node in front of this kernel stays a full dependency.  On a live stream the attribute also relaxes the order
    // against a preceding cudaMemsetAsync (measured: a padded-bias torch.zeros + cast pair lost the race) -- and eager launches
    // are bound by the host anyway.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cfg.numAttrs = (pdl_enabled() && cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Stream-ordered zero fill by one of OUR kernels (PDL-aware, a kernel node in captured graphs) instead of cudaMemsetAsync:
// keeps the programmatic chain of the step graph unbroken and every producer / consumer of the buffer inside one ordering domain.
int zero_async(void* ptr, size_t bytes, cudaStream_t s);

#define VP_CHECK_ARG(cond, ...)                  \
    do {                                         \
        if (!(cond)) {                           \
            vp::set_error(__VA_ARGS__);          \
            return VP_EINVAL;                    \
        }                                        \
    } while (0)

#define VP_CHECK_LAUNCH(name)                                                        \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            vp::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__)); \
            return VP_ECUDA;                                                         \
        }                                                                            \
        vp::count_launch();                                                          \
    } while (0)

template <typename T> struct Cvt;
template <> struct Cvt<float> {
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Cvt<bf16> {
    static __device__ __forceinline__ float ld(const bf16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
};

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
    switch (act) {
        case VP_ACT_RELU: return v > 0.f ? v : 0.f;
        case VP_ACT_LRELU: return v > 0.f ? v : slope * v;
        case VP_ACT_TANH: return tanhf(v);
        case VP_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}
// derivative of act at pre-activation v; with (act & VP_ACT_FROM_OUTPUT) v is the post-activation value
__device__ __forceinline__ float act_grad(float v, int act, float slope) {
    if (act & VP_ACT_FROM_OUTPUT) {
        switch (act & 15) {
            case VP_ACT_RELU: return v > 0.f ? 1.f : 0.f;
            case VP_ACT_LRELU: return v > 0.f ? 1.f : slope;
            case VP_ACT_TANH: return 1.f - v * v;
            case VP_ACT_SIGMOID: return v * (1.f - v);
            default: return 1.f;
        }
    }
    switch (act) {
        case VP_ACT_RELU: return v > 0.f ? 1.f : 0.f;
        case VP_ACT_LRELU: return v > 0.f ? 1.f : slope;
        case VP_ACT_TANH: { float t = tanhf(v); return 1.f - t * t; }
        case VP_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-v)); return s * (1.f - s); }
        default: return 1.f;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------------
// The generic "tap GEMM": every Conv2d / ConvTranspose2d / Linear forward and data-gradient is
//   D[n, gy*ds + doy, gx*ds + dox, :] = sum_t  A[n, gy*as + ty_t, gx*as + tx_t, :] . Wp[widx_t][:, :]
// over an iteration grid (gy,gx) in [0,gh)x[0,gw), one launch per output phase.
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxTaps = 64;

struct TapList {
    int ntaps;
    int8_t ty[kMaxTaps];
    int8_t tx[kMaxTaps];
    int8_t widx[kMaxTaps];
};

struct TapGemm {
    const void* A;   // [n, ha, wa, K]
    const void* Wp;  // [taps_total][N][K]
    void* D;         // [n, hd, wd, N]
    const float* bias;  // [N] or null
    int n, ha, wa, K;
    int hd, wd, N;
    int gh, gw;     // iteration grid
    int as;         // stride applied to grid coords when reading A
    int ds, doy, dox;  // stride / offset applied to grid coords when writing D
    int act;
    float slope;
    int out_dtype;  // dtype of D (VP_F32 / VP_BF16); A and Wp are in the call's dtype
    // weight addressing (tcgen05 engine): element (tap, n, k) at Wp[tap*w_st + n*w_sn + k*w_sk].  Packed panels: w_st = N*K,
    // w_sn = K, w_sk = 1.  One of w_sn / w_sk must be 1: w_sk == 1 is a K-major operand, w_sn == 1 an MN-major one
    // (the module's own channels-last weight read without any re-packing).
    int64_t w_st, w_sn, w_sk;
    // optional (tcgen05 engine, bf16 output, N % 64 == 0, N <= 512, no bias / activation): per-CTA partial column sums and
    // sums of squares of the stored output, stat_parts[cta][2][N] fp32; *stat_nparts receives the number of CTAs.
    float* stat_parts;
    int stat_capacity;
    int* stat_nparts;
    // optional fused BatchNorm-backward reduction (tcgen05 engine, the data gradient that FOLLOWS a conv -> BatchNorm -> ReLU block):
    // D is dL/da of that block, bn_y its pre-norm conv output [n, hd, wd, N] bf16.  The epilogue forms d = D * (y*scale + shift > 0)
    // and accumulates sum d and sum d*(y - mean) per channel into stat_parts[cta][2][N] (vp_norm_bwd_finish_parts adds them up):
    // the separate reduce pass over (y, da) disappears.  D itself is stored unmasked.
    const void* bn_y;
    const float *bn_scale, *bn_shift, *bn_mean;
    TapList taps;
};

// wgrad:  dWp[widx_t][gc][ac] += sum_{n,gy,gx} G[n,gy,gx,gc] * A[n, gy*as+ty_t, gx*as+tx_t, ac]
struct TapWgrad {
    const void* G;  // [n, gh, gw, GC]
    const void* A;  // [n, ha, wa, AC]
    float* dWp;     // [taps][GC][AC] (packed) or any layout with AC contiguous: element (tap, gc, ac) at tap*o_st + gc*o_sg + ac
    int64_t o_st, o_sg;
    int accumulate; // 0: the launcher clears dWp first when it reduces with atomics; 1: dWp is known to hold zeros already
    int n, gh, gw, GC;
    int ha, wa, AC;
    int as;
    TapList taps;
};

int launch_tapgemm_simt(const TapGemm& p, int dtype, cudaStream_t s);
int launch_tapwgrad_simt(const TapWgrad& p, int dtype, cudaStream_t s);
// tcgen05 engine: returns VP_EUNSUPPORTED when the shape is not eligible
int launch_tapgemm_tc(const TapGemm& p, cudaStream_t s);
// all output-parity phases of one layer in a single persistent launch (phases share everything but grid/offset/taps)
int launch_tapgemm_tc_multi(const TapGemm* phases, int nphases, cudaStream_t s);
int launch_tapwgrad_tc(const TapWgrad& p, cudaStream_t s);
// tap-row weight gradient (G and the A halo shared by all taps of a filter row); VP_EUNSUPPORTED when not of that form
int launch_tapwgrad_win(const TapWgrad& p, int kh, int kw, int pad, cudaStream_t s);
// stride-1 tap sets with the activation halo re-used across taps; VP_EUNSUPPORTED when not of that form
int launch_tapgemm_win(const TapGemm* phases, int nphases, cudaStream_t s);
// two output-parity phases per tile, shared shifts as one N = 128 MMA (64-channel stride-2 layers); VP_EUNSUPPORTED otherwise
int launch_tapgemm_pair(const TapGemm* phases, int nphases, cudaStream_t s);
// stride-2 gather form with per-parity sub-lattice halos shared by the taps of a class; VP_EUNSUPPORTED otherwise
int launch_tapgemm_gwin(const TapGemm& p, cudaStream_t s);
bool tc_available();
int num_sms();            // SMs persistent grids are sized for (device count, or the vp_set_sm_limit cap)
void set_sm_limit(int n);

}  // namespace vp
