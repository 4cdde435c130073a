/*
 * vaeplay_b200.h -- C ABI of libvaeplay_b200.so, the B200 (sm_100a) kernel library under the
 * vae-play VAE training step.
 *
 * The reference (kungyao/vae-play) has no FFI of its own: its hot path is the nn.Module API of
 * models/blocks.py and models/networks.py, and every arithmetic call goes to torch.nn
 * (SURVEY.md section 8b).  The entry points below are what a binding for that path binds instead
 * of ATen/cuDNN/cuBLAS; each cites the reference call site it replaces.  INTEGRATION.md shows the
 * ctypes stub used by vae_play_b200/_lib.py.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*.
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*), never allocate,
 *     never synchronise, and are re-entrant per stream.
 *   - return value: 0 on success, a negative VP_E* code otherwise (vp_last_error() has the text).
 *   - activations are channels-last (NHWC) in `dtype` (VP_F32: fp32 check mode on CUDA cores;
 *     VP_BF16: bf16 storage, tcgen05 tensor-core contractions with fp32 accumulation).
 *     Parameters, statistics, losses and gradients of parameters are always fp32.
 *   - weights: three routes.  (1) IN PLACE (vp_conv_*_cl): the layer's own weight, kept in channels-last element
 *     order, is read by the TMA as a tcgen05 operand and its gradient written in the same order -- no packed copies
 *     (bf16, both channel counts multiples of 64).  (2) THIN (vp_thin_conv_*): single-channel layers read the fp32
 *     weight directly.  (3) PACKED (vp_conv_fwd/dgrad/wgrad): tap-major panels Wp[tap][N][K] built by vp_pack_weight
 *     (tap = ky*kw + kx), gradients in the same packed layout scattered back by vp_unpack_wgrad -- the fp32 check
 *     mode and every other shape.
 *   - one stream at a time per process for calls that use the workspace registered with vp_set_workspace.
 */
#ifndef VAEPLAY_B200_H
#define VAEPLAY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VP_ABI_VERSION 2

enum { VP_F32 = 0, VP_BF16 = 1 };
enum { VP_ACT_NONE = 0, VP_ACT_RELU = 1, VP_ACT_LRELU = 2, VP_ACT_TANH = 3, VP_ACT_SIGMOID = 4 };
/* backward only: OR into `act` when the tensor passed as x holds act(x) instead of the pre-activation */
#define VP_ACT_FROM_OUTPUT 16
enum { VP_OK = 0, VP_EINVAL = -1, VP_ECUDA = -2, VP_EUNSUPPORTED = -3 };
/* engine selection for the contractions: AUTO = tcgen05 when dtype is bf16 and the shape is
 * eligible, else the CUDA-core kernel; SIMT / TC force one (TC returns VP_EUNSUPPORTED if not eligible). */
enum { VP_ENGINE_AUTO = 0, VP_ENGINE_SIMT = 1, VP_ENGINE_TC = 2 };

/* Geometry of one Conv2d / ConvTranspose2d layer (nn.Conv2d at models/networks.py:14,101 and
 * models/blocks.py:10-17; nn.ConvTranspose2d at models/networks.py:38, models/network_Style_GAN.py:49).
 * `x` is always the layer's forward input [n,hi,wi,ci], `y` its forward output [n,ho,wo,co].
 * nn.Linear (models/networks.py:65-70,88; models/blocks.py:39) is the case kh=kw=1, hi=wi=ho=wo=1. */
typedef struct VpConvGeom {
    int32_t n, hi, wi, ci;
    int32_t ho, wo, co;
    int32_t kh, kw;
    int32_t stride, pad;
    int32_t transposed; /* 0: Conv2d / Linear, 1: ConvTranspose2d */
} VpConvGeom;

const char* vp_last_error(void);
int vp_abi_version(void);
/* compute capability major*10+minor of the current device, or a negative error code */
int vp_device_arch(void);

/* Scratch memory for split-K partial sums of skinny contractions (the fc layers: a few output tiles, thousands of K
 * iterations): a DEVICE buffer owned by the caller, registered for the CURRENT device (one buffer per device) and used
 * stream-ordered (one stream at a time per device) by every later call on that device until replaced; without one such
 * contractions run unsplit on a few SMs.  >= 4 * (rows * out_features) bytes to be useful. */
int vp_set_workspace(void* ptr, size_t bytes);

/* Cap the SMs the persistent kernels size their grids for (n <= 0: all of them).  Used while a collective (the
 * data-parallel gradient all-reduce, vae_play_b200/parallel.py) runs next to the rest of backward: a persistent grid must be
 * fully resident, so the SMs the collective's CTAs occupy are left out.  Read at launch time on the host. */
int vp_set_sm_limit(int n);

/* ---- weight layout ------------------------------------------------------------------------------- */
/* Wp[t][n][k] = (dtype) w[n*stride_n + k*stride_k + t*stride_t],  t in [0,taps).  `w` fp32 (torch layout).
 * (Conv2d: stride_t = 1.  The NCHW-flatten Linear layers of models/networks.py:65,88 are expressed as
 * 8x8-tap layers over the channels-last map by choosing the strides; see vae_play_b200/functional.py.) */
int vp_pack_weight(const float* w, void* wp, int dtype, int taps, int n, int k,
                   int64_t stride_n, int64_t stride_k, int64_t stride_t, void* stream);
/* dw[n*stride_n + k*stride_k + t*stride_t] = dwp[t][n][k]   (fp32 -> fp32, torch layout) */
int vp_unpack_wgrad(const float* dwp, float* dw, int taps, int n, int k,
                    int64_t stride_n, int64_t stride_k, int64_t stride_t, void* stream);

/* ---- contractions (implicit GEMM over filter taps) ------------------------------------------------- */
/* y = conv(x, w) (+bias) then optional pointwise activation (only meaningful when no norm follows).
 * wp: packed [taps][co][ci] in `dtype`; x in `dtype`; y in `out_dtype` (fp32 output of a bf16 contraction
 * is used for the mu/logvar heads).  Replaces nn.Conv2d / nn.ConvTranspose2d / nn.Linear forward. */
int vp_conv_fwd(const VpConvGeom* g, const void* x, const void* wp, const float* bias, void* y,
                int dtype, int out_dtype, int act, float slope, int engine, void* stream);
/* dx = dL/dx from dy.  wp_t: packed [taps][ci][co] in `dtype`.  Replaces the autograd dgrad. */
int vp_conv_dgrad(const VpConvGeom* g, const void* dy, const void* wp_t, void* dx,
                  int dtype, int out_dtype, int engine, void* stream);
/* dwp (fp32, packed: [taps][co][ci] for Conv2d/Linear, [taps][ci][co] for ConvTranspose2d) = dL/dw.
 * dwp is zeroed by the call (split-K partial sums are accumulated into it). */
int vp_conv_wgrad(const VpConvGeom* g, const void* x, const void* dy, float* dwp,
                  int dtype, int engine, void* stream);

/* ---- normalisation + activation over channels-last rows ------------------------------------------- */
/* x: [groups*rows_per_group, c].  BatchNorm2d/1d: groups=1 (models/networks.py:16,40,66,89;
 * blocks.py:21).  InstanceNorm2d: groups=n, rows_per_group=h*w (blocks.py:23).
 * sums: double [2][groups][c] scratch (zeroed by the call) -> per-channel sum and sum of squares. */
int vp_norm_stats(const void* x, double* sums, int dtype, int64_t groups, int64_t rows_per_group, int c,
                  void* stream);
/* From sums: mean/invstd (saved for backward, fp32 [groups][c]) and the fused affine
 * scale = gamma*invstd, shift = beta - mean*scale (fp32 [groups][c]; gamma/beta may be NULL = 1/0).
 * If running_mean/var are non-NULL they are blended in place with `momentum` (unbiased variance),
 * torch convention running = (1-momentum)*running + momentum*batch.  num_batches_tracked (nullable): the
 * BatchNorm module's int64 device counter, incremented by one (saves the separate torch `+= 1` launch). */
int vp_norm_finalize(const double* sums, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float momentum, float eps,
                     float* mean, float* invstd, float* scale, float* shift,
                     int64_t groups, int64_t rows_per_group, int c, int64_t* num_batches_tracked, void* stream);
/* a = act(x*scale + shift) (scale/shift may be NULL = identity; bias-free pointwise activation). */
int vp_norm_apply_act(const void* x, const float* scale, const float* shift, void* a, int dtype,
                      int64_t groups, int64_t rows_per_group, int c, int act, float slope, void* stream);
/* backward, pass 1: d = da * act'(x*scale+shift); sums (double [2][groups][c], zeroed by the call) get
 * sum(d) and sum(d*xhat), xhat = (x-mean)*invstd.  With mean == NULL (no norm) only sum(d) is
 * produced (bias gradient) and `dx` (if non-NULL) receives d.  */
int vp_norm_bwd_reduce(const void* x, const void* da, const float* mean, const float* invstd,
                       const float* scale, const float* shift, double* sums, void* dx_or_null, int dtype,
                       int64_t groups, int64_t rows_per_group, int c, int act, float slope, void* stream);
/* backward, pass 2: dx = scale * (d - sum(d)/m - xhat*sum(d*xhat)/m); dgamma = sum(d*xhat), dbeta = sum(d)
 * (fp32 [c], written when non-NULL and groups == 1). */
int vp_norm_bwd_apply(const void* x, const void* da, const float* mean, const float* invstd,
                      const float* scale, const float* shift, const double* sums,
                      void* dx, float* dgamma, float* dbeta, int dtype,
                      int64_t groups, int64_t rows_per_group, int c, int act, float slope, void* stream);
/* Train-mode BatchNorm (+ activation) over x [rows, c] with FEW rows (<= 8192) in one launch per direction -- the
 * BatchNorm1d behind the fc layers (models/networks.py:66,89): statistics + finalize + apply (forward), reduce + apply
 * (backward).  Same outputs as the vp_norm_* sequence; c a multiple of 4, tensors 16-byte aligned. */
int vp_bn_rows_fwd(const void* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                   float momentum, float eps, void* a, float* mean, float* invstd, float* scale, float* shift,
                   int dtype, int64_t rows, int c, int act, float slope, int64_t* num_batches_tracked, void* stream);
int vp_bn_rows_bwd(const void* x, const void* da, const float* mean, const float* invstd, const float* scale,
                   const float* shift, void* dx, float* dgamma, float* dbeta, int dtype, int64_t rows, int c,
                   int act, float slope, void* stream);
/* out[c] = sum over rows of x[rows,c] (bias gradients).  out fp32, written (not accumulated);
 * scratch_c: double [2][c] (zeroed by the call). */
int vp_colsum(const void* x, float* out, double* scratch_c, int dtype, int64_t rows, int c, void* stream);

/* ---- reparameterisation + KL (models/networks.py:228-231 and :270; train_Style_GAN.py:156-160,218) -- */
/* eps ~ N(0,1) from Philox4x32-10 with the exact element->counter mapping of Tensor.normal_() on
 * CUDA (ATen DistributionTemplates.h:50-91, curand_normal4), so that (seed, offset) taken from
 * torch's CUDA generator reproduce the reference's eps bit for bit.  n_total = rows*z.
 * offset_dev (nullable): device uint64 added to `offset` at run time (CUDA-graph replay).
 * If eps_in != NULL it is used instead of drawing.  mu/logvar: fp32 with row stride ld.
 * z: `dtype` [rows,z]; eps_out fp32 [rows,z] (nullable); kl fp32 [rows]. */
int vp_reparam_kl_fwd(const float* mu, const float* logvar, int64_t ld, const float* eps_in,
                      uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int num_sms,
                      void* z, int z_dtype, float* eps_out, float* kl, int64_t rows, int zdim, void* stream);
/* dmu = dz + dkl*mu ; dlogvar = dz*eps*0.5*exp(0.5*logvar) + dkl*0.5*(exp(logvar)-1).
 * dz (`dz_dtype`, nullable), dkl fp32 [rows] (nullable).  dmu/dlogvar in `out_dtype` with row stride ld_out. */
int vp_reparam_kl_bwd(const float* mu, const float* logvar, int64_t ld, const float* eps,
                      const void* dz, int dz_dtype, const float* dkl,
                      void* dmu, void* dlogvar, int out_dtype, int64_t ld_out, int64_t rows, int zdim, void* stream);
/* out[i] = N(0,1) sample i of the same stream (Tensor.normal_() drop-in; models/networks.py:241). */
int vp_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                     int num_sms, void* stream);
/* *offset_dev += inc  (advance a device-resident generator offset between graph replays) */
int vp_philox_advance(uint64_t* offset_dev, uint64_t inc, void* stream);

/* ---- reconstruction losses (train.py:62, train_Style_GAN.py:220, train_BE.py:58-59, tools/ops.py:12-19) */
/* kind 0: mean((xt-x)^2)   kind 1: mean(|xt-x|).  x, xt fp32, n elements.
 * loss_acc: double[1] and counter: u32[1], zero before first use (the kernel leaves them zero again);
 * loss: fp32[1] = sum/n written by the last block to finish. */
int vp_recon_loss_fwd(const float* x, const float* xt, int64_t n, int kind, double* loss_acc,
                      unsigned int* counter, float* loss, void* stream);
/* dxt = gscale[0] * d loss / d xt   (gscale: device fp32 scalar, the upstream gradient) */
int vp_recon_loss_bwd(const float* x, const float* xt, int64_t n, int kind, const float* gscale,
                      float* dxt, void* stream);
/* 0.5*BCEWithLogits(mean) + dice(sigmoid(logits)) fused: per-sample sums then the scalar.
 * acc: double [rows][4] (zeroed by the call; kept for the backward); counter u32[1] zero before first use;
 * loss fp32[1].  rows = batch, per = elements per sample. */
int vp_bce_dice_fwd(const float* logits, const float* target, int64_t rows, int64_t per, float bce_weight,
                    double* acc, unsigned int* counter, float* loss, void* stream);
int vp_bce_dice_bwd(const float* logits, const float* target, int64_t rows, int64_t per, float bce_weight,
                    const double* acc, const float* gscale, float* dlogits, void* stream);

/* acc[0] += scale * sum(v[0..n))  (single block, deterministic): folds sum_b kl_b (train.py:63) into the loss */
int vp_sum_into(const float* v, int64_t n, float scale, float* acc, void* stream);
/* out[i] = scale * g[0], i in [0,n): the gradient of that sum */
int vp_fill_from(const float* g, float scale, float* out, int64_t n, void* stream);

/* ---- layout conversion at the graph edges ------------------------------------------------------------ */
int vp_nchw_to_nhwc(const float* x_nchw, void* y_nhwc, int dtype, int n, int c, int h, int w, void* stream);
int vp_nhwc_to_nchw(const void* x_nhwc, float* y_nchw, int dtype, int n, int c, int h, int w, void* stream);
/* elementwise cast between fp32 and `dtype` ([n] elements) */
int vp_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);
/* sum += a (fp32, n elements): gradient accumulation helper */
int vp_axpy(float alpha, const float* a, float* sum, int64_t n, void* stream);

/* ---- contractions on the module's own weight (no packed panels) -----------------------------------------------------
 * w_cl: bf16 copy of the layer's weight in channels-last element order, i.e. exactly the bytes of the nn.Parameter when it is
 * kept in torch.channels_last memory format: nn.Conv2d [co][kh][kw][ci], nn.ConvTranspose2d [ci][kh][kw][co], nn.Linear
 * [out][in].  The TMA reads it in place, as a K-major or an MN-major tcgen05 operand depending on the direction.
 * dw_cl: fp32 gradient in the same element order.  prezeroed != 0: the caller guarantees that dw_cl already holds zeros
 * (e.g. cleared by vp_rmsprop_step_shadow), so the call skips its own clearing pass.
 * bf16 activations, tcgen05 engine only:
 * VP_EUNSUPPORTED when the shape is not eligible (reduction channels not a multiple of 64, ...). */
int vp_conv_fwd_cl(const VpConvGeom* g, const void* x, const void* w_cl, const float* bias, void* y, int out_dtype,
                   int act, float slope, void* stream);
int vp_conv_dgrad_cl(const VpConvGeom* g, const void* dy, const void* w_cl, void* dx, int out_dtype, void* stream);
/* Forward of a layer that is followed by BatchNorm (EncoderBlock / DecoderBlock, models/networks.py:14-16,38-40): y in bf16
 * without bias / activation, plus the batch statistics of y taken in the GEMM epilogue from the values being stored --
 * stat_parts[*nparts][2][co] fp32 per-CTA partial column sums and sums of squares (stat_capacity = parts the buffer holds;
 * 2 * #SMs always suffices).  co a multiple of 64, <= 512.  Replaces the separate vp_norm_stats pass over y. */
int vp_conv_fwd_cl_stats(const VpConvGeom* g, const void* x, const void* w_cl, void* y, float* stat_parts,
                         int stat_capacity, int* nparts, void* stream);
/* vp_norm_finalize from such partial sums (added in double, fixed order) */
int vp_norm_finalize_parts(const float* parts, int nparts, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, float momentum, float eps,
                           float* mean, float* invstd, float* scale, float* shift, int64_t rows, int c,
                           int64_t* num_batches_tracked, void* stream);
int vp_conv_wgrad_cl(const VpConvGeom* g, const void* x, const void* dy, float* dw_cl, int prezeroed, void* stream);
/* vp_conv_dgrad_cl of the layer that FOLLOWS a conv -> BatchNorm -> ReLU block (models/networks.py:27-30,42-46), fused with the
 * first pass of that block's BatchNorm backward: dx (= dL/da of the block, bf16) is stored as usual, and the epilogue -- which
 * TMA-loads the matching tile of the block's pre-norm output y_prev [n,hi,wi,ci] -- accumulates d = dx * (y*scale + shift > 0):
 * parts[*nparts][2][ci] fp32 = per-CTA (sum d, sum d*(y - mean)).  vp_norm_bwd_finish_parts turns them into the `sums` that
 * vp_norm_bwd_apply consumes; the separate vp_norm_bwd_reduce pass over (y, da) disappears.  VP_EUNSUPPORTED (nothing launched)
 * when no kernel with this epilogue serves the shape. */
int vp_conv_dgrad_cl_bnred(const VpConvGeom* g, const void* dy, const void* w_cl, void* dx, const void* y_prev,
                           const float* scale, const float* shift, const float* mean, float* parts, int capacity,
                           int* nparts, void* stream);
int vp_norm_bwd_finish_parts(const float* parts, int nparts, const float* invstd, double* sums, int c, void* stream);
/* the same fusion for the decoder's output layer (models/networks.py:101: 64 -> 1 channels): its thin data-gradient kernel
 * produces dL/da of the last DecoderBlock -- the largest activation of the step -- and reduces it against that block's y_prev */
/* dx [n][hi][wi][1] = data gradient of a stride-1 Conv2d with ONE INPUT channel (the discriminator's first layer,
 * models/networks.py:159, whose input x_tilde carries a gradient): the thin-output forward kernel on dy with flipped taps.
 * co % 64 == 0 (pad 32 -> 64 first), k <= 5.  dx in out_dtype. */
int vp_thin_conv_dgrad_in1(const VpConvGeom* g, const void* dy, const float* w, void* dx, int out_dtype, void* stream);
int vp_thin_conv_dgrad_bnred(const VpConvGeom* g, const void* dy, const float* w, void* dx, const void* y_prev,
                             const float* scale, const float* shift, const float* mean, float* parts, int capacity,
                             int* nparts, void* stream);
/* dst[b][c][r] = src[b][r][c] (same dtype): channels-last 8x8 map <-> the NCHW-flatten order of the fc layers */
int vp_transpose_bt(const void* src, void* dst, int dtype, int batch, int rows, int cols, void* stream);

/* ---- thin layers on tcgen05, straight from the fp32 master weight ------------------------------------------------
 * Convolutions with a single channel on one side: the first EncoderBlock conv (models/networks.py:14 with
 * channel_in = 1), the decoder's output conv (models/networks.py:101) and their gradients.  x / dy / y are bf16
 * channels-last (y / dx in `out_dtype`); `w` is the nn.Conv2d weight itself, fp32 [co][ci][kh][kw], and `dw` its
 * gradient in the same layout (zeroed by the call): no packed panels.  VP_EUNSUPPORTED when the shape is not thin.
 *   fwd:   ci == 1 (stride <= 3, kernel <= 8x8; co a multiple of 32, <= 128)   or   co*kh*kw <= 32 at stride 1 (kernel
 *          <= 5x5, ci a multiple of 64, <= 256)
 *   dgrad: co == 1 at stride 1 (ci a multiple of 32, <= 128)
 *   wgrad: ci == 1 (co a multiple of 64)   or   co == 1 at stride 1 (ci a multiple of 64) */
int vp_thin_conv_fwd(const VpConvGeom* g, const void* x, const float* w, const float* bias, void* y, int out_dtype,
                     int act, float slope, void* stream);
/* single-channel-input forward + BatchNorm partial sums of y (as vp_conv_fwd_cl_stats) */
int vp_thin_conv_fwd_stats(const VpConvGeom* g, const void* x, const float* w, void* y, float* stat_parts,
                           int stat_capacity, int* nparts, void* stream);
int vp_thin_conv_dgrad(const VpConvGeom* g, const void* dy, const float* w, void* dx, int out_dtype, void* stream);
int vp_thin_conv_wgrad(const VpConvGeom* g, const void* x, const void* dy, float* dw, int prezeroed, void* stream);

/* ---- optimiser: torch.optim.RMSprop (train.py:136-140; alpha .99, eps 1e-8, no momentum, not centered) as ONE
 * multi-tensor kernel over fp32 masters:  sq = alpha*sq + (1-alpha)*g*g;  p -= lr * g / (sqrt(sq) + eps).
 * params/grads/sq: host arrays of `count` device pointers, numel: host array of element counts. */
int vp_rmsprop_step(void* const* params, const void* const* grads, void* const* sq, const int64_t* numel, int count,
                    float lr, float alpha, float eps, float weight_decay, void* stream);
/* same, and shadows[i] (nullable per entry) receives the bf16 copy of the updated parameter in the same element order:
 * the operand the in-place contractions (vp_conv_*_cl) read next step, so no cast pass is needed.  zero_grads != 0: each
 * gradient is cleared after it has been consumed, so that the next step's weight-gradient kernels can accumulate into
 * it without a memset (vp_conv_wgrad_cl / vp_thin_conv_wgrad with prezeroed = 1). */
int vp_rmsprop_step_shadow(void* const* params, void* const* grads, void* const* sq, void* const* shadows,
                           const int64_t* numel, int count, float lr, float alpha, float eps, float weight_decay,
                           int zero_grads, void* stream);

/* ---- channel padding: layers whose channel counts are not multiples of 64 on the tcgen05 kernels ---------------------------
 * (blocks.Conv2d with 3 / 4 / 6 / 16 / 32 ... channels; the VAE-GAN discriminator's 1 -> 32 -> 64 layers)
 * The activation is copied into a zero-padded [rows, cp] tensor, the weight into a zero-padded bf16 channels-last panel
 * [d0p][taps][d1p] (d0 / d1 = the weight's first two axes, so the panel is what vp_conv_*_cl expect for the padded layer), the
 * vp_conv_*_cl kernels run on the padded geometry, and the fp32 weight gradient is gathered back into the parameter's layout.
 * Zero channels contribute exact zeros: results equal the unpadded contraction. */
int vp_pad_channels(const void* src, int c, void* dst, int cp, int64_t rows, int dtype, void* stream);
int vp_pad_weight_cl(const float* w, void* dst_bf16, int d0, int d1, int taps, int64_t s0, int64_t s1, int64_t st, int d0p,
                     int d1p, void* stream);
int vp_unpad_wgrad_cl(const float* dwp, float* dw, int d0, int d1, int taps, int64_t s0, int64_t s1, int64_t st, int d1p,
                      void* stream);

/* Data-parallel exchange in bf16: dst_bf16[i] = src[i] (round to nearest even) and, with zero_src != 0, src[i] = 0 in the same
 * pass (the fp32 gradient bucket is cleared for the next step's accumulating weight-gradient kernels).  n % 4 == 0, 16-byte
 * aligned.  After the all-reduce, vp_rmsprop_step_wire reads the bf16 buffers directly. */
int vp_pack_grads_bf16(float* src, void* dst_bf16, int64_t n, int zero_src, void* stream);
int vp_rmsprop_step_wire(void* const* params, void* const* grads, void* const* sq, void* const* shadows,
                         const void* const* wire_grads, const int64_t* numel, int count, float lr, float alpha, float eps,
                         float weight_decay, int zero_grads, void* stream);

/* ---- the remaining terms of VaeGan.loss / the train.py step (models/networks.py:264-281, train.py:62-67), fp32 ------------ */
/* out[r] = 0.5 * sum_j (a[r,j] - b[r,j])^2 (feature MSE between discriminator layers, :273); bwd: da = g[r] (a - b), db = -da */
int vp_feature_mse_fwd(const float* a, const float* b, float* out, int64_t rows, int64_t cols, void* stream);
int vp_feature_mse_bwd(const float* a, const float* b, const float* g, float* da, float* db, int64_t rows, int64_t cols,
                       void* stream);
/* out = -log(sign * p + offset): (1, 1e-3) for -log(D(x) + 1e-3), (-1, 1 + 1e-3) for -log(1 - D(.) + 1e-3) (:276-278) */
int vp_neglog_fwd(const float* p, float* out, int64_t n, float sign, float offset, void* stream);
int vp_neglog_bwd(const float* p, const float* g, float* dp, int64_t n, float sign, float offset, void* stream);
/* out[b] = -log(s[b, label[b]]): the pick of F.cross_entropy after its softmax (train_Style_GAN.py:219,226: applied to the
 * discriminator's softmax PROBABILITIES, i.e. a second softmax -- vp_softmax_fwd -- comes first); label int64 [rows] */
int vp_nll_pick_fwd(const float* s, const int64_t* label, float* out, int64_t rows, int cols, void* stream);
int vp_nll_pick_bwd(const float* s, const int64_t* label, const float* g, float* ds, int64_t rows, int cols, void* stream);
/* out[0] = scale * sum smooth_l1(a - b) with beta = 1 (F.smooth_l1_loss(reduction="sum") / B, :279); g: device scalar or NULL */
int vp_smooth_l1_sum_fwd(const float* a, const float* b, float* out, int64_t n, float scale, void* stream);
int vp_smooth_l1_sum_bwd(const float* a, const float* b, const float* g, float* da, float* db, int64_t n, float scale,
                         void* stream);
/* kl[r] = -0.5 * sum_j(-exp(lv) - mu^2 + lv + 1) (:270); backward = vp_reparam_kl_bwd with dz = NULL */
int vp_kl_fwd(const float* mu, const float* logvar, int64_t ld, float* kl, int64_t rows, int zdim, void* stream);

/* torch.optim.Adam (no amsgrad; train_BE.py:131, train_Style_GAN.py) as one multi-tensor kernel over fp32 masters:
 *   g <- g + wd p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).
 * t = `step` (1-based) or *step_dev when non-NULL (a device counter the caller advances with vp_philox_advance: CUDA-graph
 * replay).  shadows / zero_grads as vp_rmsprop_step_shadow. */
int vp_adam_step(void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq, void* const* shadows,
                 const int64_t* numel, int count, float lr, float beta1, float beta2, float eps, float weight_decay,
                 int64_t step, const uint64_t* step_dev, int zero_grads, void* stream);

/* ---- operators of models/blocks.py / models/network_Style_GAN.py around the contractions (channels-last, `dtype`) ----------
 * All one-pass, HBM-bound kernels; forward / backward pairs as the autograd wrappers of vae_play_b200/functional_blocks.py
 * call them.  rows = n*h*w pixels. */
/* dst[r, dst_off : dst_off+nc] (+)= src[r, src_off : src_off+nc]: torch.cat on the channel axis (StyleUp, network_Style_GAN.py:62;
 * Generator.encode :140; Discriminator :222) is two calls, its backward (a slice) one call each; accumulate != 0 adds. */
int vp_copy_channels(const void* src, int src_c, int src_off, void* dst, int dst_c, int dst_off, int nc, int64_t rows,
                     int dtype, int accumulate, void* stream);
/* AddCoords (blocks.py:97-112): out[..., :c] = x, out[..., c] = column index, out[..., c+1] = row index
 * (normalize != 0: (i / extent - 0.5) / 0.5).  out has c + 2 channels.  Backward = vp_copy_channels of the first c. */
int vp_add_coords(const void* x, void* out, int dtype, int64_t n, int h, int w, int c, int normalize, void* stream);
/* F.interpolate(scale_factor=2, mode='bilinear') (align_corners=False; blocks.py:145): y [n,2h,2w,c]; bwd is the exact adjoint */
int vp_upsample2x_fwd(const void* x, void* y, int dtype, int64_t n, int h, int w, int c, void* stream);
int vp_upsample2x_bwd(const void* dy, void* dx, int dtype, int64_t n, int h, int w, int c, void* stream);
/* nn.AdaptiveAvgPool2d((oh, ow)) (SCSEBlock blocks.py:56; networks_BE_GAN.py:99; networks_BP.py:53): y [n,oh,ow,c] */
int vp_avgpool_fwd(const void* x, void* y, int dtype, int64_t n, int h, int w, int c, int oh, int ow, void* stream);
int vp_avgpool_bwd(const void* dy, void* dx, int dtype, int64_t n, int h, int w, int c, int oh, int ow, void* stream);
/* SCSE gate (blocks.py:64-65): y = x * cse[n, c] + x * sse[n, pixel].  bwd: dx = dy * (cse + sse), dsse[pixel] = sum_c dy*x,
 * dcse_f32[n, c] = sum_pixels dy*x (fp32, zeroed by the call). */
int vp_scse_fwd(const void* x, const void* cse, const void* sse, void* y, int dtype, int64_t n, int64_t hw, int c, void* stream);
int vp_scse_bwd(const void* x, const void* cse, const void* sse, const void* dy, void* dx, float* dcse_f32, void* dsse,
                int dtype, int64_t n, int64_t hw, int c, void* stream);
/* myConv2d blend (network_Style_GAN.py:78-79): y = a1 * (1 - label[n]) + a2 * label[n]; label fp32 [n]; per = elements per sample */
int vp_blend_fwd(const void* a1, const void* a2, const float* label, void* y, int dtype, int64_t n, int64_t per, void* stream);
int vp_blend_bwd(const void* dy, const float* label, void* d1, void* d2, int dtype, int64_t n, int64_t per, void* stream);
/* softmax over the last axis of [rows, cols] (blocks.py:73,87; network_Style_GAN.py:228); bwd: dx = y * (dy - sum(dy*y)) */
int vp_softmax_fwd(const void* x, void* y, int dtype, int64_t rows, int cols, void* stream);
int vp_softmax_bwd(const void* y, const void* dy, void* dx, int dtype, int64_t rows, int cols, void* stream);
/* torch.bmm of the attention block (blocks.py:86,90): C[b] (+)= op(A[b]) . op(B[b]), C [batch, m, n] dense; element (i, k)
 * of op(A) at a[b*sa_b + i*sa_i + k*sa_k], element (k, j) of op(B) at b[b*sb_b + k*sb_k + j*sb_j]; fp32 accumulation. */
int vp_bmm(const void* a, const void* b, void* c, int dtype, int batch, int m, int n, int k, int64_t sa_b, int64_t sa_i,
           int64_t sa_k, int64_t sb_b, int64_t sb_k, int64_t sb_j, int accumulate, void* stream);
/* y = gamma[0] * a + x (attention residual, blocks.py:93; gamma a device fp32 scalar) */
int vp_scale_add(const float* gamma, const void* a, const void* x, void* y, int dtype, int64_t n, void* stream);
/* acc[0] = sum(a * b) in double (zeroed by the call): the gradient of gamma */
int vp_dot(const void* a, const void* b, double* acc, int dtype, int64_t n, void* stream);
/* compute_dice_loss on probabilities (tools/ops.py:12-19): 1 - mean_b (2 sum(p t) + smooth) / (sum p + sum t + smooth).
 * acc: double [rows][3] (zeroed by the call, kept for the backward); counter u32[1] zero before first use; loss fp32[1]. */
int vp_dice_fwd(const float* p, const float* t, int64_t rows, int64_t per, float smooth, double* acc, unsigned int* counter,
                float* loss, void* stream);
int vp_dice_bwd(const float* t, int64_t rows, int64_t per, float smooth, const double* acc, const float* gscale, float* dp,
                void* stream);
/* |depthwise 3x3 edge filter| of edge_loss (tools/ops.py:187-211): e = |conv(x, [[-1,-1,-1],[-1,8,-1],[-1,-1,-1]] / 8)|, zero
 * padding, single-channel fp32 maps [n, h, w]; sign (nullable) receives the sign of the response for the backward. */
int vp_edge_fwd(const float* x, float* e, float* sign_or_null, int64_t n, int h, int w, void* stream);
int vp_edge_bwd(const float* de, const float* sign, float* dx, int64_t n, int h, int w, void* stream);

/* number of kernels this library has launched in this process (the bench's gpu_launches claim) */
uint64_t vp_launch_count(void);
/* number of bf16 contractions that ran on the CUDA-core engine instead of tcgen05 (0 on the bf16 hot path: every layer shape
 * of the reference reaches a tensor-core kernel directly, through the thin kernels or through channel padding) */
uint64_t vp_simt_bf16_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VAEPLAY_B200_H */
