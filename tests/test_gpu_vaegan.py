"""GPU parity of config 4 (BASELINE.json): the full ``VaeGan.forward`` 6-tuple, ``VaeGan.loss`` and the five accumulating
``backward(retain_graph=True)`` calls of the reference step (train.py:43-73) against tests/golden/vaegan64_b4.npz, which
oracle/gen_golden_vaegan.py wrote from the UNMODIFIED reference in float64 (plus the reference's own fp32-vs-fp64 deviation
per tensor, used to calibrate the fp32 bound).  Run on the B200 box with ``pytest -m gpu``.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.gen_golden_vaegan import LAMBDA_MSE, digest, synth_vaegan_inputs, synth_vaegan_params
from tests.util import load, rel, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vp():
    import vae_play_b200
    return vae_play_b200


def npy(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


def build(vp, prec, persistent=False):
    import vae_play_b200.functional as VF
    from vae_play_b200.models.networks import VaeGan
    vp.set_precision(prec)
    vp.set_engine("auto")
    VF.set_grad_sinks({})
    net = VaeGan(64, 128)
    P = synth_vaegan_params(0)
    missing = net.load_state_dict({k: torch.from_numpy(v) for k, v in P.items()}, strict=False)
    assert not missing.unexpected_keys
    net = net.cuda().train()
    if persistent:
        VF.persistent_grads(list(net.parameters()))
    return net


def run_step(net, five=True, batch=4):
    """The reference step, line for line (train.py:43-73), on the mirror."""
    from vae_play_b200.models.networks import VaeGan
    x_np, eps_np, zp_np, t_np = synth_vaegan_inputs(0, batch)
    x, eps, z_p, targets = (torch.from_numpy(a).cuda() for a in (x_np, eps_np, zp_np, t_np))
    b = len(x)
    x_tilde, disc_class, disc_layer, mus, log_variances, params = net(x, eps=eps, z_p=z_p)
    dl_o, dl_p, dl_s = disc_layer[:b], disc_layer[b:-b], disc_layer[-b:]
    dc_o, dc_p, dc_s = disc_class[:b], disc_class[b:-b], disc_class[-b:]
    nle, kl, mse, bce_o, bce_p, bce_s, l1 = VaeGan.loss(x, x_tilde, dl_o, dl_p, dl_s, dc_o, dc_p, dc_s, mus, log_variances, targets, params)
    loss_recon = F.mse_loss(x, x_tilde)
    loss_encoder = torch.sum(kl) + torch.sum(mse)
    loss_discriminator = torch.sum(bce_o) + torch.sum(bce_p) + torch.sum(bce_s)
    loss_decoder = torch.sum(LAMBDA_MSE * mse) - (1.0 - LAMBDA_MSE) * loss_discriminator
    loss_aux = l1
    out = {"x_tilde": x_tilde, "disc_class": disc_class, "disc_layer": disc_layer, "mus": mus, "log_variances": log_variances,
           "params": params, "kl": kl, "mse": mse, "bce_dis_original": bce_o, "bce_dis_predicted": bce_p, "bce_dis_sampled": bce_s,
           "l1_enc_param": l1, "nle": nle,
           "losses": torch.stack([loss_recon, loss_encoder, loss_decoder, loss_discriminator, loss_aux])}
    single = {}
    if five:
        dparams = [(k, p) for k, p in net.named_parameters() if k.startswith("discriminator.")]
        for name, loss in (("only_loss_discriminator", loss_discriminator), ("only_sum_mse", torch.sum(mse))):
            gs = torch.autograd.grad(loss, [p for _, p in dparams], retain_graph=True, allow_unused=True)
            for (k, _), g in zip(dparams, gs):
                if g is not None:
                    single[f"{name}/{k}"] = npy(g)
        net.zero_grad()
        loss_recon.backward(retain_graph=True)
        loss_encoder.backward(retain_graph=True)
        loss_decoder.backward(retain_graph=True)
        loss_discriminator.backward(retain_graph=True)
        loss_aux.backward()
    return out, single


def fixture(batch=4):
    g = load(f"vaegan64_b{batch}.npz")
    dev = dict(zip([str(k) for k in g["ref_fp32_dev_keys"]], [float(v) for v in g["ref_fp32_dev_vals"]]))
    return g, dev


def grad_dev(got, g, key):
    """(max-norm deviation, relative deviation of the l2 norm) of one gradient from the fixture (full tensor or digest)."""
    want = g["grad/" + key]
    if bool(g["gradfull/" + key][0]):
        return rel(got, want), rel_l2(got, want)
    d = digest(got)          # digest: (sum, l2, max, 256 strided samples) -- the samples are compared against the tensor's max
    return float(np.abs(d[3:] - want[3:]).max() / want[2]), abs(d[1] - want[1]) / want[1]


# The one tensor (128 elements) whose fp32 deviation exceeds max(1e-5, 3 x the reference's own fp32-vs-fp64 deviation) at batch 4:
# measured 2.03e-5 against a reference deviation of 4.05e-6 (a 4-term cancelling column sum); at batch 16 it is 7.6e-6 (ratio 0.76).
FP32_EXCEPTIONS = {(4, "encoder.l_mu.bias"): 3e-5}
# ReLU flips: with ~1e7 ReLU inputs per step some pre-activations lie within fp32 round-off of zero; the implementation under
# test and the reference's own fp32 run need not flip the SAME elements, and one flipped element changes every gradient
# UPSTREAM of that ReLU by O(1 / elements per channel) -- measured 8e-5 .. 2e-3 on the discriminator's conv.2 / conv.3 at batch
# 16 (while everything downstream of the flipped layer, fc.*, stays at 1e-6).  tests/test_gpu_parity.py pins the pattern through
# the NumPy oracle for the VAE path; there is no oracle of the discriminator graph, so a bounded number of tensors may exceed
# the per-tensor bound by a flip-sized amount instead.
FLIP_BOUND, FLIP_MAX_TENSORS = 5e-3, 8


@pytest.mark.parametrize("batch", [4, 16])
def test_vaegan_step_golden_fp32(vp, batch):
    """fp32 check mode: every forward output, every accumulated parameter gradient (five backward calls) and every single-loss
    discriminator gradient to max(1e-5, 3 x the reference's own fp32 deviation) PER TENSOR (measured: <= 0.37 of that bound
    for 61 of 62 parameters at batch 4, all 62 at batch 16; see FP32_EXCEPTIONS for the one exception)."""
    g, dev = fixture(batch)
    try:
        net = build(vp, "fp32")
        out, single = run_step(net, batch=batch)
        bad = []
        for k in ("x_tilde", "disc_class", "mus", "log_variances", "params", "kl", "mse", "losses", "bce_dis_original", "bce_dis_predicted",
                  "bce_dis_sampled", "l1_enc_param"):
            t = max(1e-5, 3 * dev.get(k, 0.0))
            r = rel(npy(out[k]).reshape(g[k].shape), g[k])
            if r >= t:
                bad.append((k, r, t))
        assert tuple(out["disc_layer"].shape) == tuple(g["disc_layer_shape"])
        d = digest(npy(out["disc_layer"]))
        assert np.abs(d[3:] - g["disc_layer_digest"][3:]).max() / g["disc_layer_digest"][2] < 1e-5
        assert np.abs(digest(npy(out["nle"]))[3:] - g["nle_digest"][3:]).max() / g["nle_digest"][2] < 1e-5
        grads = {k: npy(p.grad) for k, p in net.named_parameters()}
        assert all(p.grad is not None for p in net.parameters())
        grads.update(single)
        assert len(single) >= 20
        for k, got in grads.items():
            t = FP32_EXCEPTIONS.get((batch, k), max(1e-5, 3 * dev[k]))
            r, _ = grad_dev(got, g, k)
            if r >= t:
                bad.append((k, r, t))
        flips = [b_ for b_ in bad if b_[1] < FLIP_BOUND and ("conv" in b_[0] or "fc.0" in b_[0] or "fc.1" in b_[0])]
        hard = [b_ for b_ in bad if b_ not in flips]
        assert not hard and len(flips) <= FLIP_MAX_TENSORS, "\n".join(f"{k}: rel {r:.3e} >= {t:.2e}" for k, r, t in bad)
    finally:
        vp.set_precision("bf16")


U_RATIO = 2.0 ** (24 - 9)      # unit roundoff of bf16 storage / fp32


@pytest.mark.parametrize("batch", [4, 16])
def test_vaegan_step_golden_bf16(vp, batch):
    """bf16 tensor-core mode.  Forward outputs within the end-to-end bf16 bound of the VAE step tests.  Gradients: the
    reference's own fp32-vs-fp64 deviation of each tensor measures how much that tensor amplifies rounding noise (the
    feature-MSE term of loss_encoder is ~1e4 and flows back through three train-mode BatchNorms per network; at batch 4 / 16
    single ReLU flips move whole channels).  Scaled by the ratio of the unit roundoffs it predicts the bf16 noise floor:
        tol = max(floor, 2^15 * ref_fp32_dev),  floor = 3e-2 without ReLUs on the path (param_encoder), 2.5e-1 with
    Tensors with tol < 1 are held to it in max-norm (all of param_encoder, the discriminator's feature path, biases ...); for
    the ill-conditioned rest (tol >= 1: the reference's fp32 run itself is off by > 3e-5) only the gradient's l2 norm is
    compared, loosely -- a structurally wrong backward (missing term, wrong sign, aliasing) still fails it.  The
    discriminator's ACCUMULATED gradients cancel to 1e-6 of their terms (loss_decoder carries -(1 - 1e-6) loss_discriminator,
    train.py:66; the reference's fp32 run is off by up to 49 %) and are checked through the single-loss gradients instead."""
    g, dev = fixture(batch)
    net = build(vp, "bf16")
    out, single = run_step(net, batch=batch)
    for k, t in (("x_tilde", 2.5e-2), ("mus", 3e-2), ("log_variances", 3e-2), ("params", 2.5e-2), ("disc_class", 2.5e-2), ("kl", 2.5e-2),
                 ("mse", 1e-2), ("losses", 1e-2)):
        r = rel(npy(out[k]).reshape(g[k].shape), g[k])
        assert r < t, f"{k}: rel {r:.3e}"
    grads = {k: npy(p.grad) for k, p in net.named_parameters() if not k.startswith("discriminator.")}
    grads.update(single)
    bad, tight = [], 0
    for k, got in grads.items():
        assert np.isfinite(got).all(), k
        # linear-only path (param_encoder: eight Linears, no ReLU): bf16 rounding only.  Everything else sits upstream of
        # ReLUs whose pattern differs in ~0.4 % of the elements at bf16 noise: measured 4e-2 .. 1.6e-1 on BatchNorm affine
        # gradients (sums over few elements per channel at these batch sizes)
        floor = 3e-2 if k.startswith("param_encoder.") else 2.5e-1
        tol = max(floor, U_RATIO * dev[k])
        r, rn = grad_dev(got, g, k)
        if tol < 1.0:
            tight += 1
            if r >= tol:
                bad.append((k, "max-norm", r, tol))
        elif rn >= 0.6:
            bad.append((k, "l2 norm", rn, 0.6))
    assert not bad, "\n".join(f"{k}: {w} {r:.3e} >= {t:.2e}" for k, w, r, t in bad)
    assert tight >= 20, tight


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_weight_used_twice_with_grad_slots(vp, prec):
    """One weight used twice in the same backward (the decoder runs on z and on z_p, the discriminator in REC and GAN mode):
    with persistent gradient slots installed the SECOND use must not overwrite the slot the first use was handed
    (the slot was given out twice and autograd summed two aliases of the same memory).  Gradients with slots ==
    gradients without, for the whole five-backward step."""
    import vae_play_b200.functional as VF
    try:
        ref = build(vp, prec, persistent=False)
        run_step(ref)
        want = {k: npy(p.grad) for k, p in ref.named_parameters()}
        net = build(vp, prec, persistent=True)
        run_step(net)
        tol = 1e-4 if prec == "fp32" else 5e-2         # same kernels, same operands; fp32 atomics order / bf16 wgrad split order
        for k, p in net.named_parameters():
            if prec == "bf16" and not k.startswith("param_encoder."):
                # bf16: two runs of the SAME flow already differ by O(10 %) on the ill-conditioned tensors (fp32 atomics order ->
                # BatchNorm statistics -> ReLU flips); the linear-only param_encoder path is the exact check, fp32 covers the rest
                assert torch.isfinite(p.grad).all(), k
                continue
            if k.startswith("discriminator."):
                assert rel_l2(npy(p.grad), want[k]) < 0.5, k
                continue
            r = rel_l2(npy(p.grad), want[k])
            assert r < tol, f"{k}: rel-L2 {r:.3e} with gradient slots vs without"
        # and a minimal case: y = layer(layer(x)) with one shared weight
        vp.set_precision(prec)
        lay = VF.TapLayer("linear", 64, 64)
        w = torch.nn.Parameter(torch.randn(64, 64, device="cuda") * 0.1)
        x = torch.randn(32, 1, 1, 64, device="cuda").to(VF.act_dtype())
        res = []
        for slots in (False, True):
            VF.set_grad_sinks({})
            w.grad = None
            if slots:
                VF.persistent_grads([w])
            h, _ = VF.fused_layer(x, w, None, None, None, lay, VF.NormCfg(None), "tanh", 0.0, True, None)
            y, _ = VF.fused_layer(h, w, None, None, None, lay, VF.NormCfg(None), "none", 0.0, True, None)
            y.float().sum().backward()
            res.append(npy(w.grad))
        assert rel_l2(res[1], res[0]) < (1e-5 if prec == "fp32" else 2e-2)
        xs = x.float().double().reshape(32, 64)
        wd = w.detach().double()
        if prec == "fp32":
            hd = torch.tanh(xs @ wd.T)
            gy = torch.ones(32, 64, device="cuda", dtype=torch.float64)
            dh = (gy @ wd) * (1 - hd * hd)
            want_w = gy.T @ hd + dh.T @ xs
            assert rel(res[1], npy(want_w)) < 1e-5
    finally:
        VF.set_grad_sinks({})
        vp.set_precision("bf16")


def test_eval_mode_forward(vp):
    """eval(): BatchNorm uses the running statistics (reference networks.py:248-258: returns (x_tilde, params), or decoded
    samples when x is None)."""
    import vae_play_b200.functional as VF
    try:
        net = build(vp, "fp32")
        run_step(net, five=False)               # one training forward moves the running statistics off their init values
        net.eval()
        x_np, eps_np, _, _ = synth_vaegan_inputs(0)
        x, eps = torch.from_numpy(x_np).cuda(), torch.from_numpy(eps_np).cuda()
        with torch.no_grad():
            x_tilde, params = net(x, eps=eps)
            samples = net(None, gen_size=3)
        assert tuple(x_tilde.shape) == (4, 1, 64, 64) and tuple(params.shape) == (4, 3) and tuple(samples.shape) == (3, 1, 64, 64)
        # float64 torch evaluation of the same eval-mode graph from the module's own parameters / running statistics
        sd = {k: v.detach().double() for k, v in net.state_dict().items()}
        h = x.double()
        for i in range(3):
            h = F.conv2d(h, sd[f"encoder.conv.{i}.conv.weight"], None, stride=2, padding=2)
            h = F.batch_norm(h, sd[f"encoder.conv.{i}.bn.running_mean"], sd[f"encoder.conv.{i}.bn.running_var"], sd[f"encoder.conv.{i}.bn.weight"],
                             sd[f"encoder.conv.{i}.bn.bias"], False, 0.0, 1e-5).relu()
        h = h.reshape(len(h), -1) @ sd["encoder.fc.0.weight"].T
        h = F.batch_norm(h, sd["encoder.fc.1.running_mean"], sd["encoder.fc.1.running_var"], sd["encoder.fc.1.weight"], sd["encoder.fc.1.bias"],
                         False, 0.0, 1e-5).relu()
        mu = h @ sd["encoder.l_mu.weight"].T + sd["encoder.l_mu.bias"]
        lv = h @ sd["encoder.l_var.weight"].T + sd["encoder.l_var.bias"]
        z = eps.double() * torch.exp(0.5 * lv) + mu
        h = z @ sd["decoder.fc.0.weight"].T
        h = F.batch_norm(h, sd["decoder.fc.1.running_mean"], sd["decoder.fc.1.running_var"], sd["decoder.fc.1.weight"], sd["decoder.fc.1.bias"],
                         False, 0.0, 1e-5).relu().reshape(len(z), -1, 8, 8)
        for i in range(3):
            h = F.conv_transpose2d(h, sd[f"decoder.conv.{i}.conv.weight"], None, stride=2, padding=2, output_padding=1)
            h = F.batch_norm(h, sd[f"decoder.conv.{i}.bn.running_mean"], sd[f"decoder.conv.{i}.bn.running_var"], sd[f"decoder.conv.{i}.bn.weight"],
                             sd[f"decoder.conv.{i}.bn.bias"], False, 0.0, 1e-5).relu()
        want = torch.sigmoid(F.conv2d(h, sd["decoder.conv.3.0.weight"], sd["decoder.conv.3.0.bias"], padding=2))
        assert rel(npy(x_tilde), npy(want)) < 2e-5
    finally:
        vp.set_precision("bf16")
        VF.set_grad_sinks({})


def test_fused_single_backward_equals_five_backward(vp):
    """train_steps.vaegan_step(fused=True): ONE backward of the summed loss gives the gradients the reference accumulates over
    its five backward(retain_graph=True) calls (train.py:68-73).  fp32 check mode, batch 16 fixture: encoder / decoder /
    param_encoder gradients agree with the five-call mode to fp32 round-off; the discriminator's pure-lambda gradients
    (BatchNorm of the last block, fc.*: lambda = 1e-6 times the GAN-loss gradient, which the reference forms as the
    difference of two nearly equal fp32 numbers and gets wrong by 10-50 %) match the FLOAT64 truth to a ReLU flip (2e-3 measured:
    the same conv.3 flip test_vaegan_step_golden_fp32[16] sees) in the fused mode."""
    from vae_play_b200 import train_steps as TS
    g, dev = fixture(16)
    try:
        grads = {}
        for fused in (False, True):
            net = build(vp, "fp32")
            x_np, eps_np, zp_np, t_np = synth_vaegan_inputs(0, 16)
            x, eps, z_p, targets = (torch.from_numpy(a).cuda() for a in (x_np, eps_np, zp_np, t_np))
            net.zero_grad()
            losses, parts = TS.vaegan_losses(net, x, targets, eps=eps, z_p=z_p)
            TS.vaegan_backward(losses, parts, fused=fused)
            grads[fused] = {k: npy(p.grad) for k, p in net.named_parameters()}
            for i, k in enumerate(("loss_recon", "loss_encoder", "loss_decoder", "loss_discriminator", "loss_aux")):
                assert abs(float(losses[k]) - g["losses"][i]) <= 2e-5 * abs(g["losses"][i]) + 1e-7, k
        for k in grads[True]:
            if not k.startswith("discriminator."):
                assert rel_l2(grads[True][k], grads[False][k]) < 1e-4, k
        pure_lambda = [k for k in grads[True] if k.startswith("discriminator.fc.") or k.startswith("discriminator.conv.3.bn.")]
        assert len(pure_lambda) >= 6
        for k in pure_lambda:
            r, _ = grad_dev(grads[True][k], g, k)
            assert r < FLIP_BOUND, f"{k}: fused-mode deviation from the float64 truth {r:.3e} (reference fp32: {dev[k]:.2e})"
    finally:
        vp.set_precision("bf16")


def test_async_wgrad_with_reused_weights(vp):
    """functional.set_async_wgrad on the full VAE-GAN step: the decoder runs on z and z_p and the discriminator in REC and GAN
    mode, so their weights receive two gradients per backward -- the first into the persistent slot on the side stream, the
    later ones into scratch buffers that are added to the slot on the same stream (autograd sees None for them).  ONE forward,
    backward with the side stream off / off / on / on: the same gradients up to the run-to-run noise of the main-stream path."""
    import vae_play_b200.functional as VF
    import vae_play_b200.functional_blocks as VB
    from vae_play_b200 import train_steps as TS
    from vae_play_b200.models.networks import VaeGan
    vp.set_precision("bf16")
    vp.set_engine("auto")
    VF.set_fuse_bn_backward(False)          # bit-comparable backward passes
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    try:
        with torch.cuda.stream(side):
            torch.manual_seed(5)
            net = VaeGan(64, 128).cuda().train()
            params = list(net.parameters())
            flat = VF.persistent_grads(params)
            x, targets = torch.rand(8, 1, 64, 64, device="cuda"), torch.rand(8, 3, device="cuda")
            eps, z_p = torch.randn(8, 128, device="cuda"), torch.randn(8, 128, device="cuda")
            losses, parts = TS.vaegan_losses(net, x, targets, eps=eps, z_p=z_p)
            lam = TS.LAMBDA_MSE
            total = VB.weighted_sums([parts["recon"], parts["l1"], parts["kl"], parts["mse"], parts["bce_o"], parts["bce_p"], parts["bce_s"]],
                                     [1.0, 1.0, 1.0, 1.0 + lam, lam, lam, lam])
            res = []
            for on in (False, False, True, True):
                VF.set_async_wgrad(on)
                for p in params:
                    p.grad = None
                flat.zero_()
                for e in VF._GRAD_SINKS.values():
                    e[3], e[4] = True, False
                total.backward(retain_graph=True)
                VF.join_async()
                torch.cuda.synchronize()
                res.append([p.grad.detach().double().cpu().numpy() if p.grad is not None else None for p in params])
        names = [k for k, _ in net.named_parameters()]
        # split-K data gradients (fp32 red.add in arrival order) followed by bf16 storage make even two identical backward passes
        # differ by a few flipped roundings: the side stream must stay within a small multiple of that run-to-run noise
        for k, a, a2, b, c in zip(names, *res):
            if a is None:
                assert a2 is None and b is None and c is None, k
                continue
            nrm = np.sqrt((a * a).sum()) + 1e-30
            noise = np.sqrt(((a2 - a) ** 2).sum()) / nrm
            for other in (b, c):
                d = np.sqrt(((other - a) ** 2).sum()) / nrm
                assert d < max(1e-5, 4 * noise), f"{k}: async vs sync {d:.3e}, sync vs sync {noise:.3e}"
    finally:
        VF.set_async_wgrad(False)
        VF.set_fuse_bn_backward(True)
        VF.set_grad_sinks({})
