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


def run_step(net, five=True):
    """The reference step, line for line (train.py:43-73), on the mirror."""
    from vae_play_b200.models.networks import VaeGan
    x_np, eps_np, zp_np, t_np = synth_vaegan_inputs(0)
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


def fixture():
    g = load("vaegan64_b4.npz")
    dev = dict(zip([str(k) for k in g["ref_fp32_dev_keys"]], [float(v) for v in g["ref_fp32_dev_vals"]]))
    return g, dev


def compare_grad(got, g, key, tol, what):
    want = g["grad/" + key]
    if bool(g["gradfull/" + key][0]):
        r = rel(got, want)
    else:      # digest: (sum, l2, max, 256 strided samples) -- compare the samples against the tensor's max
        d = digest(got)
        r = float(np.abs(d[3:] - want[3:]).max() / want[2])
        r = max(r, abs(d[1] - want[1]) / want[1])
    assert r < tol, f"{what} {key}: rel {r:.3e} >= {tol:.2e}"
    return r


def test_vaegan_step_golden_fp32(vp):
    """fp32 check mode: every forward output to 1e-5 (or 3x the reference's own fp32 deviation), every accumulated parameter
    gradient and every single-loss discriminator gradient to max(1e-5, 3 x reference fp32 deviation) PER TENSOR."""
    g, dev = fixture()
    try:
        net = build(vp, "fp32")
        out, single = run_step(net)
        for k in ("x_tilde", "disc_class", "mus", "log_variances", "params", "kl", "mse", "losses", "bce_dis_original", "bce_dis_predicted",
                  "bce_dis_sampled", "l1_enc_param"):
            t = max(1e-5, 3 * dev.get(k, 0.0))
            r = rel(npy(out[k]).reshape(g[k].shape), g[k])
            assert r < t, f"{k}: rel {r:.3e} >= {t:.1e}"
        assert tuple(out["disc_layer"].shape) == tuple(g["disc_layer_shape"])
        d = digest(npy(out["disc_layer"]))
        assert np.abs(d[3:] - g["disc_layer_digest"][3:]).max() / g["disc_layer_digest"][2] < 1e-5
        assert np.abs(digest(npy(out["nle"]))[3:] - g["nle_digest"][3:]).max() / g["nle_digest"][2] < 1e-5
        worst = {}
        for k, p in net.named_parameters():
            assert p.grad is not None, k
            worst[k] = compare_grad(npy(p.grad), g, k, max(1e-5, 3 * dev[k]), "accumulated grad") / max(1e-5, 3 * dev[k])
        for k, got in single.items():
            compare_grad(got, g, k, max(1e-5, 3 * dev[k]), "single-loss grad")
        assert len(single) >= 20
    finally:
        vp.set_precision("bf16")


def test_vaegan_step_golden_bf16(vp):
    """bf16 tensor-core mode at batch 4 (BatchNorm1d over 4 / 12 samples): forward outputs within the end-to-end bf16 bound
    of the VAE step tests; well-conditioned gradients (encoder / decoder / param_encoder accumulated over the five backward
    calls, and the discriminator's single-loss gradients -- the REC-feature path included) within a loose bound that only a
    structurally wrong backward would miss.  The discriminator's ACCUMULATED gradients cancel to 1e-6 of their terms in the
    reference itself (fp32 deviates from fp64 by up to 49 %, ref_fp32_dev) and are not compared."""
    g, dev = fixture()
    net = build(vp, "bf16")
    out, single = run_step(net)
    for k, t in (("x_tilde", 3e-2), ("mus", 1.2e-1), ("log_variances", 1.2e-1), ("params", 1.2e-1), ("disc_class", 1.2e-1), ("kl", 1.2e-1)):
        r = rel(npy(out[k]).reshape(g[k].shape), g[k])
        assert r < t, f"{k}: rel {r:.3e}"
    assert abs(float(out["losses"][0]) - g["losses"][0]) / g["losses"][0] < 3e-2
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        if k.startswith("discriminator."):
            continue
        want = g["grad/" + k]
        got = npy(p.grad)
        if bool(g["gradfull/" + k][0]):
            assert rel_l2(got, want) < 0.35, (k, rel_l2(got, want))
        else:
            d = digest(got)
            assert abs(d[1] - want[1]) / want[1] < 0.35, (k, d[1], want[1])
    n = 0
    for k, got in single.items():
        want = g["grad/" + k]
        if bool(g["gradfull/" + k][0]):
            assert rel_l2(got, want) < 0.35, (k, rel_l2(got, want))
        else:
            assert abs(digest(got)[1] - want[1]) / want[1] < 0.35, k
        n += 1
    assert n >= 20


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
            if k.startswith("discriminator.") and prec == "bf16":
                continue                                 # cancel to 1e-6 of their terms (see above): noise in bf16
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
