"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the reference-generated
golden fixtures.  Run on the B200 box with ``pytest -m gpu``.
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import philox
from oracle import vae_numpy as vn
from tests.util import TOL_BF16, TOL_FP32, digest, load, ref_dev, rel, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vp():
    import vae_play_b200
    return vae_play_b200


@pytest.fixture(params=["fp32", "bf16"])
def mode(request, vp):
    vp.set_precision(request.param)
    vp.set_engine("auto")
    yield request.param
    vp.set_precision("bf16")


def tol_for(mode):
    return TOL_FP32 if mode == "fp32" else TOL_BF16


def cu(a, requires_grad=False):
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32))).cuda()
    return t.requires_grad_(requires_grad)


def npy(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


def close(got, want, tol, what=""):
    r = rel(got, want)
    assert r < tol, f"{what}: rel {r:.3e} >= tol {tol:.1e}"


def gclose(mode, got, want, t, what=""):
    """Gradient of a single golden op.  fp32: max-norm bound.  bf16: the tiny fixtures contain ReLU/LeakyReLU
    elements whose pre-activation is within bf16 rounding of zero, so an element-wise bound is meaningless
    there; bound the relative L2 error instead (element-wise gradient parity with the activation pattern
    pinned is asserted by test_vae_step_vs_oracle)."""
    if mode == "fp32":
        close(got, want, t, what)
    else:
        r = rel_l2(got, want)
        assert r < 0.15, f"{what}: rel-L2 {r:.3e}"


# ------------------------------------------------------------------------------------------------
# RNG: bit-exact against torch's own CUDA normal_() (the reference's eps draw, networks.py:230)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 7, 1000, 256 * 128, 303104, 303105, 1184 * 256 * 4 + 3, 3_000_001])
def test_philox_bit_exact_vs_torch(vp, n):
    dev = torch.device("cuda", 0)
    for seed in (0, 1234, 2 ** 40 + 17):
        torch.cuda.manual_seed(seed)
        pre = torch.empty(5, device=dev).normal_()          # move the generator off offset 0
        want = torch.empty(n, device=dev).normal_()
        off_after_want = torch.cuda.default_generators[0].get_offset()
        torch.cuda.manual_seed(seed)
        pre2 = torch.empty(5, device=dev).normal_()
        got = vp.philox_normal((n,), dev)
        off_after_got = torch.cuda.default_generators[0].get_offset()
        assert torch.equal(pre, pre2)
        assert off_after_got == off_after_want
        assert torch.equal(got.view(torch.int32), want.view(torch.int32)), f"n={n} seed={seed}"


def test_philox_matches_oracle_integers(vp):
    # the NumPy oracle reproduces the same Philox counters; floats agree to libm-vs-intrinsic round-off
    dev = torch.device("cuda", 0)
    n, seed = 4096, 99
    torch.cuda.manual_seed(seed)
    got = npy(vp.philox_normal((n,), dev))
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    want = philox.aten_normal(n, seed, 0, num_sms=sms).astype(np.float64)
    assert np.abs(got - want).max() < 2e-5


def test_reparam_uses_generator_stream(vp):
    dev = torch.device("cuda", 0)
    mu = torch.randn(64, 128, device=dev)
    lv = torch.randn(64, 128, device=dev) * 0.1
    torch.cuda.manual_seed(7)
    eps = torch.empty(64, 128, device=dev).normal_()
    want = eps * torch.exp(0.5 * lv) + mu
    torch.cuda.manual_seed(7)
    z, kl = vp.reparam_kl(mu, lv)
    close(npy(z), npy(want), 1e-6, str('npy(want)'))
    want_kl = -0.5 * torch.sum(-lv.exp() - mu ** 2 + lv + 1, 1)
    close(npy(kl), npy(want_kl), 1e-5, str('npy(want_kl)'))


# ------------------------------------------------------------------------------------------------
# per-operator golden fixtures (tests/golden/ops.npz, produced by the reference's own classes)
# ------------------------------------------------------------------------------------------------
def _load_block(m, ops, name, has_bn_affine=True):
    with torch.no_grad():
        m.conv.weight.copy_(cu(ops[f"{name}/w"]))
        m.bn.weight.copy_(cu(ops[f"{name}/g"]))
        m.bn.bias.copy_(cu(ops[f"{name}/b"]))


def test_encoder_block_golden(vp, mode):
    from vae_play_b200.models.networks import EncoderBlock
    ops = load("ops.npz")
    m = EncoderBlock(6, 10).cuda().train()
    _load_block(m, ops, "eb")
    x = cu(ops["eb/x"], True)
    y, ypre = m(x, out=True)
    t = tol_for(mode)
    close(npy(ypre), ops["eb/ypre"], t, str('ops["eb/ypre"]'))
    close(npy(y), ops["eb/y"], t, str('ops["eb/y"]'))
    y.backward(cu(ops["eb/dy"]))
    gclose(mode, npy(x.grad), ops["eb/dx"], 3 * t, str('ops["eb/dx"]'))
    gclose(mode, npy(m.conv.weight.grad), ops["eb/dw"], 3 * t, str('ops["eb/dw"]'))
    gclose(mode, npy(m.bn.weight.grad), ops["eb/dg"], 3 * t, str('ops["eb/dg"]'))
    gclose(mode, npy(m.bn.bias.grad), ops["eb/db"], 3 * t, str('ops["eb/db"]'))
    close(npy(m.bn.running_mean), ops["eb/rm"], t, str('ops["eb/rm"]'))
    close(npy(m.bn.running_var), ops["eb/rv"], t, str('ops["eb/rv"]'))
    assert int(m.bn.num_batches_tracked) == 1


def test_decoder_block_golden(vp, mode):
    from vae_play_b200.models.networks import DecoderBlock
    ops = load("ops.npz")
    m = DecoderBlock(10, 6).cuda().train()
    _load_block(m, ops, "db")
    x = cu(ops["db/x"], True)
    y = m(x)
    t = tol_for(mode)
    close(npy(y), ops["db/y"], t, str('ops["db/y"]'))
    y.backward(cu(ops["db/dy"]))
    gclose(mode, npy(x.grad), ops["db/dx"], 3 * t, str('ops["db/dx"]'))
    gclose(mode, npy(m.conv.weight.grad), ops["db/dw"], 3 * t, str('ops["db/dw"]'))
    gclose(mode, npy(m.bn.weight.grad), ops["db/dg"], 3 * t, str('ops["db/dg"]'))
    gclose(mode, npy(m.bn.bias.grad), ops["db/db"], 3 * t, str('ops["db/db"]'))
    close(npy(m.bn.running_mean), ops["db/rm"], t, str('ops["db/rm"]'))
    close(npy(m.bn.running_var), ops["db/rv"], t, str('ops["db/rv"]'))


@pytest.mark.parametrize("name,ci,co,k,s,bn,act", [
    ("c_k3s1_batch_relu", 5, 7, 3, 1, "batch", "relu"), ("c_k4s2_inst_lrelu", 4, 6, 4, 2, "instance", "lrelu"),
    ("c_k1s1_none_tanh", 6, 3, 1, 1, None, "tanh"), ("c_k5s1_none_none", 3, 2, 5, 1, None, None),
    ("c_k3s2_batch_lrelu", 4, 8, 3, 2, "batch", "lrelu")])
def test_blocks_conv2d_golden(vp, mode, name, ci, co, k, s, bn, act):
    from vae_play_b200.models.blocks import Conv2d
    ops = load("ops.npz")
    g = lambda key: ops[f"{name}/{key}"]
    m = Conv2d(ci, co, k, stride=s, bn=bn, activate=act).cuda().train()
    with torch.no_grad():
        m.conv[0].weight.copy_(cu(g("w")))
        if bn is None:
            m.conv[0].bias.copy_(cu(g("bias")))
        if bn == "batch":
            m.conv[1].weight.copy_(cu(g("g")))
            m.conv[1].bias.copy_(cu(g("b")))
    x = cu(g("x"), True)
    y = m(x)
    t = tol_for(mode)
    close(npy(y), g("y"), t, str('g("y")'))
    y.backward(cu(g("dy")))
    gclose(mode, npy(x.grad), g("dx"), 3 * t, str('g("dx")'))
    gclose(mode, npy(m.conv[0].weight.grad), g("dw"), 3 * t, str('g("dw")'))
    if bn is None:
        gclose(mode, npy(m.conv[0].bias.grad), g("dbias"), 3 * t, str('g("dbias")'))
    if bn == "batch":
        gclose(mode, npy(m.conv[1].weight.grad), g("dg"), 3 * t, str('g("dg")'))
        gclose(mode, npy(m.conv[1].bias.grad), g("db"), 3 * t, str('g("db")'))


def test_conv_transpose_k4_bias_golden(vp, mode):
    import vae_play_b200.functional as VF
    ops = load("ops.npz")
    layer = VF.TapLayer("convT", 5, 4, k=4, stride=2, pad=1, out_pad=0)
    w = cu(ops["ct4/w"], True)
    b = cu(ops["ct4/bias"], True)
    x = cu(ops["ct4/x"], True)
    y, _ = VF.fused_layer(VF.to_channels_last(x), w, b, None, None, layer, VF.NormCfg(None), "none", 0.0, True, None)
    y = VF.from_channels_last(y)
    t = tol_for(mode)
    close(npy(y), ops["ct4/y"], t, str('ops["ct4/y"]'))
    y.backward(cu(ops["ct4/dy"]))
    gclose(mode, npy(x.grad), ops["ct4/dx"], 3 * t, str('ops["ct4/dx"]'))
    gclose(mode, npy(w.grad), ops["ct4/dw"], 3 * t, str('ops["ct4/dw"]'))
    gclose(mode, npy(b.grad), ops["ct4/dbias"], 3 * t, str('ops["ct4/dbias"]'))


def test_blocks_linear_golden(vp, mode):
    from vae_play_b200.models.blocks import Linear
    ops = load("ops.npz")
    m = Linear(9, 7, bias=True, activate="lrelu").cuda()
    with torch.no_grad():
        m.fc[0].weight.copy_(cu(ops["lin/w"]))
        m.fc[0].bias.copy_(cu(ops["lin/bias"]))
    x = cu(ops["lin/x"], True)
    y = m(x)
    t = tol_for(mode)
    close(npy(y), ops["lin/y"], t, str('ops["lin/y"]'))
    y.backward(cu(ops["lin/dy"]))
    gclose(mode, npy(x.grad), ops["lin/dx"], 3 * t, str('ops["lin/dx"]'))
    gclose(mode, npy(m.fc[0].weight.grad), ops["lin/dw"], 3 * t, str('ops["lin/dw"]'))
    gclose(mode, npy(m.fc[0].bias.grad), ops["lin/dbias"], 3 * t, str('ops["lin/dbias"]'))


def test_reparam_kl_golden(vp):
    ops = load("ops.npz")
    mu, lv = cu(ops["rp/mu"], True), cu(ops["rp/lv"], True)
    z, kl = vp.reparam_kl(mu, lv, eps=cu(ops["rp/eps"]))
    close(npy(z), ops["rp/z"], 1e-6, str('ops["rp/z"]'))
    close(npy(kl), ops["rp/kl"], 1e-6, str('ops["rp/kl"]'))
    (kl.sum() + (z * cu(ops["rp/dz"])).sum()).backward()
    close(npy(mu.grad), ops["rp/dmu"], 1e-6, str('ops["rp/dmu"]'))
    close(npy(lv.grad), ops["rp/dlv"], 1e-6, str('ops["rp/dlv"]'))
    # packed (mu | logvar) form used on the hot path
    packed = torch.cat([cu(ops["rp/mu"]), cu(ops["rp/lv"])], dim=1).requires_grad_(True)
    z2, kl2 = vp.reparam_kl(packed, None, eps=cu(ops["rp/eps"]))
    (kl2.sum() + (z2 * cu(ops["rp/dz"])).sum()).backward()
    assert torch.equal(z2, z) and torch.equal(kl2, kl)
    close(npy(packed.grad[:, :16]), ops["rp/dmu"], 1e-6, str('ops["rp/dmu"]'))
    close(npy(packed.grad[:, 16:]), ops["rp/dlv"], 1e-6, str('ops["rp/dlv"]'))


def test_losses_golden(vp):
    ops = load("ops.npz")
    x = cu(ops["ls/x"])
    for nm, fn in (("mse", vp.mse_loss), ("l1", vp.l1_loss)):
        xt = cu(ops["ls/xt"], True)
        l = fn(x, xt)
        l.backward()
        close(npy(l), ops[f"ls/{nm}"], 1e-6, str('ops[f"ls/{nm}"]'))
        close(npy(xt.grad), ops[f"ls/{nm}_dxt"], 1e-6, str('ops[f"ls/{nm}_dxt"]'))
    # the scratch accumulator must be clean for a second call
    xt = cu(ops["ls/xt"], True)
    close(npy(vp.mse_loss(x, xt)), ops["ls/mse"], 1e-6, str('ops["ls/mse"]'))
    logits = cu(ops["ls/logits"], True)
    l = vp.bce_dice_loss(logits, cu(ops["ls/t"]), 0.5)
    l.backward()
    close(npy(l), ops["ls/bce_dice"], 1e-6, str('ops["ls/bce_dice"]'))
    close(npy(logits.grad), ops["ls/bce_dice_dlogits"], 1e-5, str('ops["ls/bce_dice_dlogits"]'))


# ------------------------------------------------------------------------------------------------
# the whole VAE step against the golden fixtures and against the oracle on fresh seeds
# ------------------------------------------------------------------------------------------------
def build_vae(img, cin, z, seed):
    from vae_play_b200.models.networks import Decoder, Encoder
    L = int(math.log2(img // 8))
    P = vn.synth_vae_params(img, z, cin, cin, seed)
    enc = Encoder(channel_in=cin, z_size=z, iter_level=L)
    dec = Decoder(z_size=z, size=enc.size, channel_out=cin, iter_level=L)
    enc.load_state_dict({k[8:]: torch.from_numpy(v) for k, v in P.items() if k.startswith("encoder.")}, strict=False)
    dec.load_state_dict({k[8:]: torch.from_numpy(v) for k, v in P.items() if k.startswith("decoder.")}, strict=False)
    return enc.cuda().train(), dec.cuda().train(), P


def run_step(vp, enc, dec, x_np, eps_np):
    """One VAE step through the CUDA path.  Returns (outputs, grads, running stats, ReLU layer activations)."""
    import vae_play_b200.functional as VF
    x = cu(x_np)
    trace = VF.trace_activations(True)
    mulv = enc.forward_packed(x)
    z, kl = VF.reparam_kl(mulv, None, eps=cu(eps_np), z_dtype=VF.act_dtype())
    xt = VF.from_channels_last(dec.forward_cl(z.reshape(len(z), 1, 1, -1)))
    VF.trace_activations(False)
    loss = VF.vae_loss(x, xt, kl)
    enc.zero_grad()
    dec.zero_grad()
    loss.backward()
    zd = mulv.shape[1] // 2
    out = {"mu": npy(mulv[:, :zd]), "logvar": npy(mulv[:, zd:]), "z": npy(z), "x_tilde": npy(xt), "kl": npy(kl),
           "loss": npy(loss)}
    grads = {}
    running = {}
    for pref, m in (("encoder", enc), ("decoder", dec)):
        for k, p in m.named_parameters():
            grads[f"{pref}.{k}"] = npy(p.grad)
        for k, b in m.named_buffers():
            if "running" in k:
                running[f"{pref}.{k}"] = npy(b)
    # fused-layer call order: encoder.conv.0..L-1, encoder.fc, decoder.fc, decoder.conv.0..L-1, decoder output conv
    L = len(enc.conv)
    names = [f"encoder.conv.{i}" for i in range(L)] + ["encoder.fc", "decoder.fc"] + [f"decoder.conv.{i}" for i in range(L)]
    acts = {}
    for name, a in zip(names, trace):
        a = a.float().permute(0, 3, 1, 2).contiguous()          # NHWC -> NCHW (torch feature order when flattened)
        acts[name] = npy(a.reshape(len(a), -1) if name.endswith(".fc") else a)
    return out, grads, running, acts


def check_masks(acts, want_acts, mode, atol):
    """ReLU activation patterns of the CUDA forward vs the oracle: they may differ only where the oracle's
    pre-activation is within rounding noise of zero, and only for a vanishing fraction of elements.
    Returns the CUDA patterns (to pin the oracle's backward, see oracle.vae_numpy.act_bwd)."""
    noise, max_frac = (1e-5, 1e-5) if mode == "fp32" else (6e-2, 1e-2)
    masks = {}
    for name, a in acts.items():
        pre = want_acts[name + ".pre_act"]
        m = a > 0
        masks[name] = m
        bad = m != (pre > 0)
        frac = bad.mean()
        assert frac <= max_frac, f"{name}: {frac:.2e} of the ReLU pattern differs"
        if bad.any():
            worst = np.abs(pre[bad]).max() / np.abs(pre).max()
            assert worst <= noise, f"{name}: pattern differs at |pre-activation| = {worst:.2e} of max (noise floor {noise:.0e})"
        close(a, want_acts[name + ".out"], atol, name + ".out")
    return masks


def step_tolerances(mode, batch):
    """(forward tol, activation tol, gradient tol), max|a-b|/max|b| per tensor, for the END-TO-END step.

    fp32 check mode: 1e-5 on outputs; 2e-5 on inner activations (BatchNorm1d over 4 samples at 128x128);
    5e-5 on parameter gradients (measured <= 3.9e-5: cancelling sums of ~1e6 fp32 products behind BatchNorm;
    the reference's own fp32-vs-fp64 deviation there is 1.2e-5..1.7e-5, tests/golden ref_fp32_dev_*).
    bf16 mode: every stored activation/gradient is rounded to 8 mantissa bits (2^-9 relative); ONE layer on
    identical inputs stays below the north-star 1e-2 (the *_golden op tests), but the 9-layer chain with a
    BatchNorm rescaling after each layer accumulates ~24 such roundings: measured 1.0e-2..1.8e-2 on mu/logvar
    and 1.7e-2..3.2e-2 on gradients at batch 6..32, larger when BatchNorm1d normalises over only 4 samples.
    Gradients are compared with the ReLU pattern pinned (check_masks): a flipped ReLU is an O(1) local
    change of the gradient that no tolerance can absorb."""
    if mode == "fp32":
        return TOL_FP32, 2e-5, 5e-5
    if batch >= 8:
        return 2.5e-2, 3e-2, 5e-2
    return 3e-2, 1.2e-1, 2.5e-1


@pytest.mark.parametrize("case", ["vae64_c1_b4", "vae64_c3_b4", "vae128_c1_b4"])
def test_vae_step_golden(vp, mode, case):
    """Forward outputs, loss and BatchNorm running statistics against the reference-generated fixtures."""
    g = load(case + ".npz")
    img, cin, b, z, seed = [int(v) for v in g["meta"]]
    enc, dec, _ = build_vae(img, cin, z, seed)
    x, eps = vn.synth_batch(b, img, cin, z, seed)
    out, grads, running, _ = run_step(vp, enc, dec, x, eps)
    dev = ref_dev(g)
    ft, at, gt = step_tolerances(mode, b)
    for key in ("mu", "logvar", "z", "x_tilde", "kl", "loss"):
        close(out[key], g[key], max(ft, 3 * dev.get(key, 0.0)), key)
    for key in g.files:
        if key.startswith("running/"):
            close(running[key[8:]], g[key], max(ft, 1e-5), key)
        if key.startswith("grad/"):
            # digest = (sum, l2 norm, max, 64 samples): the norm must agree; element-wise gradient parity is
            # test_vae_step_vs_oracle's job (ReLU pattern pinned), the reference's own fp32 run flips ReLUs too
            got, want = digest(grads[key[5:]]), g[key]
            assert abs(got[1] - want[1]) / want[1] < (5e-2 if mode == "fp32" else 0.3), key


@pytest.mark.parametrize("img,cin,b,seed", [(64, 1, 4, 0), (64, 1, 8, 5), (64, 3, 6, 6), (64, 1, 32, 7), (128, 1, 4, 2)])
def test_vae_step_vs_oracle(vp, mode, img, cin, b, seed):
    """Per-layer activations, outputs, loss, running statistics and every parameter gradient vs the oracle."""
    z = 128
    enc, dec, P = build_vae(img, cin, z, seed)
    x, eps = vn.synth_batch(b, img, cin, z, seed)
    out, grads, running, acts = run_step(vp, enc, dec, x, eps)
    ft, at, gt = step_tolerances(mode, b)
    fwd = vn.vae_step(P, x, eps)
    masks = check_masks(acts, fwd["acts"], mode, at)
    want = vn.vae_step(P, x, eps, masks=masks)
    for key in ("mu", "logvar", "z", "x_tilde", "kl", "loss"):
        close(out[key], want[key], ft, key)
    for key, gref in want["grads"].items():
        close(grads[key], gref, gt, key)
    for key, rref in want["running"].items():
        close(running[key], rref, max(ft, 1e-5), key)


# ------------------------------------------------------------------------------------------------
# size-independent properties at the benchmark size (B=256, 64x64): adjointness of fwd/dgrad/wgrad,
# normalisation invariants, loss identities
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,cin,cout,hw", [("conv", 64, 128, 32), ("convT", 256, 128, 16), ("conv", 1, 64, 64), ("conv_out", 64, 1, 64),
                                              ("linear", 16384, 1024, 1), ("linear", 128, 16384, 1)])
def test_adjoint_identities_full_size(vp, mode, kind, cin, cout, hw):
    import vae_play_b200.functional as VF
    torch.manual_seed(3)
    B = 256
    dt = VF.act_dtype()
    if kind == "conv":
        layer = VF.TapLayer("conv", cin, cout, k=5, stride=2, pad=2)
        w = torch.randn(cout, cin, 5, 5, device="cuda") * 0.05
    elif kind == "convT":
        layer = VF.TapLayer("convT", cin, cout, k=5, stride=2, pad=2, out_pad=1)
        w = torch.randn(cin, cout, 5, 5, device="cuda") * 0.05
    elif kind == "conv_out":
        layer = VF.TapLayer("conv", cin, cout, k=5, stride=1, pad=2)
        w = torch.randn(cout, cin, 5, 5, device="cuda") * 0.05
    else:
        layer = VF.TapLayer("linear", cin, cout)
        w = torch.randn(cout, cin, device="cuda") * 0.05
    if w.dim() == 4 and cin % 64 == 0 and cout % 64 == 0:
        w = w.contiguous(memory_format=torch.channels_last)      # the in-place (no packing) route in bf16 mode
    x = torch.randn(B, hw, hw, cin, device="cuda").to(dt)
    y = layer.fwd(x, w, None)
    dy = torch.randn(y.shape, device="cuda").to(dt)
    dx = layer.dgrad(dy, w, tuple(x.shape))
    dw = layer.wgrad(x, dy, w)
    # quantise w the way the kernels see it so that the identities are exact up to accumulation error
    wq = w.to(dt).double()
    a = (y.double() * dy.double()).sum().item()
    b_ = (x.double() * dx.double()).sum().item()
    c = (wq * dw.double()).sum().item()
    scale = (y.double().abs() * dy.double().abs()).sum().item()
    t = 2e-3 if mode == "bf16" else 1e-6
    assert abs(a - b_) / scale < t, (a, b_, scale)
    assert abs(a - c) / scale < t, (a, c, scale)


def test_batchnorm_invariants_full_size(vp, mode):
    import vae_play_b200.functional as VF
    torch.manual_seed(4)
    dt = VF.act_dtype()
    B, H, Cn = 256, 32, 64
    x = (torch.randn(B, H, H, Cn, device="cuda") * 3 + 1.5).to(dt)
    layer = VF.TapLayer("conv", Cn, Cn, k=1)
    w = torch.eye(Cn, device="cuda").reshape(Cn, Cn, 1, 1).contiguous()
    bn = torch.nn.BatchNorm2d(Cn).cuda().train()
    a, y = VF.fused_layer(x, w, None, bn.weight, bn.bias, layer, VF.NormCfg("batch"), "none", 0.0, True, bn)
    af = a.double().reshape(-1, Cn)
    assert af.mean(0).abs().max().item() < (2e-2 if mode == "bf16" else 1e-4)
    assert (af.var(0, unbiased=False) - 1).abs().max().item() < (2e-2 if mode == "bf16" else 1e-3)
    xf = x.double().reshape(-1, Cn)
    close(npy(bn.running_mean), 0.1 * xf.mean(0).cpu().numpy(), (1e-2 if mode == "bf16" else 1e-5), str('0.1 * xf.mean(0).cpu().numpy()'))


def test_abi_error_codes(vp):
    from vae_play_b200 import _lib
    lib = _lib.load()
    g = _lib.VpConvGeom(2, 8, 8, 4, 5, 5, 4, 3, 3, 1, 1, 0)  # inconsistent output size
    buf = torch.zeros(1024, device="cuda")
    p = C.c_void_p(buf.data_ptr())
    rc = lib.vp_conv_fwd(C.byref(g), p, p, None, p, 0, 0, 0, 0.0, 0, None)
    assert rc == -1 and b"inconsistent" in lib.vp_last_error()
    rc = lib.vp_conv_fwd(None, p, p, None, p, 0, 0, 0, 0.0, 0, None)
    assert rc == -1
    g2 = _lib.VpConvGeom(2, 8, 8, 4, 8, 8, 4, 3, 3, 1, 1, 0)
    rc = lib.vp_conv_fwd(C.byref(g2), p, p, None, p, 0, 0, 0, 0.0, 2, None)  # TC engine with fp32 dtype
    assert rc == -3
    with pytest.raises(_lib.VaePlayError):
        vp.mse_loss(torch.zeros(4), torch.zeros(4))  # CPU tensors: no CPU path


# ------------------------------------------------------------------------------------------------
# tcgen05 engine vs the CUDA-core engine on identical bf16 operands (fp32 outputs compared)
# ------------------------------------------------------------------------------------------------
TC_SHAPES = [
    # kind, cin, cout, hw, batch
    ("conv", 64, 128, 32, 4), ("conv", 128, 256, 16, 4), ("conv", 64, 128, 32, 3),
    ("convT", 256, 256, 8, 4), ("convT", 256, 128, 16, 2), ("convT", 128, 64, 32, 2),
    ("linear", 16384, 1024, 1, 16), ("linear", 128, 16384, 1, 16), ("linear", 1024, 256, 1, 40),
    ("conv_s1", 64, 1, 64, 2), ("conv_s1", 64, 3, 32, 2), ("conv_k3", 64, 96, 20, 3),
    # channels-last weights: forward / dgrad / wgrad on the module's own weight (K-major and MN-major operands), no packing
    ("conv_cl", 64, 128, 32, 4), ("conv_cl", 128, 256, 16, 3), ("convT_cl", 256, 256, 8, 4), ("convT_cl", 256, 128, 16, 2),
    ("convT_cl", 128, 64, 32, 2), ("conv_s1_cl", 64, 128, 24, 2), ("conv_k3_cl", 128, 64, 20, 3),
    # ragged grids through the stride-2 sub-lattice window kernel (forward of a conv, data gradient of a transposed conv)
    ("conv_cl", 64, 128, 26, 3), ("conv_cl", 128, 64, 50, 2), ("convT_cl", 128, 64, 13, 2), ("convT_cl", 64, 128, 27, 2),
]


def _tc_layer(kind, cin, cout):
    import vae_play_b200.functional as VF
    g = torch.Generator(device="cuda").manual_seed(11)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g) * 0.05
    if kind.endswith("_cl"):
        layer, w = _tc_layer(kind[:-3], cin, cout)
        return layer, w.contiguous(memory_format=torch.channels_last)
    if kind == "conv":
        return VF.TapLayer("conv", cin, cout, k=5, stride=2, pad=2), r(cout, cin, 5, 5)
    if kind == "conv_s1":
        return VF.TapLayer("conv", cin, cout, k=5, stride=1, pad=2), r(cout, cin, 5, 5)
    if kind == "conv_k3":
        return VF.TapLayer("conv", cin, cout, k=3, stride=1, pad=1), r(cout, cin, 3, 3)
    if kind == "convT":
        return VF.TapLayer("convT", cin, cout, k=5, stride=2, pad=2, out_pad=1), r(cin, cout, 5, 5)
    return VF.TapLayer("linear", cin, cout), r(cout, cin)


@pytest.mark.parametrize("kind,cin,cout,hw,b", TC_SHAPES)
def test_tc_engine_matches_simt(vp, kind, cin, cout, hw, b):
    import vae_play_b200.functional as VF
    vp.set_precision("bf16")
    layer, w = _tc_layer(kind, cin, cout)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(cout, device="cuda", generator=g) if kind.startswith("conv_") else None
    try:
        vp.set_engine("simt")
        y_ref = layer.fwd(x, w, bias, out_dtype=torch.float32)
        dy = torch.randn(y_ref.shape, device="cuda", generator=g).to(torch.bfloat16)
        dx_ref = layer.dgrad(dy, w, tuple(x.shape), out_dtype=torch.float32)
        dw_ref = layer.wgrad(x, dy, w)
        vp.set_engine("tc")
        layer._cache.clear()
        y = layer.fwd(x, w, bias, out_dtype=torch.float32)
        y16 = layer.fwd(x, w, bias, "relu")
        torch.cuda.synchronize()
        close(npy(y), npy(y_ref), 1e-4, f"{kind} fwd tc-vs-simt")
        close(npy(y16), npy(torch.relu(y_ref).to(torch.bfloat16)), 8e-3, f"{kind} fwd bf16+relu epilogue")
        if cin % 64 == 0 and cout % 64 == 0:
            dx = layer.dgrad(dy, w, tuple(x.shape), out_dtype=torch.float32)
            torch.cuda.synchronize()
            close(npy(dx), npy(dx_ref), 1e-4, f"{kind} dgrad tc-vs-simt")
            dw = layer.wgrad(x, dy, w)
            torch.cuda.synchronize()
            close(npy(dw), npy(dw_ref), 1e-4, f"{kind} wgrad tc-vs-simt")
    finally:
        vp.set_engine("auto")


@pytest.mark.parametrize("cin,cout,stride,hw,b", [(1, 64, 2, 64, 3), (64, 1, 1, 32, 3), (3, 64, 2, 32, 2), (64, 3, 1, 16, 2),
                                                    (1, 64, 2, 20, 2), (64, 1, 1, 10, 5)])
def test_thin_channel_wgrad(vp, cin, cout, stride, hw, b):
    """First/last-layer weight gradients (one side has <= 4 channels) use a dedicated streaming kernel."""
    import vae_play_b200.functional as VF
    vp.set_precision("bf16")
    vp.set_engine("auto")
    g = torch.Generator(device="cuda").manual_seed(2)
    layer = VF.TapLayer("conv", cin, cout, k=5, stride=stride, pad=2)
    w = torch.randn(cout, cin, 5, 5, device="cuda", generator=g) * 0.1
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g).to(torch.bfloat16)
    ho = (hw + 4 - 5) // stride + 1
    dy = torch.randn(b, ho, ho, cout, device="cuda", generator=g).to(torch.bfloat16)
    dw = layer.wgrad(x, dy, w)
    want = torch.nn.grad.conv2d_weight(x.double().permute(0, 3, 1, 2), w.shape, dy.double().permute(0, 3, 1, 2),
                                       stride=stride, padding=2)
    close(npy(dw), npy(want), 1e-5, f"thin wgrad {cin}->{cout}")


@pytest.mark.parametrize("cin,cout,stride,hw,b", [(1, 64, 2, 64, 3), (3, 64, 2, 20, 2), (64, 1, 1, 32, 3), (64, 3, 1, 10, 2)])
def test_thin_channel_fwd_and_dgrad(vp, cin, cout, stride, hw, b):
    """First-layer forward (Cin <= 4) and last-layer data gradient (Cout <= 4) use the thin-K streaming kernel."""
    import torch.nn.functional as F
    import vae_play_b200.functional as VF
    vp.set_precision("bf16")
    vp.set_engine("auto")
    g = torch.Generator(device="cuda").manual_seed(3)
    layer = VF.TapLayer("conv", cin, cout, k=5, stride=stride, pad=2)
    w = torch.randn(cout, cin, 5, 5, device="cuda", generator=g) * 0.1
    wq = w.to(torch.bfloat16).double()
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(cout, device="cuda", generator=g)
    if cin <= 4:
        y = layer.fwd(x, w, bias, out_dtype=torch.float32)
        want = F.conv2d(x.double().permute(0, 3, 1, 2), wq, bias.double(), stride=stride, padding=2).permute(0, 2, 3, 1)
        close(npy(y), npy(want), 1e-5, "thin fwd")
        y16 = layer.fwd(x, w, bias, "relu")
        close(npy(y16), npy(torch.relu(want)), 6e-3, "thin fwd bf16 relu")
    else:
        ho = (hw + 4 - 5) // stride + 1
        dy = torch.randn(b, ho, ho, cout, device="cuda", generator=g).to(torch.bfloat16)
        dx = layer.dgrad(dy, w, tuple(x.shape), out_dtype=torch.float32)
        want = torch.nn.grad.conv2d_input((b, cin, hw, hw), wq, dy.double().permute(0, 3, 1, 2), stride=stride, padding=2)
        close(npy(dx), npy(want.permute(0, 2, 3, 1)), 1e-5, "thin dgrad")


@pytest.mark.parametrize("cin,cout,k,stride,hw,b", [(1, 64, 5, 2, 64, 5), (1, 64, 7, 3, 40, 3), (1, 32, 5, 1, 21, 2), (1, 128, 3, 1, 17, 2),
                                                      (64, 1, 5, 1, 64, 5), (128, 1, 5, 1, 19, 3), (64, 2, 3, 1, 33, 2), (64, 1, 3, 1, 9, 2)])
def test_thin_tc_kernels(vp, cin, cout, k, stride, hw, b):
    """tcgen05 thin-layer kernels (operands built in shared memory from the fp32 master weight) against float64 torch:
    forward with bias + sigmoid, data gradient and weight gradient, ragged grids and both thin sides."""
    import torch.nn.functional as F
    import vae_play_b200.functional as VF
    vp.set_precision("bf16")
    vp.set_engine("auto")
    pad = (k - 1) // 2
    g = torch.Generator(device="cuda").manual_seed(11)
    layer = VF.TapLayer("conv", cin, cout, k=k, stride=stride, pad=pad)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) * 0.1
    wq = w.to(torch.bfloat16).double()
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(cout, device="cuda", generator=g)
    ho = (hw + 2 * pad - k) // stride + 1
    dy = torch.randn(b, ho, ho, cout, device="cuda", generator=g).to(torch.bfloat16)
    xd, dyd = x.double().permute(0, 3, 1, 2), dy.double().permute(0, 3, 1, 2)
    n0 = vp._lib.launch_count()
    assert layer._thin("fwd", torch.bfloat16, w)
    y = layer.fwd(x, w, bias, "sigmoid", out_dtype=torch.float32)
    want = torch.sigmoid(F.conv2d(xd, wq, bias.double(), stride=stride, padding=pad)).permute(0, 2, 3, 1)
    close(npy(y), npy(want), 1e-5, "thin tc fwd")
    y16 = layer.fwd(x, w, None)
    want = F.conv2d(xd, wq, None, stride=stride, padding=pad).permute(0, 2, 3, 1)
    close(npy(y16), npy(want), 6e-3, "thin tc fwd bf16")
    assert vp._lib.launch_count() - n0 == 2          # one kernel each: no packing passes
    if layer._thin("dgrad", torch.bfloat16, w):
        dx = layer.dgrad(dy, w, tuple(x.shape), out_dtype=torch.float32)
        want = torch.nn.grad.conv2d_input((b, cin, hw, hw), wq, dyd, stride=stride, padding=pad)
        close(npy(dx), npy(want.permute(0, 2, 3, 1)), 1e-5, "thin tc dgrad")
    else:
        assert cin == 1 or cout > 1
    if layer._thin("wgrad", torch.bfloat16, w):
        dw = layer.wgrad(x, dy, w)
        want = torch.nn.grad.conv2d_weight(xd, w.shape, dyd, stride=stride, padding=pad)
        close(npy(dw), npy(want), 1e-5, "thin tc wgrad")
    else:
        assert (cout % 64 != 0 and cin == 1) or cout > 1


@pytest.mark.parametrize("kind,cin,cout,k,S", [("conv", 64, 128, 5, 1), ("conv", 3, 64, 5, 1), ("convT", 128, 64, 5, 1),
                                                ("conv", 40, 96, 3, 1), ("linear", 96, 40, 1, 1)])
def test_pack_unpack_bit_exact(vp, kind, cin, cout, k, S):
    """Weight packing is index shuffling: bit-exact against the host emulation, and unpack inverts it."""
    import vae_play_b200.functional as VF
    from vae_play_b200 import _lib
    from tests.test_host_cpu import _emulate_pack
    if kind == "conv":
        layer, shape = VF.TapLayer("conv", cin, cout, k=k, stride=1, pad=k // 2), (cout, cin, k, k)
    elif kind == "convT":
        layer, shape = VF.TapLayer("convT", cin, cout, k=k, stride=2, pad=2, out_pad=1), (cin, cout, k, k)
    else:
        layer, shape = VF.TapLayer("linear", cin, cout), (cout, cin)
    w = torch.randn(*shape, device="cuda")
    for which in ("fwd", "dgrad"):
        p = layer._recipe(which, w)
        wp = torch.empty(p.taps * p.n * p.k, dtype=torch.float32, device="cuda")
        _lib.call("vp_pack_weight", VF._ptr(w), VF._ptr(wp), 0, p.taps, p.n, p.k, p.sn, p.sk, p.st, VF._stream())
        want = _emulate_pack(w.cpu().numpy(), p)
        assert np.array_equal(wp.cpu().numpy().reshape(want.shape), want), (kind, which)
        back = torch.zeros_like(w)
        _lib.call("vp_unpack_wgrad", VF._ptr(wp), VF._ptr(back), p.taps, p.n, p.k, p.sn, p.sk, p.st, VF._stream())
        assert torch.equal(back, w), (kind, which)


def test_fused_rmsprop_matches_torch(vp):
    """FusedRMSprop == torch.optim.RMSprop (the reference's optimiser, train.py:136-140) step for step."""
    from vae_play_b200.optim import FusedRMSprop
    torch.manual_seed(0)
    shapes = [(64, 1, 5, 5), (128,), (1024, 16384), (3,), (256, 128, 5, 5), (7, 13)]
    pa = [torch.randn(*s, device="cuda").requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = FusedRMSprop(pa, lr=1e-4)
    ob = torch.optim.RMSprop(pb, lr=1e-4)
    for step in range(3):
        for x, y in zip(pa, pb):
            g = torch.randn_like(x) * (0.1 + step)
            x.grad = g.clone()
            y.grad = g.clone()
        v0 = pa[0]._version
        oa.step()
        ob.step()
        assert pa[0]._version > v0
        for x, y in zip(pa, pb):
            close(npy(x), npy(y), 1e-5, "rmsprop param")
            close(npy(oa.state[x]["square_avg"]), npy(ob.state[y]["square_avg"]), 1e-5, "rmsprop state")


@pytest.mark.parametrize("kind,cin,cout,hw,b", [("enc", 64, 128, 32, 6), ("enc", 128, 256, 16, 5), ("enc", 1, 64, 64, 7), ("enc", 64, 128, 20, 3),
                                                 ("dec", 128, 64, 32, 4), ("dec", 256, 256, 8, 9), ("dec", 256, 128, 16, 3), ("dec", 128, 64, 13, 2)])
def test_epilogue_batchnorm_statistics(vp, kind, cin, cout, hw, b):
    """BatchNorm statistics taken in the GEMM epilogue (from the bf16 values being stored, per-CTA partial sums): the
    running statistics match float64 statistics of the stored pre-norm tensor, and activations / gradients match the
    path with a separate statistics pass."""
    import vae_play_b200.functional as VF
    from vae_play_b200.models.networks import DecoderBlock, EncoderBlock, _bn_cfg
    vp.set_precision("bf16")
    vp.set_engine("auto")
    torch.manual_seed(5)
    blk = (EncoderBlock if kind == "enc" else DecoderBlock)(cin, cout).cuda().train()
    with torch.no_grad():
        blk.bn.weight.uniform_(0.5, 1.5)
        blk.bn.bias.uniform_(-0.5, 0.5)
    x = torch.randn(b, hw, hw, cin, device="cuda").to(torch.bfloat16)
    res = {}
    try:
        for on in (False, True):
            vp.set_epilogue_stats(on)
            blk.bn.running_mean.zero_(); blk.bn.running_var.fill_(1.0)
            blk.zero_grad(set_to_none=True)
            xin = x.clone().requires_grad_(True)
            n0 = vp._lib.launch_count()
            a, y = VF.fused_layer(xin, blk.conv.weight, None, blk.bn.weight, blk.bn.bias, blk._layer, _bn_cfg(blk.bn), "relu", 0.0, True, blk.bn)
            launches = vp._lib.launch_count() - n0
            a.float().square().sum().backward()
            yd = y.detach().double().reshape(-1, cout)
            m = yd.shape[0]
            want_rm = 0.9 * yd.mean(0)                                           # momentum 0.9 from running_mean = 0
            want_rv = 0.1 + 0.9 * yd.var(0, unbiased=False) * m / (m - 1)
            res[on] = (npy(a), npy(blk.bn.running_mean), npy(blk.bn.running_var), npy(blk.conv.weight.grad), npy(blk.bn.weight.grad), launches,
                       want_rm.cpu().numpy(), want_rv.cpu().numpy())
    finally:
        vp.set_epilogue_stats(True)
    assert blk._layer.stats_in_epilogue(torch.bfloat16, blk.conv.weight)
    if y.numel() // cout > 8192:                             # (smaller tensors take the single-launch few-rows kernel when the epilogue path is off)
        assert res[True][5] < res[False][5]                  # the statistics pass is gone
    for on in (False, True):
        # the batch mean is a cancelling sum (|mean| << std): bound its error by the std, i.e. by sqrt(running_var)
        assert np.max(np.abs(res[on][1] - res[on][6]) / np.sqrt(res[on][7])) < 2e-6, ("running_mean", on)
        close(res[on][2], res[on][7], 2e-6, f"running_var (epilogue={on})")
    close(res[True][0], res[False][0], 1e-2, "activations")  # one bf16 ulp where a rounding boundary moves
    assert rel_l2(res[True][3], res[False][3]) < 5e-3 and rel_l2(res[True][4], res[False][4]) < 5e-3


def test_persistent_grads_and_zeroing_optimizer(vp):
    """persistent_grads + FusedRMSprop(zero_grads=True): the weight-gradient kernels accumulate into slots the optimiser
    cleared (no memset) and autograd adopts the slots as .grad (no clone); the bf16 operand copies are refreshed by the
    optimiser kernel.  Gradients agree with the plain flow (fresh tensors, memset, separate cast pass)."""
    import copy
    import vae_play_b200.functional as VF
    from vae_play_b200.models.networks import VaeGan
    from vae_play_b200.optim import FusedRMSprop
    vp.set_precision("bf16")
    vp.set_engine("auto")
    # (1) one layer, identical operands: slot + accumulate == fresh tensor + memset, up to the order of the fp32 atomics
    g = torch.Generator(device="cuda").manual_seed(3)
    layer = VF.TapLayer("convT", 128, 64, k=5, stride=2, pad=2, out_pad=1)
    w = torch.nn.Parameter((torch.randn(128, 64, 5, 5, device="cuda", generator=g) * 0.05).contiguous(memory_format=torch.channels_last))
    xa = torch.randn(4, 16, 16, 128, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(4, 32, 32, 64, device="cuda", generator=g).to(torch.bfloat16)
    try:
        VF.set_grad_sinks({})
        plain = layer.wgrad(xa, dy, w)
        flat = VF.persistent_grads([w])
        n0 = vp._lib.launch_count()
        slot = layer.wgrad(xa, dy, w)
        assert slot.data_ptr() == flat.data_ptr() and slot.stride() == w.stride()
        assert VF._GRAD_SINKS[w.data_ptr()][3] is False            # handed out: no longer known to be zero
        close(npy(slot), npy(plain), 1e-5, "wgrad into a persistent slot")
        again = layer.wgrad(xa, dy, w)                              # slot not cleared -> must be memset, not accumulated
        close(npy(again), npy(plain), 1e-5, "wgrad into a dirty slot")
    finally:
        VF.set_grad_sinks({})
    # (2) three training steps of the whole model
    torch.manual_seed(0)
    ref = VaeGan(64, 128).cuda().train()
    x = torch.rand(8, 1, 64, 64, device="cuda")
    eps = torch.randn(8, 128, device="cuda")
    first = []
    for persistent in (False, True):
        VF.set_grad_sinks({})
        m = copy.deepcopy(ref)
        params = list(m.encoder.parameters()) + list(m.decoder.parameters())
        opt = FusedRMSprop(params, lr=1e-6, zero_grads=persistent)
        flat = VF.persistent_grads(params) if persistent else None
        big = [p for p in params if p.dim() == 4 or p.numel() > 1_000_000]      # every conv / fc weight (9 tensors, 99.9 % of the parameters)
        for step in range(3):
            opt.zero_grad(set_to_none=True)
            if persistent and step > 0:
                assert float(flat.abs().max()) == 0.0                            # cleared by the optimiser kernel
                assert all(VF._GRAD_SINKS[p.data_ptr()][3] for p in big)         # -> the wgrad kernels skip their memset
            xt, mulv, kl = m.vae_forward(x, eps=eps)
            VF.vae_loss(x, xt, kl).backward()
            if step == 0:
                first.append([npy(p.grad) for p in params])
            if persistent:
                # the gradients were written into the slots and adopted by autograd as they are (no clone, no copy back)
                assert all(p.grad.data_ptr() == flat.data_ptr() + 4 * VF._GRAD_SINKS[p.data_ptr()][1] for p in big)
            v = [p._version for p in big]
            opt.step()
            for p, v0 in zip(big, v):
                lay_sh = VF._SHADOWS.get(p.data_ptr())
                if lay_sh is not None:                                           # bf16 copy refreshed by the optimiser kernel, bit for bit
                    assert p._version > v0 and torch.equal(lay_sh[1], p.detach().to(torch.bfloat16))
    VF.set_grad_sinks({})
    assert len(big) == 9
    for a, b in zip(*first):
        # same weights, same inputs; bf16 training at batch 8 is only reproducible up to a few ReLU / rounding flips
        # (two runs of the SAME flow differ by up to ~1e-2 here): this only guards against gross errors, (1) is the exact check
        assert rel_l2(a, b) < 0.1


def test_two_stage_backward_matches_single(vp):
    """bench.py's data-parallel step cuts the backward at the output of the encoder's conv stack (so that the gradient
    all-reduce of everything downstream can overlap the conv-stack backward): same gradients as one backward call through
    the SAME forward graph, with the weight gradients written in place into persistent slots as in bench.py, and no kernel
    launched twice (the cut sits behind VF.grad_cut: naming a block's own output in ``inputs`` would run its backward in
    both stages)."""
    import vae_play_b200.functional as VF
    from vae_play_b200 import _lib
    from vae_play_b200.models.networks import VaeGan
    vp.set_precision("bf16")
    vp.set_engine("auto")
    VF.set_fuse_bn_backward(False)        # fp32 shared-memory atomics in the fused reduction: not bit-comparable run to run
    torch.manual_seed(1)
    m = VaeGan(64, 128).cuda().train()
    x = torch.rand(16, 1, 64, 64, device="cuda")
    eps = torch.randn(16, 128, device="cuda")
    params = list(m.encoder.parameters()) + list(m.decoder.parameters())
    conv_params = [p for blk in m.encoder.conv for p in blk.parameters()]
    ids = {id(p) for p in conv_params}
    try:
        flat = VF.persistent_grads(params)
        taps, cut = [], {}
        xt, mulv, kl = m.vae_forward(x, eps=eps, taps=taps)
        loss = VF.vae_loss(x, xt, kl)
        a3 = taps[0]
        a3.register_hook(lambda g: cut.__setitem__("g", g))
        n0 = _lib.launch_count()
        loss.backward(inputs=[p for p in params if id(p) not in ids] + [a3], retain_graph=True)
        assert all(p.grad is None for p in conv_params) and all(p.grad is not None for p in params if id(p) not in ids)
        n_torch = []
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            taps[1].backward(cut["g"], inputs=conv_params, retain_graph=True)
            torch.cuda.synchronize()
        n_torch = [e.key for e in prof.key_averages() if "at::native" in e.key or "Memcpy" in e.key]
        assert not n_torch, f"stage 2 launched library kernels (a retained .grad being cloned / added into?): {n_torch}"
        n_staged = _lib.launch_count() - n0
        slot = lambda p: flat.data_ptr() <= p.grad.data_ptr() < flat.data_ptr() + 4 * flat.numel()
        names = {id(p): k for k, p in m.named_parameters()}
        outside = [names[id(p)] for p in params if p.dim() > 1 and not slot(p)]
        assert not [k for k in outside if "l_mu" not in k and "l_var" not in k], f"weight gradients that left their persistent slots: {outside}"
        staged = [npy(p.grad) for p in params]
        VF.set_grad_sinks({})
        for p in params:
            p.grad = None
        n0 = _lib.launch_count()
        loss.backward()
        n_single = _lib.launch_count() - n0
        for a, p in zip(staged, params):
            assert rel_l2(a, npy(p.grad)) < 1e-4      # same kernels on the same saved activations; only the order of fp32 atomics differs
        assert n_staged <= n_single, (n_staged, n_single)     # the single pass zero-fills fresh gradient buffers the slots do not need
    finally:
        VF.set_fuse_bn_backward(True)
        VF.set_grad_sinks({})


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_discriminator_golden(vp, prec):
    """VAE-GAN Discriminator (models/networks.py:151-195, config 4): 'REC' features, 'GAN' probabilities and parameter
    gradients against the fixture written by oracle/gen_golden_disc.py from the unmodified reference in float64."""
    import os
    from oracle.gen_golden_disc import digest, synth_disc_inputs, synth_disc_params
    from vae_play_b200.models.networks import Discriminator
    want = np.load(os.path.join(os.path.dirname(__file__), "golden", "disc64_b2.npz"))
    vp.set_precision(prec)
    vp.set_engine("auto")
    try:
        d = Discriminator(channel_in=1, recon_level=3, iter_level=3)
        P = synth_disc_params(0)
        d.load_state_dict({k: torch.from_numpy(v) for k, v in P.items()}, strict=False)
        d = d.cuda().train()
        xs, probe = synth_disc_inputs(0)
        xt = [torch.from_numpy(x).cuda() for x in xs]
        rec = d(*xt, "REC")
        assert tuple(rec.shape) == tuple(want["rec_shape"])
        d.zero_grad()
        gan = d(*xt, "GAN")
        (gan * torch.from_numpy(probe).cuda()).sum().backward()
        tol = 2e-5 if prec == "fp32" else 2.5e-2
        close(digest(npy(rec))[3:], want["rec_digest"][3:], tol, "REC features (strided samples)")
        close(digest(npy(rec))[1:2], want["rec_digest"][1:2], tol, "REC features (l2)")
        close(npy(gan), want["gan"], tol, "GAN probabilities")
        if prec == "fp32":
            close(digest(npy(d.conv[0][0].weight.grad))[3:], want["grad_conv0"][3:], 2e-4, "grad conv.0")
            close(digest(npy(d.conv[2].conv.weight.grad))[3:], want["grad_conv2"][3:], 2e-4, "grad conv.2")
            close(npy(d.fc[3].weight.grad), want["grad_fc3"], 2e-4, "grad fc.3")
        else:
            assert rel_l2(npy(d.fc[3].weight.grad), want["grad_fc3"]) < 0.1
    finally:
        vp.set_precision("bf16")


def test_bf16_wire_pack_and_optimizer(vp):
    """Data-parallel exchange in bf16: GradBuckets.pack() writes bf16(grad) into the wire buffer and clears the fp32 slots in
    the same pass; FusedRMSprop(wire=...) then gives exactly torch.optim.RMSprop on the bf16-rounded gradients."""
    import vae_play_b200.functional as VF
    from vae_play_b200.optim import FusedRMSprop
    from vae_play_b200.parallel import GradBuckets
    torch.manual_seed(0)
    shapes = [(64, 64, 3, 3), (130,), (257, 33), (7,)]
    ps = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    ps[0].data = ps[0].data.contiguous(memory_format=torch.channels_last)
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    gb = GradBuckets(ps, world_size=1, bucket_mb=0.05, overlap=False, wire_dtype=torch.bfloat16)
    try:
        assert len(gb.buckets) >= 2
        views = {}
        for bi in range(len(gb.buckets)):
            views.update(gb.wire_views(bi))
        opt = FusedRMSprop(ps, lr=1e-3, zero_grads=False, wire=views)
        ref = torch.optim.RMSprop(qs, lr=1e-3)
        for step in range(3):
            grads = [torch.randn_like(p) for p in ps]
            for p, q, g in zip(ps, qs, grads):
                slot = gb.slot[id(p)][1]
                slot.copy_(g)
                p.grad = slot
                q.grad = g.to(torch.bfloat16).float()
            gb.pack()
            for p, g in zip(ps, grads):
                assert torch.equal(views[id(p)].float(), g.to(torch.bfloat16).float())         # bf16(grad), element for element (strided views too)
                assert float(p.grad.abs().max()) == 0.0                                       # fp32 slot cleared by the same pass
            opt.step()
            ref.step()
            for p, q in zip(ps, qs):
                close(npy(p), npy(q), 1e-6, f"RMSprop on wire gradients, step {step}")
    finally:
        gb.remove()


@pytest.mark.parametrize("thin", [False, True])
def test_fused_bn_backward_matches_unfused(vp, thin):
    """BatchNorm backward, first pass: done by the epilogue of the consumer's data-gradient kernel (vp_conv_dgrad_cl_bnred, taken
    at the bench size by the stride-2 sub-lattice kernel for decoder.conv.1 <- decoder.conv.2; with ``thin`` also
    vp_thin_conv_dgrad_bnred for decoder.conv.2 <- the output layer) == the separate reduce pass.  Same operands, same dx; only
    the summation (per-CTA fp32 partials, then double) differs."""
    import copy
    import vae_play_b200.functional as VF
    from vae_play_b200 import _lib
    from vae_play_b200.models.networks import VaeGan
    vp.set_precision("bf16")
    vp.set_engine("auto")
    VF.set_grad_sinks({})
    torch.manual_seed(0)
    m = VaeGan(64, 128).cuda().train()
    x = torch.rand(256, 1, 64, 64, device="cuda")
    eps = torch.randn(256, 128, device="cuda")
    res, launches = [], []
    # ONE forward, two backward passes over the same saved activations (the forward's own BatchNorm statistics come from fp32
    # shared-memory atomics: two forwards already differ by bf16 rounding flips that grow to ~1e-2 in the gradients)
    xt, mulv, kl = m.vae_forward(x, eps=eps)
    loss = VF.vae_loss(x, xt, kl)
    try:
        for fused in (False, True):
            VF.set_fuse_bn_backward(fused, thin=thin)
            for p in m.parameters():
                p.grad = None
            n0 = _lib.launch_count()
            loss.backward(retain_graph=True)
            torch.cuda.synchronize()
            launches.append(_lib.launch_count() - n0)
            res.append({k: npy(p.grad) for k, p in m.named_parameters() if p.grad is not None})
    finally:
        VF.set_fuse_bn_backward(True)
    assert launches[1] <= launches[0] - (2 if thin else 1), launches     # one reduce pass per fused block disappeared
    for k in res[0]:
        r = rel_l2(res[1][k], res[0][k])
        # the block whose reduction moved and everything downstream of it in backward order: same dx, only the summation
        # differs (measured 7e-5 on the block's own weight gradient).  Further upstream the difference passes through bf16
        # storage of each dx like any other rounding noise.
        first = 2 if thin else 1         # decoder.conv.<first>: the last block (in forward order) whose reduce pass is fused
        tight = any(k.startswith(f"decoder.conv.{i}.") for i in range(first, 4))
        assert r < (5e-4 if tight else 2e-2), f"{k}: rel-L2 {r:.3e} fused vs unfused BatchNorm backward"


def test_host_io_pipeline(vp):
    """HostBatchPipeline / ScalarReadback (vae_play_b200/host_io.py): every fed batch arrives intact and in order while the
    consumer keeps the device busy, staging buffers are not overwritten before their reader is done, and scalar read-backs
    return each step's value in order."""
    from vae_play_b200.host_io import HostBatchPipeline, ScalarReadback
    torch.manual_seed(0)
    batches = [torch.rand(64, 1, 64, 64).pin_memory() for _ in range(7)]
    pipe, reader = HostBatchPipeline((64, 1, 64, 64)), ScalarReadback(2)
    sink = torch.zeros(64, 1, 64, 64, device="cuda")
    busy = torch.rand(4096, 4096, device="cuda")
    got = []
    pipe.feed(batches[0])
    for i in range(len(batches)):
        xd = pipe.take()
        if i + 1 < len(batches):
            pipe.feed(batches[i + 1])
        for _ in range(3):
            busy = busy * 1.0001          # keep the consumer's stream behind the copy stream
        sink.copy_(xd, non_blocking=True)
        vals = torch.stack([sink.sum(), sink.flatten()[i]])
        pipe.release()
        if reader.pending() == 2:
            got.append(reader.pop())
        reader.push(vals)
    while reader.pending():
        got.append(reader.pop())
    assert len(got) == len(batches)
    for i, (b, g) in enumerate(zip(batches, got)):
        assert abs(g[0] - float(b.sum())) < 1e-3 * float(b.sum()) and g[1] == float(b.flatten()[i]), i
    with pytest.raises(ValueError):
        pipe.feed(torch.rand(64, 1, 64, 64))          # not pinned
    with pytest.raises(RuntimeError):
        pipe.feed(batches[0]); pipe.feed(batches[0]); pipe.feed(batches[0])


def test_async_wgrad_matches_sync(vp):
    """Weight gradients on the side stream (functional.set_async_wgrad) == on the main stream: ONE forward, two backward passes
    over the same graph into persistent slots, eagerly and inside a captured CUDA graph (fork / join inside the capture)."""
    import vae_play_b200.functional as VF
    from vae_play_b200.models.networks import VaeGan
    vp.set_precision("bf16")
    vp.set_engine("auto")
    VF.set_fuse_bn_backward(False)        # bit-comparable backward passes (see test_two_stage_backward_matches_single)
    torch.manual_seed(3)
    m = VaeGan(64, 128).cuda().train()
    x = torch.rand(64, 1, 64, 64, device="cuda")
    eps = torch.randn(64, 128, device="cuda")
    params = list(m.encoder.parameters()) + list(m.decoder.parameters())
    # everything on ONE non-default stream: autograd pins each AccumulateGrad node to the stream of the first forward that used
    # the parameter, and the legacy default stream cannot take part in a capture
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    try:
      with torch.cuda.stream(side):
          flat = VF.persistent_grads(params)
          xt, mulv, kl = m.vae_forward(x, eps=eps)
          loss = VF.vae_loss(x, xt, kl)
          res = []
          for on in (False, True, True):
              VF.set_async_wgrad(on)
              for p in params:
                  p.grad = None
              flat.zero_()
              for e in VF._GRAD_SINKS.values():
                  e[3], e[4] = True, False
              loss.backward(retain_graph=True)
              VF.join_async()
              torch.cuda.synchronize()
              res.append([npy(p.grad) for p in params])
          for a, b, c in zip(*res):
              assert rel_l2(b, a) < 1e-5 and rel_l2(c, a) < 1e-5
          # captured: the side stream forks from and joins the capture stream inside the graph
          for p in params:
              p.grad = None
          flat.zero_()
          for e in VF._GRAD_SINKS.values():
              e[3], e[4] = True, False
          g = torch.cuda.CUDAGraph()

          def fb():
              xt, mulv, kl = m.vae_forward(x, eps=eps)
              VF.vae_loss(x, xt, kl).backward()
              VF.join_async()
          for p in params:
              p.grad = None
          fb()                                      # warm-up outside the capture
          torch.cuda.synchronize()
          for p in params:
              p.grad = None
          flat.zero_()
          for e in VF._GRAD_SINKS.values():
              e[3], e[4] = True, False
          with torch.cuda.graph(g, stream=side):
              fb()
          flat.zero_()
          g.replay()
          torch.cuda.synchronize()
          for a, p in zip(res[0], params):
              assert rel_l2(npy(p.grad), a) < 0.1       # a second forward: its BatchNorm statistics differ by atomics order (bf16 flips, measured <= 5e-2 at batch 64); a missing or torn gradient would be O(1)
              assert np.isfinite(npy(p.grad)).all()
    finally:
        VF.set_async_wgrad(False)
        VF.set_fuse_bn_backward(True)
        VF.set_grad_sinks({})


def test_three_stage_backward_matches_single(vp):
    """The data-parallel step of vae_play_b200.engine cuts the backward twice (decoder | sample + heads + encoder.fc | encoder
    convs) so that each group's gradient exchange can start as soon as the group is complete: the three calls -- the middle one
    rooted at (z, kl) with d loss / d kl = 1 -- give the gradients of one backward call through the same forward graph."""
    import vae_play_b200.functional as VF
    from vae_play_b200.models.networks import VaeGan
    vp.set_precision("bf16")
    vp.set_engine("auto")
    VF.set_fuse_bn_backward(False)
    torch.manual_seed(2)
    m = VaeGan(64, 128).cuda().train()
    x = torch.rand(16, 1, 64, 64, device="cuda")
    eps = torch.randn(16, 128, device="cuda")
    params = list(m.encoder.parameters()) + list(m.decoder.parameters())
    conv_params = [p for blk in m.encoder.conv for p in blk.parameters()]
    dec_params = list(m.decoder.parameters())
    skip = {id(p) for p in conv_params} | {id(p) for p in dec_params}
    mid_params = [p for p in params if id(p) not in skip]
    try:
        VF.persistent_grads(params)
        taps, cut = [], {}
        xt, mulv, kl = m.vae_forward(x, eps=eps, taps=taps)
        loss = VF.vae_loss(x, xt, kl, mse_scale=0.5)
        a_out, a_in, z_out, z_in = taps
        z_out.register_hook(lambda g: cut.__setitem__("gz", g))
        loss.backward(inputs=dec_params + [z_out], retain_graph=True)
        z_out.grad = None
        assert all(p.grad is None for p in conv_params + mid_params) and all(p.grad is not None for p in dec_params)
        a_out.register_hook(lambda g: cut.__setitem__("g", g))
        torch.autograd.backward([z_in, kl], [cut["gz"], torch.ones_like(kl)], inputs=mid_params + [a_out], retain_graph=True)
        a_out.grad = None
        assert all(p.grad is None for p in conv_params) and all(p.grad is not None for p in mid_params)
        a_in.backward(cut["g"], inputs=conv_params, retain_graph=True)
        staged = [npy(p.grad) for p in params]
        VF.set_grad_sinks({})
        for p in params:
            p.grad = None
        loss.backward()
        names = [k for k, _ in list(m.encoder.named_parameters()) + list(m.decoder.named_parameters())]
        for k, a, p in zip(names, staged, params):
            assert rel_l2(a, npy(p.grad)) < 1e-4, k
    finally:
        VF.set_fuse_bn_backward(True)
        VF.set_grad_sinks({})
