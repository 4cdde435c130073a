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
from tests.util import TOL_BF16, TOL_FP32, digest, load, ref_dev, rel

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
    assert rel(npy(z), npy(want)) < 1e-6
    want_kl = -0.5 * torch.sum(-lv.exp() - mu ** 2 + lv + 1, 1)
    assert rel(npy(kl), npy(want_kl)) < 1e-5


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
    assert rel(npy(ypre), ops["eb/ypre"]) < t
    assert rel(npy(y), ops["eb/y"]) < t
    y.backward(cu(ops["eb/dy"]))
    assert rel(npy(x.grad), ops["eb/dx"]) < 3 * t
    assert rel(npy(m.conv.weight.grad), ops["eb/dw"]) < 3 * t
    assert rel(npy(m.bn.weight.grad), ops["eb/dg"]) < 3 * t
    assert rel(npy(m.bn.bias.grad), ops["eb/db"]) < 3 * t
    assert rel(npy(m.bn.running_mean), ops["eb/rm"]) < t
    assert rel(npy(m.bn.running_var), ops["eb/rv"]) < t
    assert int(m.bn.num_batches_tracked) == 1


def test_decoder_block_golden(vp, mode):
    from vae_play_b200.models.networks import DecoderBlock
    ops = load("ops.npz")
    m = DecoderBlock(10, 6).cuda().train()
    _load_block(m, ops, "db")
    x = cu(ops["db/x"], True)
    y = m(x)
    t = tol_for(mode)
    assert rel(npy(y), ops["db/y"]) < t
    y.backward(cu(ops["db/dy"]))
    assert rel(npy(x.grad), ops["db/dx"]) < 3 * t
    assert rel(npy(m.conv.weight.grad), ops["db/dw"]) < 3 * t
    assert rel(npy(m.bn.weight.grad), ops["db/dg"]) < 3 * t
    assert rel(npy(m.bn.bias.grad), ops["db/db"]) < 3 * t
    assert rel(npy(m.bn.running_mean), ops["db/rm"]) < t
    assert rel(npy(m.bn.running_var), ops["db/rv"]) < t


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
    assert rel(npy(y), g("y")) < t
    y.backward(cu(g("dy")))
    assert rel(npy(x.grad), g("dx")) < 3 * t
    assert rel(npy(m.conv[0].weight.grad), g("dw")) < 3 * t
    if bn is None:
        assert rel(npy(m.conv[0].bias.grad), g("dbias")) < 3 * t
    if bn == "batch":
        assert rel(npy(m.conv[1].weight.grad), g("dg")) < 3 * t
        assert rel(npy(m.conv[1].bias.grad), g("db")) < 3 * t


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
    assert rel(npy(y), ops["ct4/y"]) < t
    y.backward(cu(ops["ct4/dy"]))
    assert rel(npy(x.grad), ops["ct4/dx"]) < 3 * t
    assert rel(npy(w.grad), ops["ct4/dw"]) < 3 * t
    assert rel(npy(b.grad), ops["ct4/dbias"]) < 3 * t


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
    assert rel(npy(y), ops["lin/y"]) < t
    y.backward(cu(ops["lin/dy"]))
    assert rel(npy(x.grad), ops["lin/dx"]) < 3 * t
    assert rel(npy(m.fc[0].weight.grad), ops["lin/dw"]) < 3 * t
    assert rel(npy(m.fc[0].bias.grad), ops["lin/dbias"]) < 3 * t


def test_reparam_kl_golden(vp):
    ops = load("ops.npz")
    mu, lv = cu(ops["rp/mu"], True), cu(ops["rp/lv"], True)
    z, kl = vp.reparam_kl(mu, lv, eps=cu(ops["rp/eps"]))
    assert rel(npy(z), ops["rp/z"]) < 1e-6
    assert rel(npy(kl), ops["rp/kl"]) < 1e-6
    (kl.sum() + (z * cu(ops["rp/dz"])).sum()).backward()
    assert rel(npy(mu.grad), ops["rp/dmu"]) < 1e-6
    assert rel(npy(lv.grad), ops["rp/dlv"]) < 1e-6
    # packed (mu | logvar) form used on the hot path
    packed = torch.cat([cu(ops["rp/mu"]), cu(ops["rp/lv"])], dim=1).requires_grad_(True)
    z2, kl2 = vp.reparam_kl(packed, None, eps=cu(ops["rp/eps"]))
    (kl2.sum() + (z2 * cu(ops["rp/dz"])).sum()).backward()
    assert torch.equal(z2, z) and torch.equal(kl2, kl)
    assert rel(npy(packed.grad[:, :16]), ops["rp/dmu"]) < 1e-6
    assert rel(npy(packed.grad[:, 16:]), ops["rp/dlv"]) < 1e-6


def test_losses_golden(vp):
    ops = load("ops.npz")
    x = cu(ops["ls/x"])
    for nm, fn in (("mse", vp.mse_loss), ("l1", vp.l1_loss)):
        xt = cu(ops["ls/xt"], True)
        l = fn(x, xt)
        l.backward()
        assert rel(npy(l), ops[f"ls/{nm}"]) < 1e-6
        assert rel(npy(xt.grad), ops[f"ls/{nm}_dxt"]) < 1e-6
    # the scratch accumulator must be clean for a second call
    xt = cu(ops["ls/xt"], True)
    assert rel(npy(vp.mse_loss(x, xt)), ops["ls/mse"]) < 1e-6
    logits = cu(ops["ls/logits"], True)
    l = vp.bce_dice_loss(logits, cu(ops["ls/t"]), 0.5)
    l.backward()
    assert rel(npy(l), ops["ls/bce_dice"]) < 1e-6
    assert rel(npy(logits.grad), ops["ls/bce_dice_dlogits"]) < 1e-5


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
    import vae_play_b200.functional as VF
    x = cu(x_np)
    mulv = enc.forward_packed(x)
    z, kl = VF.reparam_kl(mulv, None, eps=cu(eps_np), z_dtype=VF.act_dtype())
    xt = VF.from_channels_last(dec.forward_cl(z.reshape(len(z), 1, 1, -1)))
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
    return out, grads, running


@pytest.mark.parametrize("case", ["vae64_c1_b4", "vae64_c3_b4", "vae128_c1_b4"])
def test_vae_step_golden(vp, mode, case):
    g = load(case + ".npz")
    img, cin, b, z, seed = [int(v) for v in g["meta"]]
    enc, dec, _ = build_vae(img, cin, z, seed)
    x, eps = vn.synth_batch(b, img, cin, z, seed)
    out, grads, running = run_step(vp, enc, dec, x, eps)
    dev = ref_dev(g)
    base = tol_for(mode)
    for key in ("mu", "logvar", "z", "x_tilde", "kl", "loss"):
        t = max(base, 3 * dev.get(key, 0.0))
        assert rel(out[key], g[key]) < t, (key, rel(out[key], g[key]))
    for key in g.files:
        if key.startswith("grad/"):
            t = max(3 * base, 3 * dev.get(key, 0.0))
            r = rel(digest(grads[key[5:]]), g[key])
            assert r < t, (key, r, t)
        if key.startswith("running/"):
            assert rel(running[key[8:]], g[key]) < max(base, 1e-5), key


@pytest.mark.parametrize("img,cin,b,seed", [(64, 1, 8, 5), (64, 3, 6, 6)])
def test_vae_step_vs_oracle(vp, mode, img, cin, b, seed):
    z = 128
    enc, dec, P = build_vae(img, cin, z, seed)
    x, eps = vn.synth_batch(b, img, cin, z, seed)
    want = vn.vae_step(P, x, eps)
    out, grads, running = run_step(vp, enc, dec, x, eps)
    base = tol_for(mode)
    # fp32: the reference's own fp32 run is only good to ~1e-5 on the BatchNorm-coupled gradients
    gt = 5e-5 if mode == "fp32" else 3 * base
    for key in ("mu", "logvar", "z", "x_tilde", "kl", "loss"):
        assert rel(out[key], want[key]) < base, (key, rel(out[key], want[key]))
    for key, gref in want["grads"].items():
        r = rel(grads[key], gref)
        assert r < gt, (key, r)
    for key, rref in want["running"].items():
        assert rel(running[key], rref) < max(base, 1e-5), key


# ------------------------------------------------------------------------------------------------
# size-independent properties at the benchmark size (B=256, 64x64): adjointness of fwd/dgrad/wgrad,
# normalisation invariants, loss identities
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,cin,cout,hw", [("conv", 64, 128, 32), ("convT", 256, 128, 16), ("conv", 1, 64, 64),
                                              ("flatten_in", 256, 1024, 8), ("flatten_out", 128, 256, 1)])
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
    elif kind == "flatten_in":
        layer = VF.TapLayer("flatten_in", cin, cout, spatial=8)
        w = torch.randn(cout, cin * 64, device="cuda") * 0.05
    else:
        layer = VF.TapLayer("flatten_out", cin, cout, spatial=8)
        w = torch.randn(cout * 64, cin, device="cuda") * 0.05
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
    assert rel(npy(bn.running_mean), 0.1 * xf.mean(0).cpu().numpy()) < (1e-2 if mode == "bf16" else 1e-5)


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
