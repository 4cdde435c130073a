"""tcgen05 kernels against float64 ON THE SAME bf16-QUANTISED OPERANDS, one layer at a time, at the benchmark's layer shapes
(VAE 64x64: encoder.conv.1/2, decoder.conv.0/1/2, encoder.fc, decoder.fc, the thin first / last layers), plus the
reference-pinned per-operator fixtures with the activation pattern pinned (max-norm bounds instead of rel-L2).

north_star tolerance: rel <= 1e-2 in bf16, measured as max|a-b| / max|b| per tensor.  With identical (already quantised)
operands the only errors left are the fp32 accumulation order and ONE rounding of the stored output:
    y, dx stored as bf16  <= 1e-2 (asserted; expected ~2^-9 = 2e-3 of the largest element)
    y, dx read back fp32  <= 1e-4
    dw (always fp32)      <= 1e-4
Run on the B200 box with ``pytest -m gpu``.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import vae_numpy as vn
from tests.util import load, rel

pytestmark = pytest.mark.gpu

TOL_BF16_STORE = 1e-2
TOL_FP32_OUT = 1e-4


@pytest.fixture(scope="module")
def vp():
    import vae_play_b200
    return vae_play_b200


def npy(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32))).cuda()


# name, kind, cin, cout, input hw, batch (the bench runs batch 256; 16 keeps the float64 check in seconds and still spans
# several 128-row tiles per SM slot for every layer; the fc layers run at the full batch)
BENCH_LAYERS = [
    ("encoder.conv.1", "conv", 64, 128, 32, 16),
    ("encoder.conv.2", "conv", 128, 256, 16, 16),
    ("decoder.conv.0", "convT", 256, 256, 8, 16),
    ("decoder.conv.1", "convT", 256, 128, 16, 16),
    ("decoder.conv.2", "convT", 128, 64, 32, 16),
    ("encoder.fc", "linear", 16384, 1024, 1, 256),
    ("decoder.fc", "linear", 128, 16384, 1, 256),
    ("encoder.l_mu|l_var", "linear", 1024, 256, 1, 256),
    ("encoder.conv.0 (thin)", "conv", 1, 64, 64, 16),
    ("decoder.conv.3 (thin)", "conv_s1", 64, 1, 64, 16),
    # 128x128 model: the 512-channel layers
    ("encoder.conv.3 @128", "conv", 256, 512, 16, 8),
    ("decoder.conv.0 @128", "convT", 512, 512, 8, 8),
]


def _make(kind, cin, cout):
    import vae_play_b200.functional as VF
    g = torch.Generator(device="cuda").manual_seed(17)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g) * 0.05
    if kind == "conv":
        layer, w = VF.TapLayer("conv", cin, cout, k=5, stride=2, pad=2), r(cout, cin, 5, 5)
    elif kind == "conv_s1":
        layer, w = VF.TapLayer("conv", cin, cout, k=5, stride=1, pad=2), r(cout, cin, 5, 5)
    elif kind == "convT":
        layer, w = VF.TapLayer("convT", cin, cout, k=5, stride=2, pad=2, out_pad=1), r(cin, cout, 5, 5)
    else:
        layer, w = VF.TapLayer("linear", cin, cout), r(cout, cin)
    if w.dim() == 4 and cin % 64 == 0 and cout % 64 == 0:
        w = w.contiguous(memory_format=torch.channels_last)          # the layout the model keeps these weights in
    return layer, w


def _reference(kind, xq, wq, dyq):
    """float64 torch on the quantised operands (NCHW views): y, dx, dw."""
    xd = xq.double().permute(0, 3, 1, 2)
    wd = wq.double()
    if kind in ("conv", "conv_s1"):
        s = 2 if kind == "conv" else 1
        y = F.conv2d(xd, wd, None, stride=s, padding=2)
        if dyq is None:
            return y.permute(0, 2, 3, 1), None, None
        dyd = dyq.double().permute(0, 3, 1, 2)
        dx = torch.nn.grad.conv2d_input(xd.shape, wd, dyd, stride=s, padding=2)
        dw = torch.nn.grad.conv2d_weight(xd, wd.shape, dyd, stride=s, padding=2)
    elif kind == "convT":
        y = F.conv_transpose2d(xd, wd, None, stride=2, padding=2, output_padding=1)
        if dyq is None:
            return y.permute(0, 2, 3, 1), None, None
        dyd = dyq.double().permute(0, 3, 1, 2)
        dx = F.conv2d(dyd, wd, None, stride=2, padding=2)                         # adjoint of the transposed conv
        dw = torch.nn.grad.conv2d_weight(dyd, wd.shape, xd, stride=2, padding=2)   # conv2d(dy, w) has weight w: dw = wgrad(dy, x)
    else:
        x2, w2 = xd.reshape(len(xd), -1), wd
        y = (x2 @ w2.T).reshape(len(xd), 1, 1, -1).permute(0, 3, 1, 2)
        if dyq is None:
            return y.permute(0, 2, 3, 1), None, None
        d2 = dyq.double().reshape(len(xd), -1)
        dx = (d2 @ w2).reshape(xd.shape)
        dw = d2.T @ x2
    return y.permute(0, 2, 3, 1), dx.permute(0, 2, 3, 1), dw


@pytest.mark.parametrize("name,kind,cin,cout,hw,b", BENCH_LAYERS)
def test_tc_layer_vs_oracle(vp, name, kind, cin, cout, hw, b):
    import vae_play_b200.functional as VF
    vp.set_precision("bf16")
    vp.set_engine("auto")
    VF.set_grad_sinks({})
    layer, w = _make(kind, cin, cout)
    g = torch.Generator(device="cuda").manual_seed(23)
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g).to(torch.bfloat16)
    wq = w.to(torch.bfloat16)                      # what the kernels multiply with (bf16 operand copy / in-kernel rounding)
    n0 = vp._lib.launch_count()
    y16 = layer.fwd(x, w, None)
    assert vp._lib.launch_count() - n0 <= 4        # the contraction (+ at most: cast of the weight, split-K workspace clear + finish), no CUDA-core fallback
    y32 = layer.fwd(x, w, None, out_dtype=torch.float32)
    dy = torch.randn(y16.shape, device="cuda", generator=g).to(torch.bfloat16)
    want_y, want_dx, want_dw = _reference(kind, x, wq, dy)
    r16, r32 = rel(npy(y16), npy(want_y)), rel(npy(y32), npy(want_y))
    assert r16 < TOL_BF16_STORE, f"{name} fwd (bf16 store): rel {r16:.3e}"
    assert r32 < TOL_FP32_OUT, f"{name} fwd (fp32 out): rel {r32:.3e}"
    if cin > 1:
        dx16 = layer.dgrad(dy, w, tuple(x.shape))
        dx32 = layer.dgrad(dy, w, tuple(x.shape), out_dtype=torch.float32)
        r16, r32 = rel(npy(dx16), npy(want_dx)), rel(npy(dx32), npy(want_dx))
        assert r16 < TOL_BF16_STORE, f"{name} dgrad (bf16 store): rel {r16:.3e}"
        assert r32 < TOL_FP32_OUT, f"{name} dgrad (fp32 out): rel {r32:.3e}"
    dw = layer.wgrad(x, dy, w)
    rw = rel(npy(dw), npy(want_dw))
    assert rw < TOL_FP32_OUT, f"{name} wgrad: rel {rw:.3e}"
    # the engine that ran is the tensor-core one: the CUDA-core engine would need packed panels (vp_pack_weight launches)
    assert layer._thin("fwd", torch.bfloat16, w) or layer._cl(torch.bfloat16, w)


@pytest.mark.parametrize("name,kind,cin,cout,hw,b", [("encoder.conv.1", "conv", 64, 128, 32, 2), ("decoder.conv.2", "convT", 128, 64, 32, 2),
                                                     ("decoder.fc", "linear", 128, 16384, 1, 8)])
def test_tc_layer_vs_numpy_oracle(vp, name, kind, cin, cout, hw, b):
    """Same check against oracle/vae_numpy.py itself (the restatement pinned to the reference), at a batch it finishes in seconds."""
    import vae_play_b200.functional as VF
    vp.set_precision("bf16")
    vp.set_engine("auto")
    VF.set_grad_sinks({})
    layer, w = _make(kind, cin, cout)
    g = torch.Generator(device="cuda").manual_seed(29)
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g).to(torch.bfloat16)
    wq = npy(w.to(torch.bfloat16))
    xn = npy(x).transpose(0, 3, 1, 2)
    y32 = layer.fwd(x, w, None, out_dtype=torch.float32)
    dy = torch.randn(y32.shape, device="cuda", generator=g).to(torch.bfloat16)
    dyn = npy(dy).transpose(0, 3, 1, 2)
    if kind == "conv":
        want_y = vn.conv2d_fwd(xn, wq, None, 2, 2)
        want_dx, want_dw = vn.conv2d_bwd(xn, wq, dyn, 2, 2)[:2]
    elif kind == "convT":
        want_y = vn.conv_transpose2d_fwd(xn, wq, None, 2, 2, 1)
        want_dx, want_dw = vn.conv_transpose2d_bwd(xn, wq, dyn, 2, 2, 1)[:2]
    else:
        want_y = vn.linear_fwd(xn.reshape(b, -1), wq).reshape(b, -1, 1, 1)
        dx2, want_dw, _ = vn.linear_bwd(xn.reshape(b, -1), wq, dyn.reshape(b, -1))
        want_dx = dx2.reshape(xn.shape)
    assert rel(npy(y32).transpose(0, 3, 1, 2), want_y) < TOL_FP32_OUT
    dx32 = layer.dgrad(dy, w, tuple(x.shape), out_dtype=torch.float32)
    assert rel(npy(dx32).transpose(0, 3, 1, 2), want_dx) < TOL_FP32_OUT
    dw = layer.wgrad(x, dy, w)
    assert rel(npy(dw), want_dw) < TOL_FP32_OUT


# ---------------------------------------------------------------------------------------------------------------------
# reference-pinned per-operator fixtures (tests/golden/ops.npz), bf16 mode, MAX-NORM bounds with the activation pattern pinned
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_block_bwd(kind, x, w, gamma, beta, dy, mask, stride, pad, norm, act, slope, bias=None):
    """Backward of conv -> [BN | IN | -] -> act through oracle/vae_numpy.py with the ReLU / LeakyReLU pattern taken from the
    implementation under test.  Returns (dx, dw, dgamma, dbeta, dbias, pre-activation)."""
    if kind == "conv":
        y = vn.conv2d_fwd(x, w, bias, stride, pad)
    else:
        y = vn.conv_transpose2d_fwd(x, w, bias, stride, pad, 1)
    if norm == "batch":
        pre, cache, _, _ = vn.batchnorm_train_fwd(y, gamma, beta)
    elif norm == "instance":
        pre, cache = vn.instancenorm_fwd(y)
    else:
        pre = y
    out = vn.act_fwd(pre, act, slope)
    d = vn.act_bwd(dy, pre, out, act, slope, mask=mask)
    dg = db = None
    if norm == "batch":
        d, dg, db = vn.batchnorm_train_bwd(d, gamma, cache)
    elif norm == "instance":
        d = vn.instancenorm_bwd(d, cache)
    dbias = d.sum(axis=(0, 2, 3)) if bias is not None else None
    if kind == "conv":
        dx, dw = vn.conv2d_bwd(x, w, d, stride, pad)[:2]
    else:
        dx, dw = vn.conv_transpose2d_bwd(x, w, d, stride, pad, 1)[:2]
    return dx, dw, dg, db, dbias, pre


def _check_pattern(pre, mask, noise=6e-2, max_frac=2e-2):
    """the implementation's activation pattern may differ from the oracle's only where the pre-activation is within bf16 noise of 0"""
    bad = mask != (pre > 0)
    assert bad.mean() <= max_frac, f"{bad.mean():.2e} of the activation pattern differs"
    if bad.any():
        assert np.abs(pre[bad]).max() / np.abs(pre).max() <= noise


GRAD_TOL_BF16 = 3e-2      # one layer, fp32-accurate fixtures vs bf16-quantised x, w, dy AND bf16-stored y: <= 3 x the 1e-2 forward bound


@pytest.mark.parametrize("blk", ["eb", "db"])
def test_block_gradients_pinned_pattern_bf16(vp, blk):
    """EncoderBlock / DecoderBlock (reference networks.py:10-46) in bf16 mode: gradients against the oracle's backward with the
    CUDA ReLU pattern pinned, element-wise max-norm (replaces the round-1 rel-L2 < 0.15 bound)."""
    from vae_play_b200.models.networks import DecoderBlock, EncoderBlock
    vp.set_precision("bf16")
    vp.set_engine("auto")
    ops = load("ops.npz")
    m = (EncoderBlock(6, 10) if blk == "eb" else DecoderBlock(10, 6)).cuda().train()
    with torch.no_grad():
        m.conv.weight.copy_(cu(ops[f"{blk}/w"]))
        m.bn.weight.copy_(cu(ops[f"{blk}/g"]))
        m.bn.bias.copy_(cu(ops[f"{blk}/b"]))
    x = cu(ops[f"{blk}/x"]).requires_grad_(True)
    y = m(x)
    y.backward(cu(ops[f"{blk}/dy"]))
    mask = npy(y) > 0
    dx, dw, dg, db, _, pre = _oracle_block_bwd("conv" if blk == "eb" else "convT", ops[f"{blk}/x"], ops[f"{blk}/w"], ops[f"{blk}/g"],
                                               ops[f"{blk}/b"], ops[f"{blk}/dy"], mask, 2, 2, "batch", "relu", 0.0)
    _check_pattern(pre, mask)
    for got, want, nm in ((x.grad, dx, "dx"), (m.conv.weight.grad, dw, "dw"), (m.bn.weight.grad, dg, "dgamma"), (m.bn.bias.grad, db, "dbeta")):
        r = rel(npy(got), want)
        assert r < GRAD_TOL_BF16, f"{blk} {nm}: rel {r:.3e}"


@pytest.mark.parametrize("name,ci,co,k,s,bn,act", [
    ("c_k3s1_batch_relu", 5, 7, 3, 1, "batch", "relu"), ("c_k4s2_inst_lrelu", 4, 6, 4, 2, "instance", "lrelu"),
    ("c_k1s1_none_tanh", 6, 3, 1, 1, None, "tanh"), ("c_k5s1_none_none", 3, 2, 5, 1, None, None),
    ("c_k3s2_batch_lrelu", 4, 8, 3, 2, "batch", "lrelu")])
def test_blocks_conv2d_gradients_pinned_pattern_bf16(vp, name, ci, co, k, s, bn, act):
    """blocks.Conv2d variants (reference blocks.py:5-34) in bf16 mode, max-norm with the activation pattern pinned."""
    from vae_play_b200.models.blocks import Conv2d
    vp.set_precision("bf16")
    vp.set_engine("auto")
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
    x = cu(g("x")).requires_grad_(True)
    y = m(x)
    y.backward(cu(g("dy")))
    mask = (npy(y) > 0) if act in ("relu", "lrelu") else None
    dx, dw, dg, db, dbias, pre = _oracle_block_bwd("conv", g("x"), g("w"), g("g") if bn == "batch" else None, g("b") if bn == "batch" else None,
                                                   g("dy"), mask, s, (k - 1) // 2, bn, act, 0.02, bias=g("bias") if bn is None else None)
    if mask is not None:
        _check_pattern(pre, mask)
    checks = [(x.grad, dx, "dx"), (m.conv[0].weight.grad, dw, "dw")]
    if bn is None:
        checks.append((m.conv[0].bias.grad, dbias, "dbias"))
    if bn == "batch":
        checks += [(m.conv[1].weight.grad, dg, "dgamma"), (m.conv[1].bias.grad, db, "dbeta")]
    for got, want, nm in checks:
        r = rel(npy(got), want)
        assert r < GRAD_TOL_BF16, f"{name} {nm}: rel {r:.3e}"


# ---------------------------------------------------------------------------------------------------------------------
# channel counts that are not multiples of 64: the padded tcgen05 route (3, 4, 6, 16, 32 ... channels)
# ---------------------------------------------------------------------------------------------------------------------
PADDED_SHAPES = [
    # kind, cin, cout, k, stride, pad, out_pad, hw, batch
    ("conv", 3, 64, 5, 2, 2, 0, 32, 4),        # RGB first layer (BASELINE config 1)
    ("conv", 32, 64, 5, 2, 2, 0, 32, 4),       # VAE-GAN discriminator conv.1 (networks.py:162)
    ("conv", 32, 32, 3, 1, 1, 0, 24, 3),       # Style_GAN full-resolution 3x3 convs (network_Style_GAN.py:96,119-120)
    ("conv", 4, 32, 3, 1, 1, 0, 20, 2),        # Generator.conv1 (:95)
    ("conv", 6, 64, 5, 1, 2, 0, 16, 2),        # Style discriminator first layer (:205)
    ("conv", 16, 48, 3, 2, 1, 0, 18, 3),
    ("conv", 64, 96, 4, 2, 1, 0, 16, 2),       # k4 s2 p1 down-sampling (:98-101), ragged output channels
    ("convT", 64, 32, 4, 2, 1, 0, 8, 3),       # Generator.final[0] (:116)
    ("convT", 10, 6, 5, 2, 2, 1, 7, 2),
    ("linear", 96, 40, 1, 1, 0, 0, 1, 48),
    ("linear", 32, 2, 1, 1, 0, 0, 1, 64),      # DirectDecoder.xy_fc[1] (networks.py:136)
]


@pytest.mark.parametrize("kind,cin,cout,k,stride,pad,out_pad,hw,b", PADDED_SHAPES)
def test_padded_route_vs_float64(vp, kind, cin, cout, k, stride, pad, out_pad, hw, b):
    """fwd / dgrad / wgrad of layers with arbitrary channel counts in bf16 mode: zero-padded onto the tcgen05 kernels, compared
    with float64 on the same bf16-quantised operands; no CUDA-core contraction is launched."""
    import vae_play_b200.functional as VF
    from vae_play_b200 import _lib
    vp.set_precision("bf16")
    vp.set_engine("auto")
    VF.set_grad_sinks({})
    g = torch.Generator(device="cuda").manual_seed(31)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g) * 0.1
    if kind == "conv":
        layer, w = VF.TapLayer("conv", cin, cout, k=k, stride=stride, pad=pad), r(cout, cin, k, k)
    elif kind == "convT":
        layer, w = VF.TapLayer("convT", cin, cout, k=k, stride=stride, pad=pad, out_pad=out_pad), r(cin, cout, k, k)
    else:
        layer, w = VF.TapLayer("linear", cin, cout), r(cout, cin)
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(cout, device="cuda", generator=g)
    simt0 = _lib.simt_bf16_count()
    y32 = layer.fwd(x, w, bias, out_dtype=torch.float32)
    y16 = layer.fwd(x, w, bias, "relu")
    dy = torch.randn(y32.shape, device="cuda", generator=g).to(torch.bfloat16)
    dx32 = layer.dgrad(dy, w, tuple(x.shape), out_dtype=torch.float32)
    dw = layer.wgrad(x, dy, w)
    assert _lib.simt_bf16_count() == simt0, "a bf16 contraction fell back to the CUDA-core engine"
    xd, wd, dyd = x.double().permute(0, 3, 1, 2), w.to(torch.bfloat16).double(), dy.double().permute(0, 3, 1, 2)
    if kind == "conv":
        want_y = F.conv2d(xd, wd, bias.double(), stride=stride, padding=pad)
        want_dx = torch.nn.grad.conv2d_input(xd.shape, wd, dyd, stride=stride, padding=pad)
        want_dw = torch.nn.grad.conv2d_weight(xd, wd.shape, dyd, stride=stride, padding=pad)
    elif kind == "convT":
        want_y = F.conv_transpose2d(xd, wd, bias.double(), stride=stride, padding=pad, output_padding=out_pad)
        want_dx = F.conv2d(dyd, wd, None, stride=stride, padding=pad)
        want_dw = torch.nn.grad.conv2d_weight(dyd, wd.shape, xd, stride=stride, padding=pad)
    else:
        x2, d2 = xd.reshape(b, cin), dyd.reshape(b, cout)
        want_y = (x2 @ wd.T + bias.double()).reshape(b, cout, 1, 1)
        want_dx = (d2 @ wd).reshape(b, cin, 1, 1)
        want_dw = d2.T @ x2
    assert rel(npy(y32), npy(want_y.permute(0, 2, 3, 1))) < TOL_FP32_OUT
    assert rel(npy(y16), npy(torch.relu(want_y).permute(0, 2, 3, 1))) < TOL_BF16_STORE
    assert rel(npy(dx32), npy(want_dx.permute(0, 2, 3, 1))) < TOL_FP32_OUT
    assert rel(npy(dw), npy(want_dw)) < TOL_FP32_OUT


def test_bf16_discriminator_step_runs_on_tensor_cores(vp):
    """VAE-GAN Discriminator (networks.py:151-195: 1 -> 32 -> 64 -> 128 -> 256 channels) forward + backward in REC and GAN mode
    in bf16: no contraction on the CUDA-core engine (round 1 ran the 1 -> 32 and 32 -> 64 layers there)."""
    import vae_play_b200.functional as VF
    from vae_play_b200 import _lib
    from vae_play_b200.models.networks import Discriminator
    vp.set_precision("bf16")
    vp.set_engine("auto")
    VF.set_grad_sinks({})
    torch.manual_seed(0)
    d = Discriminator(channel_in=1, recon_level=3, iter_level=3).cuda().train()
    xs = [torch.rand(4, 1, 64, 64, device="cuda").requires_grad_(i > 0) for i in range(3)]
    simt0 = _lib.simt_bf16_count()
    rec = d(*xs, "REC")
    gan = d(*xs, "GAN")
    (rec.float().square().sum() * 1e-3 + gan.float().sum()).backward()
    torch.cuda.synchronize()
    assert _lib.simt_bf16_count() == simt0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in d.parameters())


@pytest.mark.parametrize("cin,cout,k,hw,b,act", [(64, 4, 1, 1, 3, "relu"), (64, 4, 1, 9, 2, "none"), (64, 2, 3, 12, 2, "sigmoid"),
                                                 (128, 3, 3, 17, 2, "tanh"), (64, 32, 1, 8, 2, "none"), (64, 1, 5, 20, 2, "sigmoid"),
                                                 (256, 5, 1, 6, 3, "relu")])
def test_thin_output_layers_multi_channel(vp, cin, cout, k, hw, b, act):
    """The thin-output forward kernel (a few output channels: the decoder's 64 -> 1 layer, SCSE's 64 -> 4 squeeze, BE heads) with
    MORE than one output channel, a bias per channel, every activation, ragged bricks: float64 reference on the bf16-quantised
    operands.  (A two-channel shortcut for the bias once slipped through the one-channel bench-shape test.)"""
    import vae_play_b200.functional as VF
    vp.set_precision("bf16")
    vp.set_engine("auto")
    layer = VF.TapLayer("conv", cin, cout, k=k, stride=1, pad=k // 2)
    g = torch.Generator(device="cuda").manual_seed(5)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) * 0.1
    bias = torch.randn(cout, device="cuda", generator=g)
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g).to(torch.bfloat16)
    assert layer._thin("fwd", torch.bfloat16, w)
    y = layer.fwd(x, w, bias, act=act, out_dtype=torch.float32)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.to(torch.bfloat16).double(), bias.double(), padding=k // 2)
    ref = {"none": lambda t: t, "relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}[act](ref).permute(0, 2, 3, 1)
    r = rel(npy(y), npy(ref))
    assert r < TOL_FP32_OUT, f"thin forward {cin}->{cout} k{k} {act}: rel {r:.3e}"


@pytest.mark.parametrize("cout,k,hw,b", [(32, 5, 20, 3), (64, 5, 33, 2), (32, 3, 16, 2), (128, 5, 12, 2)])
def test_single_channel_input_dgrad(vp, cout, k, hw, b):
    """Data gradient w.r.t. a ONE-channel input (the VAE-GAN discriminator's first layer sits on x_tilde, which needs a gradient:
    models/networks.py:159): served by the thin-output kernel with flipped taps, 32-channel dy zero-padded to 64.  float64
    reference on the bf16-quantised operands; no CUDA-core contraction, no 64x-padded tensor-core launch."""
    import vae_play_b200.functional as VF
    vp.set_precision("bf16")
    vp.set_engine("auto")
    layer = VF.TapLayer("conv", 1, cout, k=k, stride=1, pad=k // 2)
    g = torch.Generator(device="cuda").manual_seed(11)
    w = torch.randn(cout, 1, k, k, device="cuda", generator=g) * 0.1
    dy = torch.randn(b, hw, hw, cout, device="cuda", generator=g).to(torch.bfloat16)
    simt0 = vp._lib.simt_bf16_count()
    dx = layer.dgrad(dy, w, (b, hw, hw, 1), out_dtype=torch.float32)
    assert vp._lib.simt_bf16_count() == simt0
    ref = torch.nn.grad.conv2d_input((b, 1, hw, hw), w.to(torch.bfloat16).double(), dy.double().permute(0, 3, 1, 2), stride=1, padding=k // 2)
    r = rel(npy(dx), npy(ref.permute(0, 2, 3, 1)))
    assert r < TOL_FP32_OUT, f"dgrad to a single-channel input, {cout} channels k{k}: rel {r:.3e}"
