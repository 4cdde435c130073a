"""Pin the CPU oracle (oracle/vae_numpy.py, oracle/philox.py) against the reference-generated fixtures.

The fixtures in tests/golden were produced by oracle/gen_golden.py from the
unmodified reference modules in float64; the oracle must reproduce them to
float64 round-off.  CPU only.
"""
import os

import numpy as np
import pytest

from oracle import philox
from oracle import vae_numpy as vn

TOL = 1e-9


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-300)


def digest(a, nsamp=64):
    a = np.asarray(a, np.float64).ravel()
    idx = np.linspace(0, a.size - 1, num=min(nsamp, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum()), np.abs(a).max()], a[idx]])


@pytest.fixture(scope="module")
def ops(golden_dir):
    return np.load(os.path.join(golden_dir, "ops.npz"))


def test_philox_known_answers():
    for ctr, key, want in philox.KAT:
        got = philox.philox4x32_10(np.array([ctr], np.uint32), np.array([key], np.uint32))[0]
        assert tuple(int(v) for v in got) == want


def test_philox_policy():
    # DistributionTemplates.h:50-62: grid = min(ceil(n/256), 148*8); offset advance in units of 4
    assert philox.aten_normal_policy(256 * 128) == (128, 4)
    assert philox.aten_normal_policy(1) == (1, 4)
    assert philox.aten_normal_policy(148 * 8 * 256 * 4 + 1) == (1184, 8)
    x = philox.aten_normal(1 << 16, 42, 0)
    assert abs(float(x.mean())) < 0.02 and abs(float(x.std()) - 1) < 0.02


def test_encoder_block(ops):
    y = vn.conv2d_fwd(ops["eb/x"], ops["eb/w"], None, 2, 2)
    assert rel(y, ops["eb/ypre"]) < TOL
    yb, cache, bm, bv = vn.batchnorm_train_fwd(y, ops["eb/g"], ops["eb/b"])
    a = vn.act_fwd(yb, "relu")
    assert rel(a, ops["eb/y"]) < TOL
    assert rel(vn.bn_running(np.zeros(10), bm, 0.9), ops["eb/rm"]) < TOL
    assert rel(vn.bn_running(np.ones(10), bv, 0.9), ops["eb/rv"]) < TOL
    dyb = vn.act_bwd(ops["eb/dy"], yb, a, "relu")
    dy, dg, db = vn.batchnorm_train_bwd(dyb, ops["eb/g"], cache)
    dx, dw, _ = vn.conv2d_bwd(ops["eb/x"], ops["eb/w"], dy, 2, 2)
    for got, key in ((dx, "dx"), (dw, "dw"), (dg, "dg"), (db, "db")):
        assert rel(got, ops["eb/" + key]) < TOL, key


def test_decoder_block(ops):
    y = vn.conv_transpose2d_fwd(ops["db/x"], ops["db/w"], None, 2, 2, 1)
    yb, cache, bm, bv = vn.batchnorm_train_fwd(y, ops["db/g"], ops["db/b"])
    a = vn.act_fwd(yb, "relu")
    assert rel(a, ops["db/y"]) < TOL
    assert rel(vn.bn_running(np.zeros(6), bm, 0.9), ops["db/rm"]) < TOL
    assert rel(vn.bn_running(np.ones(6), bv, 0.9), ops["db/rv"]) < TOL
    dyb = vn.act_bwd(ops["db/dy"], yb, a, "relu")
    dy, dg, db = vn.batchnorm_train_bwd(dyb, ops["db/g"], cache)
    dx, dw, _ = vn.conv_transpose2d_bwd(ops["db/x"], ops["db/w"], dy, 2, 2, 1)
    for got, key in ((dx, "dx"), (dw, "dw"), (dg, "dg"), (db, "db")):
        assert rel(got, ops["db/" + key]) < TOL, key


@pytest.mark.parametrize("name,k,s,bn,act", [
    ("c_k3s1_batch_relu", 3, 1, "batch", "relu"), ("c_k4s2_inst_lrelu", 4, 2, "instance", "lrelu"),
    ("c_k1s1_none_tanh", 1, 1, None, "tanh"), ("c_k5s1_none_none", 5, 1, None, None),
    ("c_k3s2_batch_lrelu", 3, 2, "batch", "lrelu")])
def test_blocks_conv2d(ops, name, k, s, bn, act):
    g = lambda key: ops[f"{name}/{key}"]
    bias = g("bias") if bn is None else None
    y = vn.conv2d_fwd(g("x"), g("w"), bias, s, (k - 1) // 2)
    if bn == "batch":
        yn, cache, _, _ = vn.batchnorm_train_fwd(y, g("g"), g("b"))
    elif bn == "instance":
        yn, cache = vn.instancenorm_fwd(y)
    else:
        yn = y
    a = vn.act_fwd(yn, act, 0.02)
    assert rel(a, g("y")) < TOL
    d = vn.act_bwd(g("dy"), yn, a, act, 0.02)
    if bn == "batch":
        d, dg, db = vn.batchnorm_train_bwd(d, g("g"), cache)
        assert rel(dg, g("dg")) < TOL and rel(db, g("db")) < TOL
    elif bn == "instance":
        d = vn.instancenorm_bwd(d, cache)
    dx, dw, dbias = vn.conv2d_bwd(g("x"), g("w"), d, s, (k - 1) // 2)
    assert rel(dx, g("dx")) < TOL and rel(dw, g("dw")) < TOL
    if bn is None:
        assert rel(dbias, g("dbias")) < TOL


def test_conv_transpose_k4(ops):
    y = vn.conv_transpose2d_fwd(ops["ct4/x"], ops["ct4/w"], ops["ct4/bias"], 2, 1, 0)
    assert rel(y, ops["ct4/y"]) < TOL
    dx, dw, db = vn.conv_transpose2d_bwd(ops["ct4/x"], ops["ct4/w"], ops["ct4/dy"], 2, 1, 0)
    assert rel(dx, ops["ct4/dx"]) < TOL and rel(dw, ops["ct4/dw"]) < TOL and rel(db, ops["ct4/dbias"]) < TOL


def test_blocks_linear(ops):
    y0 = vn.linear_fwd(ops["lin/x"], ops["lin/w"], ops["lin/bias"])
    y = vn.act_fwd(y0, "lrelu", 0.2)
    assert rel(y, ops["lin/y"]) < TOL
    d = vn.act_bwd(ops["lin/dy"], y0, y, "lrelu", 0.2)
    dx, dw, db = vn.linear_bwd(ops["lin/x"], ops["lin/w"], d)
    assert rel(dx, ops["lin/dx"]) < TOL and rel(dw, ops["lin/dw"]) < TOL and rel(db, ops["lin/dbias"]) < TOL


def test_reparam_kl(ops):
    z = vn.reparameterize(ops["rp/mu"], ops["rp/lv"], ops["rp/eps"])
    assert rel(z, ops["rp/z"]) < TOL
    assert rel(vn.kl_per_sample(ops["rp/mu"], ops["rp/lv"]), ops["rp/kl"]) < TOL
    assert rel(vn.nle(ops["rp/x"], ops["rp/xt"]), ops["rp/nle"]) < TOL
    dmu_r, dlv_r = vn.reparameterize_bwd(ops["rp/dz"], ops["rp/lv"], ops["rp/eps"])
    dmu_k, dlv_k = vn.kl_bwd(ops["rp/mu"], ops["rp/lv"], np.ones(6))
    assert rel(dmu_r + dmu_k, ops["rp/dmu"]) < TOL and rel(dlv_r + dlv_k, ops["rp/dlv"]) < TOL


def test_losses(ops):
    l, d = vn.mse_mean(ops["ls/x"], ops["ls/xt"])
    assert rel(l, ops["ls/mse"]) < TOL and rel(d, ops["ls/mse_dxt"]) < TOL
    l, d = vn.l1_mean(ops["ls/x"], ops["ls/xt"])
    assert rel(l, ops["ls/l1"]) < TOL and rel(d, ops["ls/l1_dxt"]) < TOL
    lb, db = vn.bce_with_logits_mean(ops["ls/logits"], ops["ls/t"])
    p = vn.act_fwd(ops["ls/logits"], "sigmoid")
    ld, dp = vn.dice_loss(p, ops["ls/t"])
    assert rel(0.5 * lb + ld, ops["ls/bce_dice"]) < TOL
    assert rel(0.5 * db + dp * p * (1 - p), ops["ls/bce_dice_dlogits"]) < TOL


@pytest.mark.parametrize("case", ["vae64_c1_b4", "vae64_c3_b4", "vae128_c1_b4"])
def test_vae_step(golden_dir, case):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    img, cin, b, z, seed = [int(v) for v in g["meta"]]
    P = vn.synth_vae_params(img, z, cin, cin, seed)
    x, eps = vn.synth_batch(b, img, cin, z, seed)
    out = vn.vae_step(P, x, eps)
    for key in ("mu", "logvar", "z", "x_tilde", "kl", "mse", "loss"):
        assert rel(out[key], g[key]) < 1e-8, key
    assert rel(vn.nle(x, out["x_tilde"]).sum(), g["nle_sum"]) < 1e-8
    for key in g.files:
        if key.startswith("grad/"):
            assert rel(digest(out["grads"][key[5:]]), g[key]) < 1e-7, key
        if key.startswith("running/"):
            assert rel(out["running"][key[8:]], g[key]) < 1e-8, key


@pytest.mark.parametrize("case", ["vae64_c1_b4", "vae64_c3_b4"])
def test_torch_port_matches_golden(golden_dir, case):
    import torch
    from oracle.vae_torch import VaeTorchPort
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    img, cin, b, z, seed = [int(v) for v in g["meta"]]
    P = vn.synth_vae_params(img, z, cin, cin, seed)
    x, eps = vn.synth_batch(b, img, cin, z, seed)
    port = VaeTorchPort(P, torch.float64)
    loss, out = port.step(torch.from_numpy(x).double(), torch.from_numpy(eps).double(), optimize=False)
    assert rel(loss.numpy(), g["loss"]) < 1e-10
    for key in ("mu", "logvar", "z", "x_tilde", "kl"):
        assert rel(out[key].detach().numpy(), g[key]) < 1e-9, key
    for key, gr in port.grads().items():
        assert rel(digest(gr.numpy()), g["grad/" + key]) < 1e-8, key
    for key in g.files:
        if key.startswith("running/"):
            assert rel(port.P[key[8:]].detach().numpy(), g[key]) < 1e-9, key
