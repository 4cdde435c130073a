"""CPU-only checks of the host side: the C-ABI library builds, loads and exports every symbol the
header declares; the nn.Module mirror keeps the reference's state_dict keys; packing recipes and the
Philox policy are right.  No kernels are launched here (no GPU in this container)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "vaeplay_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vp_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_header_symbol():
    from vae_play_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vaeplay_b200.h but not exported"
    # the ctypes table is generated from the header prototypes and covers exactly the header
    bound = set(_lib.SIGNATURES) | set(_lib.PLAIN)
    assert bound == set(names), (bound ^ set(names))
    assert lib.vp_abi_version() == _lib.ABI_VERSION == 2
    assert lib.vp_launch_count() == 0
    # debug probes are not part of the release ABI
    assert not any(n.startswith("vp_debug") for n in names)
    assert not hasattr(lib, "vp_debug_umma_probe")


def test_ctypes_signatures_follow_header_prototypes():
    """Argument TYPES, not just names: the binding is derived from the prototypes, and a few hand-written expectations pin
    the derivation itself (pointer / int64 / float / size_t / host out-parameter)."""
    import ctypes as C
    from vae_play_b200 import _lib
    G, p, i, i64, u64, f = C.POINTER(_lib.VpConvGeom), C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float
    expect = {
        "vp_conv_fwd": [G, p, p, p, p, i, i, i, f, i, p],
        "vp_pack_weight": [p, p, i, i, i, i, i64, i64, i64, p],
        "vp_reparam_kl_fwd": [p, p, i64, p, u64, u64, p, i, p, i, p, p, i64, i, p],
        "vp_set_workspace": [p, C.c_size_t],
        "vp_conv_fwd_cl_stats": [G, p, p, p, p, i, C.POINTER(C.c_int), p],
        "vp_axpy": [f, p, p, i64, p],
        "vp_rmsprop_step_shadow": [p, p, p, p, p, i, f, f, f, f, i, p],
    }
    for name, want in expect.items():
        assert _lib.SIGNATURES[name] == want, name
    lib = _lib.load()
    for name, args in _lib.SIGNATURES.items():
        assert list(getattr(lib, name).argtypes) == args and getattr(lib, name).restype is C.c_int, name
    # every prototype in the header has as many parameters as the parser found (no silently skipped declaration)
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vaeplay_b200.h")).read(), flags=re.S)
    for name, args in _lib.SIGNATURES.items():
        m = re.search(name + r"\s*\(([^)]*)\)", src)
        assert m and len([a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]) == len(args), name


def test_stale_library_is_detected(tmp_path, monkeypatch):
    from vae_play_b200 import build
    assert not build.is_stale()
    monkeypatch.setattr(build, "DIGEST", str(tmp_path / "none.sha256"))
    assert build.is_stale()


def test_no_cpu_path():
    import vae_play_b200 as vp
    from vae_play_b200 import _lib
    with pytest.raises(_lib.VaePlayError):
        vp.mse_loss(torch.zeros(8), torch.zeros(8))
    with pytest.raises(_lib.VaePlayError):
        vp.reparam_kl(torch.zeros(2, 4), torch.zeros(2, 4))


def test_state_dict_keys_match_reference():
    from vae_play_b200.models.networks import VaeGan
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    for img in ("64", "128"):
        m = VaeGan(int(img), 128)
        sd = m.state_dict()
        assert list(sd.keys()) == [k for k, _ in want[img]]
        assert [list(v.shape) for v in sd.values()] == [s for _, s in want[img]]


def test_init_parameters_matches_reference_scale():
    from vae_play_b200.models.networks import VaeGan
    torch.manual_seed(0)
    m = VaeGan(64, 128)
    w = m.encoder.conv[1].conv.weight
    s = 1.0 / np.sqrt(np.prod(w.shape[1:])) / np.sqrt(3)   # reference networks.py:219-224
    assert float(w.abs().max()) <= s and float(w.abs().max()) > 0.98 * s
    assert float(m.decoder.conv[3][0].bias.abs().max()) == 0.0


def _emulate_pack(w, p):
    """w: flat storage of the weight (physical element order)."""
    w = np.asarray(w).ravel()
    out = np.empty((p.taps, p.n, p.k), w.dtype)
    for t in range(p.taps):
        for n in range(p.n):
            out[t, n, :] = w[n * p.sn + np.arange(p.k) * p.sk + t * p.st]
    return out


def _storage(w):
    return torch.as_strided(w, (w.numel(),), (1,)).numpy()


@pytest.mark.parametrize("channels_last", [False, True])
def test_pack_recipes(channels_last):
    """Pack recipes are derived from the weight's strides: the same panels come out of a torch-contiguous and of a
    channels-last weight."""
    from vae_play_b200.functional import TapLayer
    torch.manual_seed(0)
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    # conv: Wp[t][co][ci] = w[co][ci][ky][kx]
    w = torch.randn(6, 4, 3, 3, dtype=torch.float64).contiguous(memory_format=fmt)
    L = TapLayer("conv", 4, 6, k=3, stride=1, pad=1)
    ref = w.numpy().reshape(6, 4, 9)
    assert np.array_equal(_emulate_pack(_storage(w), L._recipe("fwd", w)), ref.transpose(2, 0, 1))
    assert np.array_equal(_emulate_pack(_storage(w), L._recipe("dgrad", w)), ref.transpose(2, 1, 0))
    assert np.array_equal(_emulate_pack(_storage(w), L._recipe("wgrad", w)), ref.transpose(2, 0, 1))
    # convT: weight [ci][co][ky][kx]
    w = torch.randn(4, 6, 3, 3, dtype=torch.float64).contiguous(memory_format=fmt)
    L = TapLayer("convT", 4, 6, k=3, stride=2, pad=1, out_pad=1)
    ref = w.numpy().reshape(4, 6, 9)
    assert np.array_equal(_emulate_pack(_storage(w), L._recipe("fwd", w)), ref.transpose(2, 1, 0))
    assert np.array_equal(_emulate_pack(_storage(w), L._recipe("dgrad", w)), ref.transpose(2, 0, 1))
    assert np.array_equal(_emulate_pack(_storage(w), L._recipe("wgrad", w)), ref.transpose(2, 0, 1))
    # linear: weight [out][in]
    w = torch.randn(5, 7, dtype=torch.float64)
    L = TapLayer("linear", 7, 5)
    assert np.array_equal(_emulate_pack(_storage(w), L._recipe("fwd", w))[0], w.numpy())
    assert np.array_equal(_emulate_pack(_storage(w), L._recipe("dgrad", w))[0], w.numpy().T)


def test_weights_channels_last_keeps_state_dict():
    """Conv weights are kept in channels-last memory order: shapes, values and state_dict keys are unchanged and a
    checkpoint written from them loads into a plain contiguous module."""
    from vae_play_b200.models.networks import DecoderBlock, EncoderBlock
    torch.manual_seed(0)
    blk, ref = EncoderBlock(64, 128), torch.nn.Conv2d(64, 128, 5, padding=2, stride=2, bias=False)
    assert blk.conv.weight.shape == ref.weight.shape
    assert blk.conv.weight.is_contiguous(memory_format=torch.channels_last) and not blk.conv.weight.is_contiguous()
    ref.load_state_dict({"weight": blk.state_dict()["conv.weight"]})
    assert torch.equal(ref.weight, blk.conv.weight)
    blk.load_state_dict({**blk.state_dict(), "conv.weight": ref.weight.detach() * 2})
    assert blk.conv.weight.is_contiguous(memory_format=torch.channels_last) and torch.equal(blk.conv.weight, ref.weight * 2)
    assert DecoderBlock(128, 64).conv.weight.is_contiguous(memory_format=torch.channels_last)
    assert EncoderBlock(1, 64).conv.weight.is_contiguous()          # single-channel layers stay torch-contiguous (thin kernels)


def test_philox_policy_matches_oracle():
    from oracle import philox
    from vae_play_b200.functional import philox_policy
    for n in (1, 255, 256, 32768, 303104, 303105, 5_000_000):
        assert philox_policy(n, 148) == philox.aten_normal_policy(n, 148)


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the reference's CPU path on the host cores) prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-batch", "4"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "train_images_per_sec" and line["unit"] == "images/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    # "reference": the unmodified reference modules vendored into the git-ignored oracle/_ref by oracle/build_ref.py
    # (__graft_entry__.build() runs it where /root/reference exists); "port": oracle/vae_torch.py when they are absent
    have_ref = os.path.exists(os.path.join(root, "oracle", "_ref", "models", "networks.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_and_port_cpu_steps_agree():
    """The unmodified reference modules (oracle/_ref, when vendored) and the line-by-line port give the same loss on the same
    weights / batch / eps: the two possible CPU-baseline legs of bench.py time the same computation."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "oracle", "_ref", "models", "networks.py")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not vendored here (python oracle/build_ref.py needs /root/reference)")
    import bench
    from oracle import vae_numpy as vn
    from oracle.vae_torch import VaeTorchPort
    ref = bench._ReferenceCpuStep(bench._load_reference_modules(), 64, 1)
    port = VaeTorchPort(vn.synth_vae_params(64, 128, 1, 1, 0), torch.float32)
    x_np, eps_np = vn.synth_batch(4, 64, 1, 128, 0)
    x, eps = torch.from_numpy(x_np), torch.from_numpy(eps_np)
    orig = torch.Tensor.normal_
    torch.Tensor.normal_ = lambda self, *a, **k: self.copy_(eps)
    try:
        l_ref = float(ref.step(x))
    finally:
        torch.Tensor.normal_ = orig
    l_port = float(port.step(x, eps)[0])
    assert abs(l_ref - l_port) / abs(l_port) < 1e-5


def test_engine_options_match_bench_flags():
    """vae_play_b200.engine.VaeTrainer takes bench.py's argparse namespace as its options: every option it reads must exist as
    a bench flag with the same default, so `bench.py` (and tools/dp_timeline.py) really run the packaged step."""
    import argparse
    import bench
    from vae_play_b200.engine import VaeTrainer
    ap = argparse.ArgumentParser()
    bench.add_arguments(ap)
    args = vars(ap.parse_args([]))
    for k, v in VaeTrainer.default_options().items():
        assert k in args, f"bench.py has no flag for engine option {k!r}"
        assert args[k] == v, (k, args[k], v)


def test_grad_buckets_breaks_keep_stage_groups_apart():
    """GradBuckets(breaks=[p]): p starts a bucket of its own run (reverse registration order), so a backward cut into stages can
    exchange a stage's gradients without waiting for parameters of the next stage that would otherwise share the bucket."""
    import torch
    from vae_play_b200.parallel import GradBuckets
    enc = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))
    dec = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Linear(8, 8))
    params = list(enc.parameters()) + list(dec.parameters())
    plain = GradBuckets(params, 1, bucket_mb=1.0)
    assert len(plain.buckets) == 1                                   # everything fits one bucket ...
    plain.remove()
    gb = GradBuckets(params, 1, bucket_mb=1.0, breaks=[list(enc.parameters())[-1]])
    try:
        assert len(gb.buckets) == 2                                  # ... unless the encoder's last parameter must start one
        dec_ids, enc_ids = {id(p) for p in dec.parameters()}, {id(p) for p in enc.parameters()}
        assert {id(p) for p in gb.buckets[0]["params"]} == dec_ids and {id(p) for p in gb.buckets[1]["params"]} == enc_ids
        assert gb.buckets_within(list(dec.parameters())) == [0] and gb.buckets_within(params) == [0, 1]
    finally:
        gb.remove()
