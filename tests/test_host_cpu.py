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
    # and the ctypes table covers exactly the header
    bound = set(_lib.SIGNATURES) | set(_lib.PLAIN)
    assert bound == set(names), (bound ^ set(names))
    assert lib.vp_abi_version() == 1
    assert lib.vp_launch_count() == 0


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
    w = np.asarray(w).ravel()
    out = np.empty((p.taps, p.n, p.k), w.dtype)
    for t in range(p.taps):
        for n in range(p.n):
            out[t, n, :] = w[n * p.sn + np.arange(p.k) * p.sk + t * p.st]
    return out


def test_pack_recipes():
    from vae_play_b200.functional import TapLayer
    rs = np.random.RandomState(0)
    # conv: Wp[t][co][ci] = w[co][ci][ky][kx]
    w = rs.randn(6, 4, 3, 3)
    L = TapLayer("conv", 4, 6, k=3, stride=1, pad=1)
    assert np.array_equal(_emulate_pack(w, L.p_fwd), w.reshape(6, 4, 9).transpose(2, 0, 1))
    assert np.array_equal(_emulate_pack(w, L.p_dgrad), w.reshape(6, 4, 9).transpose(2, 1, 0))
    # convT: weight [ci][co][ky][kx]
    w = rs.randn(4, 6, 3, 3)
    L = TapLayer("convT", 4, 6, k=3, stride=2, pad=1, out_pad=1)
    assert np.array_equal(_emulate_pack(w, L.p_fwd), w.reshape(4, 6, 9).transpose(2, 1, 0))
    assert np.array_equal(_emulate_pack(w, L.p_dgrad), w.reshape(4, 6, 9).transpose(2, 0, 1))
    # flatten_in: Linear weight [out, C*S*S] with NCHW flatten  ->  taps over the SxS map
    C_, S, out = 3, 2, 5
    w = rs.randn(out, C_ * S * S)
    L = TapLayer("flatten_in", C_, out, spatial=S)
    assert np.array_equal(_emulate_pack(w, L.p_fwd), w.reshape(out, C_, S * S).transpose(2, 0, 1))
    # dgrad packing [T][C][out] read as a plain [T*C, out] matrix maps dh -> channels-last dx
    wp = _emulate_pack(w, L.p_dgrad).reshape(S * S * C_, out)
    dh = rs.randn(2, out)
    dx_cl = dh @ wp.T                                     # [B, (t, c)]
    dx_ref = (dh @ w).reshape(2, C_, S * S).transpose(0, 2, 1).reshape(2, -1)
    assert np.allclose(dx_cl, dx_ref)
    # flatten_out: Linear weight [C*S*S, z], output viewed [B,C,S,S]
    z = 4
    w = rs.randn(C_ * S * S, z)
    L = TapLayer("flatten_out", z, C_, spatial=S)
    wp = _emulate_pack(w, L.p_fwd).reshape(S * S * C_, z)
    zz = rs.randn(2, z)
    y_cl = zz @ wp.T
    y_ref = (zz @ w.T).reshape(2, C_, S * S).transpose(0, 2, 1).reshape(2, -1)
    assert np.allclose(y_cl, y_ref)
    wpd = _emulate_pack(w, L.p_dgrad)                    # [T][z][C]
    dy_cl = rs.randn(2, S * S, C_)
    dz = np.einsum("btc,tkc->bk", dy_cl, wpd)
    dz_ref = dy_cl.transpose(0, 2, 1).reshape(2, -1) @ w
    assert np.allclose(dz, dz_ref)


def test_philox_policy_matches_oracle():
    from oracle import philox
    from vae_play_b200.functional import philox_policy
    for n in (1, 255, 256, 32768, 303104, 303105, 5_000_000):
        assert philox_policy(n, 148) == philox.aten_normal_policy(n, 148)
