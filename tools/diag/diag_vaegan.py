"""Diagnostic (not a test): per-tensor deviation of the mirror VaeGan step from the reference fixtures, both precisions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import vae_play_b200 as vp
from tests import test_gpu_vaegan as T
from tests.util import load, rel, rel_l2

for b in (4, 16):
    g = load(f"vaegan64_b{b}.npz")
    dev = dict(zip([str(k) for k in g["ref_fp32_dev_keys"]], [float(v) for v in g["ref_fp32_dev_vals"]]))
    for prec in ("fp32", "bf16"):
        net = T.build(vp, prec)
        out, single = T.run_step(net, batch=b)
        print(f"==== batch {b} {prec}")
        for k in ("x_tilde", "disc_class", "mus", "log_variances", "params", "kl", "mse", "losses"):
            print(f"  out {k:28s} rel {rel(T.npy(out[k]).reshape(g[k].shape), g[k]):.3e}  refdev {dev.get(k, 0):.2e}")
        rows = []
        for k, p in list(net.named_parameters()):
            want = g["grad/" + k]
            got = T.npy(p.grad)
            if bool(g["gradfull/" + k][0]):
                r, r2 = rel(got, want), rel_l2(got, want)
            else:
                d = T.digest(got)
                r, r2 = float(np.abs(d[3:] - want[3:]).max() / want[2]), abs(d[1] - want[1]) / want[1]
            rows.append((r, r2, dev[k], k))
        for k, got in single.items():
            want = g["grad/" + k]
            if bool(g["gradfull/" + k][0]):
                r, r2 = rel(got, want), rel_l2(got, want)
            else:
                d = T.digest(got)
                r, r2 = float(np.abs(d[3:] - want[3:]).max() / want[2]), abs(d[1] - want[1]) / want[1]
            rows.append((r, r2, dev[k], k))
        for r, r2, dv, k in sorted(rows, key=lambda t: -t[0] / max(1e-5, 3 * t[2]))[: (25 if prec == "fp32" else 90)]:
            print(f"  grad {k:60s} max-rel {r:.3e} l2/norm {r2:.3e} refdev {dv:.2e} ratio {r / max(1e-5, 3 * dv):.2f}")
vp.set_precision("bf16")
