"""Diagnostic (not a test): per-tensor relative error of the CUDA VAE step vs the NumPy oracle (ReLU pattern pinned)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import vae_play_b200 as vp
from oracle import vae_numpy as vn
from tests.test_gpu_parity import build_vae, run_step
from tests.util import rel, rel_l2

for mode in ("bf16",):
    vp.set_precision(mode)
    for (img, cin, b, seed) in [(64, 1, 4, 0), (64, 1, 8, 5), (64, 1, 32, 7), (64, 3, 6, 6), (128, 1, 4, 2)]:
        enc, dec, P = build_vae(img, cin, 128, seed)
        x, eps = vn.synth_batch(b, img, cin, 128, seed)
        out, grads, running, acts = run_step(vp, enc, dec, x, eps)
        fwd = vn.vae_step(P, x, eps)
        masks, info = {}, {}
        for name, a in acts.items():
            pre = fwd["acts"][name + ".pre_act"]
            m = a > 0
            masks[name] = m
            bad = m != (pre > 0)
            info[name] = (f"{bad.mean():.1e}", f"{(np.abs(pre[bad]).max() / np.abs(pre).max()) if bad.any() else 0:.1e}",
                          f"{rel(a, fwd['acts'][name + '.out']):.1e}")
        want = vn.vae_step(P, x, eps, masks=masks)
        fw = {k: f"{rel(out[k], want[k]):.1e}" for k in ("mu", "logvar", "z", "x_tilde", "kl", "loss")}
        gr = {k: rel(grads[k], want["grads"][k]) for k in want["grads"]}
        worst = sorted(gr.items(), key=lambda kv: -kv[1])[:5]
        print(mode, (img, cin, b), "fwd", fw)
        print("      masks (flip frac, |pre| at flips, act err)", info)
        print("      grads worst", [(k, f"{v:.1e}") for k, v in worst])
        fw2 = {k: f"{rel_l2(out[k], want[k]):.1e}" for k in ("mu", "logvar", "z", "x_tilde", "kl")}
        gr2 = {k: rel_l2(grads[k], want["grads"][k]) for k in want["grads"]}
        worst2 = sorted(gr2.items(), key=lambda kv: -kv[1])[:5]
        a2 = max(rel_l2(a, fwd["acts"][n + ".out"]) for n, a in acts.items())
        print("      L2: fwd", fw2, "acts max", f"{a2:.1e}", "grads worst", [(k, f"{v:.1e}") for k, v in worst2], flush=True)
