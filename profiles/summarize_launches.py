"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (per-launch times are
cold-cache and serialised: read the SHARES, not the absolutes).  Usage: python profiles/summarize_launches.py file.csv [last_n]"""
import collections
import csv
import re
import sys


def main(path, last_n=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if last_n and str(last_n).startswith("step:"):
        # one whole training step: the launches after the second-to-last optimiser kernel up to and including the last one
        key = str(last_n)[5:]
        idx = [i for i, r in enumerate(rows) if key in r["Kernel Name"]]
        if len(idx) >= 2:
            rows = rows[idx[-2] + 1: idx[-1] + 1]
    elif last_n:
        rows = rows[-int(last_n):]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in rows:
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("vp::", "").replace("(anonymous namespace)::", "")
        name = name.replace("<unnamed>::", "")
        agg[name[:80] + " grid=" + row["Grid Size"].replace(" ", "")][1] += v
        agg[name[:80] + " grid=" + row["Grid Size"].replace(" ", "")][0] += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot:.1f} us total (serialised, cold cache)")
    print(f"{'kernel':100s} {'n':>5s} {'us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{k[:100]:100s} {v[0]:5d} {v[1]:10.1f} {100 * v[1] / tot:6.1f}%")
    fam = collections.defaultdict(float)
    for k, v in agg.items():
        key = ("tcgen05 contraction (fwd/dgrad)" if any(s_ in k for s_ in ("tapgemm_tc", "tapgemm_win", "tapgemm_pair", "tapgemm_gwin", "splitk"))
               else "tcgen05 contraction (wgrad)" if "tapwgrad" in k
               else "tcgen05 thin layers" if "thin_" in k
               else "cuda-core contraction" if "simt" in k
               else "norm/act" if any(s_ in k for s_ in ("norm_stream", "stats_kernel", "apply_kernel", "bwd_reduce", "finalize", "colsum", "bn_rows", "bwd_finish"))
               else "optimiser" if ("rmsprop" in k or "adam" in k)
               else "pack/cast/layout" if any(s_ in k for s_ in ("pack", "nchw", "nhwc", "cast", "transpose"))
               else "torch" if k.startswith("at::") else "other (loss, reparam, ...)")
        fam[key] += v[1]
    print("# by family")
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
        print(f"{k:40s} {v:10.1f} us {100 * v / tot:6.1f}%")


if __name__ == "__main__":
    main(*sys.argv[1:])
