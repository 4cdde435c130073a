"""Condensed SASS listing of an .ncu-rep's first kernel: every executed instruction with its stall-sample and execution counts,
repeated unrolled mbarrier polls folded.  usage: python tools/ncu_sass_regions.py report.ncu-rep [min_exec]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, data, seen = None, [], 0
for r in rows:
    if r and r[0] == "Address":
        seen += 1
        h = {k: i for i, k in enumerate(r)}
        continue
    if h and seen == 1 and len(r) > h["# Samples"] and r[h["# Samples"]].isdigit():
        data.append((int(r[h["# Samples"]]), int(r[h["Instructions Executed"]] or 0), r[h["Source"]].strip()))
tot = sum(d[0] for d in data)
print(f"# {len(data)} SASS instructions, {tot} stall samples, {sum(d[1] for d in data)} warp-instructions executed")
prev, rep_n, rep_s, rep_e = None, 0, 0, 0
def flush():
    global prev, rep_n, rep_s, rep_e
    if prev is not None and rep_n > 1:
        print(f"        ... x{rep_n} polls, samples {rep_s}, exec {rep_e}")
    prev, rep_n, rep_s, rep_e = None, 0, 0, 0
for i, (s, e, t) in enumerate(data):
    if e == 0 and s == 0:
        continue
    key = t if ("TRYWAIT" in t) else None
    if key and key == prev:
        rep_n += 1; rep_s += s; rep_e += e
        continue
    if ("BRA" in t and prev):
        rep_s += s; rep_e += e
        continue
    flush()
    print(f"{i:5d} s={s:4d} e={e:8d} {t[:110]}")
    if key:
        prev, rep_n, rep_s, rep_e = key, 1, s, e
flush()
