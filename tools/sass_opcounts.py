"""Opcode histogram per kernel of the shipped library (cuobjdump -sass; runs without a GPU).
usage: python tools/sass_opcounts.py [lib.so] > profiles/r02_sass_opcounts.txt"""
import collections, re, subprocess, sys, os
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vae_play_b200", "lib", "libvaeplay_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "ACQBULK", "PREEXIT", "SYNCS", "HMMA", "HGMMA", "FFMA", "REDG", "ATOMG", "ATOMS", "UTCBAR"]
kern, per = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::|vp::", "", kern).split("(")[0]
        per[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)[.\w]*[\s;]", line)
    if m and kern:
        per[kern][m.group(1)] += 1
        per[kern]["_total"] += 1
tot = collections.Counter()
print(f"# {os.path.basename(lib)}: SASS opcode counts per kernel (sm_100a); tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM, TMA = UTMALDG/UTMASTG/UBLKCP,")
print("# griddepcontrol.wait / launch_dependents = ACQBULK / PREEXIT, mbarrier = SYNCS; HMMA / HGMMA (legacy tensor paths) must be absent")
print(f"{'kernel':86s} {'instrs':>7s} " + " ".join(f"{k:>7s}" for k in KEY))
for k, c in per.items():
    print(f"{k[:86]:86s} {c['_total']:7d} " + " ".join(f"{c[x]:7d}" for x in KEY))
    tot.update(c)
print(f"{'TOTAL (' + str(len(per)) + ' kernels)':86s} {tot['_total']:7d} " + " ".join(f"{tot[x]:7d}" for x in KEY))
