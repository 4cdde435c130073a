"""Summarise an .ncu-rep (read with the ncu CLI, no GPU needed): key raw metrics + top stall-sampled SASS lines.
usage: python tools/ncu_summary.py report.ncu-rep [n_top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("==", name[:100])
    for h, u, v in zip(hdr, units, r):
        if h in KEYS:
            print(f"  {h:70s} {v:>16s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
data = []
for r in rows:
    if r and r[0] == "Address":
        h = {k: i for i, k in enumerate(r)}
        continue
    if h and len(r) > h["# Samples"] and r[h["# Samples"]].isdigit():
        data.append(r)
if h:
    tot = sum(int(r[h["# Samples"]]) for r in data)
    print(f"-- {tot} stall samples over {len(data)} SASS instructions; top {ntop}:")
    for r in sorted(data, key=lambda r: -int(r[h["# Samples"]]))[:ntop]:
        print(f"  {int(r[h['# Samples']]):6d} {100*int(r[h['# Samples']])/max(tot,1):5.1f}%  exec={r[h['Instructions Executed']]:>8s}  {r[h['Source']][:120]}")
