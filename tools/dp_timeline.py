"""Device timeline of ONE data-parallel training step exactly as bench.py runs it (rank 0, torch.profiler / CUPTI): every kernel
with its start offset, duration and stream, so that the position of NCCL's all-reduce kernels against the backward and optimiser
kernels can be read off.  Diagnostic, not a benchmark (CUPTI perturbs short kernels).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/dp_timeline.py [bench.py flags, e.g. --grad-wire fp32 --no-split-backward] > profiles/r02_dp2_timeline.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import bench
import vae_play_b200 as vp


def main():
    ap = argparse.ArgumentParser()
    bench.add_arguments(ap)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        bench.init_nccl(args, dev)
    vp.set_precision(args.precision)
    run = bench.StepRunner(args, args.img, world, rank, dev)
    for _ in range(30):
        run.step(run.x_dev)
    ms = run.time_resident(20) / 20
    N = 4
    run.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(N):
            run.step(run.x_dev)
        run.barrier()
    if rank != 0:
        return
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ev.sort(key=lambda e: e.time_range.start)
    # one step = the kernels between two consecutive launches of the first kernel of the step graph
    first = [i for i, e in enumerate(ev) if e.name == ev[0].name]
    per = max(1, len(first) // N)
    a, b = first[per * (N - 2)], first[per * (N - 1)]
    step = ev[a:b]
    t0 = step[0].time_range.start
    print(f"# world {world}, rank 0, flags {' '.join(sys.argv[1:]) or '(defaults)'}; step without profiler {ms * 1e3:.1f} us; "
          f"profiled step {step[-1].time_range.end - t0:.1f} us, next step starts at {ev[b].time_range.start - t0:.1f} us")
    print("# start_us   dur_us  stream  kernel")
    for e in step:
        k = e.name.replace("void ", "").replace("vp::", "").replace("(anonymous namespace)::", "")
        stream = getattr(e, "stream", None)
        mark = "  <== NCCL" if "nccl" in k.lower() else ""
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {str(stream):>6}  {k[:90]}{mark}")


if __name__ == "__main__":
    main()
