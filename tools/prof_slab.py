"""Per-kernel time of one step of the bench workload on N GPUs (torchrun), from the CUPTI
activity records collected by torch.profiler (works for kernels launched by our .so and
inside CUDA graphs).  Rank 0 prints the table.  Not a bench: profiler overhead included.

    torchrun --nproc-per-node 2 tools/prof_slab.py [n] [steps]
"""
import collections
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
import fluid2d_b200  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
world = int(os.environ.get("WORLD_SIZE", "1"))
api = fluid2d_b200.api()
so = sys.stdout
sys.stdout = sys.stderr
case = os.environ.get("F2D_PROF_CASE", "freedecay")   # rb: RayleighBenard n x n/2; karman: VonKarman 4n x n
if case == "rb":
    import cases
    f2d = cases.rb(api, tempfile.mkdtemp(), n)
elif case == "karman":
    import cases
    f2d = cases.karman(api, tempfile.mkdtemp(), n, ratio=4)
else:
    f2d = bench.build_case(api, n, 1, tempfile.mkdtemp(), world)
f2d.model.diagnostics(f2d.model.var, 0.)
for _ in range(4):
    bench.loop_body(f2d)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        bench.loop_body(f2d)
    torch.cuda.synchronize()
sys.stdout = so
agg = collections.defaultdict(lambda: [0, 0.])
first, last = None, None
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    nm = ev.name
    if "<" in nm and "(" in nm:
        nm = nm.split("(")[0]
    agg[nm][0] += 1
    agg[nm][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    t0, t1 = ev.time_range.start, ev.time_range.end
    first = t0 if first is None else min(first, t0)
    last = t1 if last is None else max(last, t1)
rank = int(os.environ.get("RANK", "0"))
if rank == 0:
    tot = sum(v[1] for v in agg.values())
    print("world %d  n %d  steps %d  kernel time/step %.3f ms  span/step %.3f ms" % (world, n, steps, tot/steps/1e3, (last-first)/steps/1e3))
    for nm, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%8.1f us/step %6.1f%% %6.1f calls/step %7.2f us/call  %s" % (t/steps, 100*t/tot, c/steps, t/c, nm[:90]))
if rank == 0 and os.environ.get("F2D_PROF_SEQ"):
    # kernel sequence of the last step: duration, idle gap before it, name
    evs = sorted([e for e in prof.events() if e.device_type.name == "CUDA"], key=lambda e: e.time_range.start)
    evs = evs[-(len(evs)//steps):]
    prev = None
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", "slab_seq.txt"), "w") as f:
        for e in evs:
            nm = e.name.replace("(anonymous namespace)::", "").replace("void ", "")
            nm = nm.split("(")[0][:48]
            gap = (e.time_range.start-prev) if prev is not None else 0.
            prev = e.time_range.end
            f.write("%8.2f gap %7.2f  %s\n" % (e.time_range.end-e.time_range.start, gap, nm))
if world > 1:
    dist.barrier()