#!/usr/bin/env python
"""bench.py -- headline benchmark of fluid2d_b200 (contract: see the task statement).

Metric (BASELINE.json): full-step cell-updates/s of the Euler model -- RK3_SSP, 5th-order
upwind advection with parabolic flux splitting, two truncated multigrid inversions per
step plus the end-of-step full solve, diagnostics -- on the freedecay initial state
(experiments/Twodim_turbulence/freedecay/freedecay.py, seed 42), doubly periodic,
4096^2 per GPU.  A "step" is one iteration of Fluid2d.loop(): set_dt, model.step,
diagnostics.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 4096] [--tracers 1]
    python bench.py --strong --n 16384 --gpus N --no-cpu     # SURVEY.md 8d case S5 (not the driver's line)
    python bench.py --impl reference ...     # the CPU arm (oracle port on the host cores)

One JSON line on stdout (rank 0); everything else goes to stderr.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))

# algorithmic bytes per cell per Euler RK3_SSP step:
# T advected tracers, n_F F-cycles in the end-of-step solve
def b_alg(T, n_F):
    return 1114.2 + 144.*T + 272.5*n_F


SMOOTH_BYTES_PER_CELL = 25.   # one Grid.smooth application: read x, b, mask; write x
VCYCLE_BYTES_PER_CELL = 139.3


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650., "fallback (B200_PROFILING.md)"


_SAMPLER_SRC = r"""
import sys, time
import pynvml as N
N.nvmlInit()
h = N.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
names = [("hw_slowdown", "HwSlowdown"), ("hw_thermal_slowdown", "HwThermalSlowdown"),
         ("sw_thermal_slowdown", "SwThermalSlowdown"), ("sw_power_cap", "SwPowerCap")]
bits = []
for nm, suffix in names:
    v = getattr(N, "nvmlClocksEventReason" + suffix, None)
    if v is None:
        v = getattr(N, "nvmlClocksThrottleReason" + suffix, None)
    if v is not None:
        bits.append((nm, v))
print("max", N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM), flush=True)
while True:
    try:
        r = N.nvmlDeviceGetCurrentClocksEventReasons(h)
    except Exception:
        r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    print("s", repr(time.time()), N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM),
          ",".join(nm for nm, b in bits if r & b), flush=True)
    time.sleep(0.005)
"""


class ClockSampler(object):
    """SM clock and throttle reasons sampled every 5 ms through NVML (nvidia_ml_py) by a CHILD
    process, so that the sampling costs the benchmark's host thread nothing (a sampling thread
    in this process was measured to cost 5 % at 4 GPUs through the GIL).  start() before the
    warm-up; window(t0, t1) keeps the samples taken during the timed region."""

    def __init__(self, index=0):
        self.proc = None
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                index = int(vis.split(",")[index])
            except (ValueError, IndexError):
                pass
        try:
            # the child writes to a file: nothing in this process wakes up while the bench runs
            self.out = tempfile.NamedTemporaryFile("w+", suffix=".clocks")
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(index)], stdout=self.out,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_ready(self, timeout=10.):
        """block until the child has delivered its first sample (NVML start-up takes a while)"""
        t_end = time.time()+timeout
        while self.proc is not None and time.time() < t_end:
            if os.path.getsize(self.out.name) > 40:
                return
            time.sleep(0.05)

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml sampler unavailable"]}
        time.sleep(0.02)
        self.proc.terminate()
        self.proc.wait()
        self.out.seek(0)
        sm, mx, reasons = [], None, set()
        for ln in self.out.read().splitlines():
            f = ln.split()
            if f and f[0] == "max":
                mx = float(f[1])
            elif f and f[0] == "s" and t0 <= float(f[1]) <= t1:
                sm.append(float(f[2]))
                if len(f) > 3:
                    reasons.update(x for x in f[3].split(",") if x)
        sm.sort()
        return {"sm_mhz": sm[len(sm)//2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def build_case(api, n, tracers, datadir, world=1, strong=False):
    """world > 1: weak scaling -- the global domain is n x (n*world), split in `world`
    y-slabs of n x n cells (npx = 1, npy = world); the n x n freedecay field is repeated in
    every slab (periodic tiling), so each GPU carries the single-GPU workload plus the halo
    exchange with its neighbours.
    strong=True (--strong, SURVEY.md 8d case S5): the global domain is n x n whatever the
    number of GPUs, split in `world` y-slabs of n/world rows; the field is a 4096^2 freedecay
    tile repeated over the domain"""
    import cases
    if strong:
        return cases.freedecay(api, datadir, n, order=5, tracer=(tracers > 1), ny=n, npy=world,
                               tile=min(n, 4096))
    if world == 1:
        return cases.freedecay(api, datadir, n, order=5, tracer=(tracers > 1))
    return cases.freedecay(api, datadir, n, order=5, tracer=(tracers > 1), ny=n*world, npy=world, tile=True)


def loop_body(f2d):
    """one iteration of Fluid2d.loop() (fluid2d.py:235-280) without I/O"""
    model = f2d.model
    f2d.set_dt(f2d.kt)
    model.step(f2d.t, f2d.dt)
    f2d.t += f2d.dt
    f2d.kt += 1
    model.diagnostics(model.var, f2d.t)


# ---------------------------------------------------------------------------
# CPU arm: the oracle port (the reference's Fortran cannot be compiled here, so this is
# kind "port"), all host threads, a bounded number of steps of the SAME workload
# ---------------------------------------------------------------------------
def cpu_arm(n, tracers, steps, warmup):
    import types
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host
    # thread (the OpenMP runtime reads the variable when the oracle library is loaded, below)
    under_torchrun = os.environ.get("OMP_NUM_THREADS") == "1" and (
        "TORCHELASTIC_RUN_ID" in os.environ or "RANK" in os.environ)
    if under_torchrun:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import model as om, kernels as K
    K.lib().f2d_oracle_set_reduce_mode(1)     # row-partial reductions (parallel)
    if under_torchrun:
        K.lib().f2d_oracle_set_num_threads(os.cpu_count() or 1)
    cores = K.lib().f2d_oracle_num_threads()

    class F(om.Fluid2d):
        def __init__(self, p, g):
            om.Fluid2d.__init__(self, p, g, fast_axpy=True)
    api = types.SimpleNamespace(Param=om.Param, Grid=om.Grid, Fluid2d=F)
    t0 = time.time()
    f2d = build_case(api, n, tracers, tempfile.mkdtemp())
    f2d.model.diagnostics(f2d.model.var, 0.)
    log("[cpu] set-up %.1f s on %d threads" % (time.time()-t0, cores))
    for _ in range(warmup):
        loop_body(f2d)
    t0 = time.time()
    for _ in range(steps):
        loop_body(f2d)
    dt = time.time()-t0
    return {"value": n*n*steps/dt, "unit": "cell-updates/s", "cores": cores, "kind": "port",
            "sample": "%d steps (after %d warm-up) of the same %dx%d Euler freedecay workload, oracle C port "
                      "with OpenMP on all host threads" % (steps, warmup, n, n),
            "ms_per_step": 1e3*dt/steps}


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 3)
    warm = min(args.warmup, 1)
    r = cpu_arm(args.n, args.tracers, steps, warm)
    line = {"impl": "reference", "metric": "cell_updates_per_s", "value": r["value"], "unit": r["unit"],
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "Euler freedecay %dx%d perio, RK3_SSP, upwind5 parabolic, T=%d"
                       % (args.n, args.n, args.tracers)},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def gpu_main(args):
    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # native libraries (NCCL's version banner) write to file descriptor 1 directly: point it at
    # stderr until the JSON line is due, so that stdout carries that one line only
    sys.stdout.flush()
    saved_fd1 = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.replicas:
        os.environ.pop("RANK", None)     # N independent single-GPU replicas (debug aid)
    import fluid2d_b200
    api = fluid2d_b200.api()
    from runtime import rt
    r = rt()
    lib = r.lib
    real_stdout = sys.stdout
    sys.stdout = sys.stderr
    sampler = ClockSampler(local) if rank == 0 else None     # child process, started well before the timed region
    n, T = args.n, args.tracers
    t0 = time.time()
    slabs = world > 1 and not args.replicas
    strong = bool(args.strong)
    if strong and (args.replicas or n % world):
        raise SystemExit("--strong: n must be a multiple of the number of GPUs (and no --replicas)")
    rows = n//world if strong else n          # rows of this rank's slab
    total_cells = n*n if strong else world*n*n
    f2d = build_case(api, n, T, tempfile.mkdtemp(), world if slabs else 1, strong)
    model = f2d.model
    model.diagnostics(model.var, 0.)
    torch.cuda.synchronize()
    log("[gpu %d] set-up %.1f s" % (rank, time.time()-t0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        loop_body(f2d)
    if sampler:
        sampler.wait_ready()
    barrier()
    # ---- timed region: K steps, device timers, inputs (134 MB fields) exceed the 126 MB L2
    lib.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nites = []
    wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        loop_body(f2d)
        nites.append(model.ope.last_solve[0])
    e1.record()
    barrier()
    wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = int(lib.launch_count())
    clocks = sampler.window(wall0, wall1) if sampler else None
    if world > 1:
        tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    ms_step = ms/args.steps
    value = total_cells*args.steps/(ms*1e-3)
    n_F = float(np.mean(nites))

    # ---- dominant kernel: the level-0 double Jacobi sweep (Grid.smooth), timed alone
    peak, peak_src = measured_peaks()
    if slabs:
        # kernel-level numbers come from a single-GPU hierarchy of the slab's size on this rank
        cm = torch.ones((rows+6, n+6), dtype=torch.float64, device="cuda")
        cm[-1, :] = 0
        cm[:, -1] = 0
        import ctypes
        mgh = ctypes.c_void_p()
        lib.mg_create(ctypes.byref(mgh), r.ptr(cm), rows+6, n+6, 1./n, 1./n, 8./9., 1., 0., r.stream)
    else:
        mgh = model.ope.gmg.h
    x0 = torch.zeros((rows+6, n+6), dtype=torch.float64, device="cuda")
    b0 = torch.randn((rows+6, n+6), dtype=torch.float64, device="cuda")
    reps = 10
    for _ in range(3):
        lib.mg_smooth(mgh, 0, r.ptr(x0), r.ptr(b0), 2, r.stream)
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(reps):
        lib.mg_smooth(mgh, 0, r.ptr(x0), r.ptr(b0), 2, r.stream)   # 2 applications: x -> t -> x
    s1.record()
    torch.cuda.synchronize()
    smooth_ms = s0.elapsed_time(s1)/(2*reps)
    achieved = SMOOTH_BYTES_PER_CELL*rows*n/(smooth_ms*1e-3)/1e9
    # ---- V-cycle (metric part 2): one Vcycle(0) through its CUDA graph
    lib.mg_vcycle(mgh, 0, r.stream)
    torch.cuda.synchronize()
    s0.record()
    for _ in range(10):
        lib.mg_vcycle(mgh, 0, r.stream)
    s1.record()
    torch.cuda.synchronize()
    vcycle_ms = s0.elapsed_time(s1)/10

    # ---- end to end: host buffers in, host buffers out, every step
    ds = model.var.dstate
    nvar = ds.nvar
    _ = model.var.state            # materialise the pinned mirror
    for _ in range(2):
        ds._host_written(None)
        loop_body(f2d)
        _ = model.var.state
    barrier()
    h2d0, d2h0 = ds.h2d_bytes, ds.d2h_bytes
    ke = min(args.steps, 5)
    e0.record()
    for _ in range(ke):
        ds._host_written(None)     # the caller's (pinned) host state is the step's input ...
        loop_body(f2d)             # ... uploaded here (H2D of every field) ...
        _ = model.var.state        # ... and the new state is read back (D2H of every field)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms = float(tmax.item())
    e2e = {"value": total_cells*ke/(e2e_ms*1e-3), "unit": "cell-updates/s",
           "h2d_bytes_per_step": (ds.h2d_bytes-h2d0)//ke, "d2h_bytes_per_step": (ds.d2h_bytes-d2h0)//ke,
           "ms_per_step": e2e_ms/ke, "steps": ke}

    sys.stderr.flush()
    os.dup2(saved_fd1, 1)
    os.close(saved_fd1)
    sys.stdout = real_stdout
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu and not strong:
        try:
            cpu = cpu_arm(n, T, 2, 1)
        except Exception as ex:   # the baseline is a report, never a reason to lose the GPU line
            cpu = {"value": None, "unit": "cell-updates/s", "cores": 0, "kind": "port", "sample": "failed: %r" % ex}
    balg = b_alg(T, n_F)
    line = {
        "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Euler freedecay %dx%d perio %s (experiments/Twodim_turbulence), RK3_SSP, "
                               "upwind5 + parabolic splitting, 2 truncated MG inversions + full solve per step, "
                               "T=%d advected tracer(s)" % (n, n, "in total" if strong else "per GPU", T),
                   "grid": [n, n], "tracers": T, "n_F_mean": n_F,
                   "parallelism": "single GPU" if world == 1 else (
                       "%d y-slabs of %dx%d (global %dx%d), peer halo exchange over NVLink, coarse levels gathered"
                       % (world, n, rows, n, rows*world) if slabs else "%d independent replicas" % world),
                   "mg_slab_levels": getattr(model.ope.gmg, "slab_levels", 0),
                   "cache": "working set %.1f GB >> 126 MB L2 (no flush needed)" % (40*(n+6)*(rows+6)*8/1e9)},
        "roofline": {"bound": "hbm", "kernel": "k_smooth2<0,0,0> (Grid.smooth = double Jacobi sweep + halo fill, level 0)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved/peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch at 4096^2, from
                     # profiles/r01_ncu_full_v10_two_vcycle_kernels.csv (269.1 MB + 104.2 MB; algorithmic
                     # 420 MB: the mask-free level reads x and b, writes x; halo re-reads hit L2)
                     "traffic": 3.733e8 if (n == 4096 and rows == 4096) else None,
                     "peak_source": peak_src, "ms_per_launch": smooth_ms,
                     "algorithmic_bytes_per_cell": SMOOTH_BYTES_PER_CELL},
        "step_hbm": {"b_alg_bytes_per_cell": balg, "achieved_gbs": balg*value/world/1e9,
                     "frac_of_peak": balg*value/world/1e9/peak},
        "vcycle_ms": vcycle_ms,
        "vcycle_frac_of_peak": VCYCLE_BYTES_PER_CELL*rows*n/(vcycle_ms*1e-3)/1e9/peak,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": launches,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--tracers", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--replicas", action="store_true", help="N>1: independent replicas instead of slabs")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: --n is the GLOBAL grid (e.g. 16384), split in --gpus y-slabs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        reference_main(args)
    else:
        gpu_main(args)


if __name__ == "__main__":
    main()
