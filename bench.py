#!/usr/bin/env python
"""bench.py -- headline benchmark of fluid2d_b200 (contract: see the task statement).

Metric (BASELINE.json): full-step cell-updates/s of the Euler model -- RK3_SSP, 5th-order
upwind advection with parabolic flux splitting, two truncated multigrid inversions per
step plus the end-of-step full solve, diagnostics -- on the freedecay initial state
(experiments/Twodim_turbulence/freedecay/freedecay.py, seed 42), doubly periodic,
4096^2 per GPU.  A "step" is one iteration of Fluid2d.loop(): set_dt, model.step,
diagnostics.

    python bench.py [--steps K] [--warmup W] [--tracers 1]      # N=1: 4096^2 (SURVEY.md 8d case S1)
    torchrun ... bench.py --gpus N ...     # N>1: case S5, 16384^2 GLOBAL split in N y-slabs (strong scaling)
    python bench.py --strong --n 16384     # case S5 on one GPU (the parallel-efficiency denominator)
    torchrun ... bench.py --gpus N --weak  # weak scaling: 4096^2 per GPU (round 1's line)
    python bench.py --impl reference ...   # the CPU arm (oracle port on the host cores)

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


# the other BASELINE.json configurations (SURVEY.md 8d cases S3, S4): bench.py --config rb | vk
CONFIGS = {
    "rb": {"nx": 2048, "ny": 1024, "tracers": 2,
           "workload": "RayleighBenard 2048x1024 (experiments/RayleighBenard): Boussinesq, xchannel, RK3_SSP, upwind5 "
                       "(aparab 0.02) on vorticity and buoyancy, torque, diffusion, no-slip walls, cool-roof forcing; "
                       "4 truncated MG inversions per step"},
    "vk": {"nx": 4096, "ny": 1024, "tracers": 1,
           "workload": "VonKarman 4096x1024 (experiments/VonKarman/karman_street.py): Euler, xchannel, RK3_SSP, upwind3, "
                       "disc obstacle + 2 islands, no-slip, diffusion, sponge layer; 4 truncated MG inversions + full "
                       "solve per step"},
}


def b_alg_config(config, T, n_F):
    """algorithmic bytes per cell per step of the masked configurations, by the rule of SURVEY.md
    8d / appendix B (every operand field of a logical operator read once, every result written
    once, masks 1 B, matrix coefficients 0; truncated inversion = celltocorner 16 + 2 V-cycles
    2 x 139.3 + orthogradient 34 = 328.6).  Derivation: DESIGN.md section 5."""
    inv = 328.6
    if config == "rb":
        # adv 3(16T+17), torque 3x25, diffusion Tx25 and forcing hook 40 at the last stage, 3 stage
        # inversions + the no-slip one, no-slip source 114, RK on all 6 fields 96x6, banom 24, diagnostics 52
        return 3*(16*T+17)+75+25*T+40+4*inv+114+96*6+24+52
    if config == "vk":
        # adv 3x33, diffusion 25, 3 stage inversions + no-slip one + island terms 65 each (5 inversions),
        # full solve 101 + 272.5 n_F, no-slip 114+16, sponge 24, RK 24x3 + 32x3 + 40x2, diagnostics 41
        return 3*33+25+4*inv+5*65+101+272.5*n_F+130+24+248+41
    return b_alg(T, n_F)


def build_config(api, config, datadir):
    import cases
    if config == "rb":
        return cases.rb(api, datadir, CONFIGS["rb"]["nx"])
    return cases.karman(api, datadir, CONFIGS["vk"]["ny"], ratio=CONFIGS["vk"]["nx"]//CONFIGS["vk"]["ny"])


def workload_string(n, T, strong, weak_world=1):
    """config.workload -- the same string in the GPU arm and in the reference arm"""
    where = "in total (global grid, y-slabs over the GPUs)" if strong else (
        "per GPU" if weak_world > 1 else "on one GPU")
    return ("Euler freedecay %dx%d perio %s (experiments/Twodim_turbulence), RK3_SSP, upwind5 + parabolic "
            "splitting, 2 truncated MG inversions + full solve per step, T=%d advected tracer(s)" % (n, n, where, T))


def kernel_bytes_per_cell(name):
    """algorithmic bytes per cell of one launch of a tagged kernel (per
    operator; matrix coefficients count 0, masks 1 B where the instantiation reads them), and
    the number of cells the launch works on -- (None, None) for the latency-bound kernels"""
    import re
    m = re.match(r"k_smooth2<mode(\d),input(\d)(,peer)?> (\d+)x(\d+)", name)
    if m:
        mode, inp = int(m.group(1)), int(m.group(2))
        b = {0: 24., 1: 16., 2: 18., 3: 26.}[inp]+(0. if mode == 1 else 1.)
        return b, int(m.group(4))*int(m.group(5))
    m = re.match(r"k_resid_restrict<mode(\d)(,peer)?> (\d+)x(\d+)", name)
    if m:
        return 18.+(0. if int(m.group(1)) == 1 else 1.), int(m.group(3))*int(m.group(4))
    m = re.match(r"k_resid_sumsq<mode(\d)> (\d+)x(\d+)", name)
    if m:
        return 24.+(0. if int(m.group(1)) == 1 else 1.), int(m.group(2))*int(m.group(3))
    m = re.match(r"k_restrict(<peer>)? (\d+)x(\d+)", name)
    if m:
        return 10., int(m.group(2))*int(m.group(3))
    m = re.match(r"k_zsmooth_rr<mode1> (\d+)x(\d+)", name)
    if m:   # b read once; t written; the coarse right-hand side written (a quarter of the cells)
        return 18., int(m.group(1))*int(m.group(2))
    m = re.match(r"k_adv<upw\d,order\d,masked(\d)> (\d+)x(\d+)", name)
    if m:
        return 32.+float(m.group(1)), int(m.group(2))*int(m.group(3))
    m = re.match(r"f2d_mask_orthogradient_stage<(\d) extra>", name)
    if m:   # psi read and written masked, u, v written; base (and extra tendency) of u, v read, stage u, v written
        return 32.+32.+16.*int(m.group(1)), None
    m = re.match(r"k_map_vec<(\d) in> (\d+) doubles", name)
    if m:
        return 8.*(int(m.group(1))+1), int(m.group(2))
    m = re.match(r"k_elementwise (\d+) doubles", name)
    if m:
        return 16., int(m.group(1))
    m = re.match(r"k_reduce1<nout(\d+)> (\d+)x(\d+)", name)
    if m:
        nout = int(m.group(1))
        # the fused Euler diagnostics read u, v, vorticity, psi, source, xr, yr and the mask once
        # each (f2d_diag_euler; ncu: 960 MB per launch at 4096^2); the single sums one field + mask
        return (57. if nout == 8 else 9.), int(m.group(2))*int(m.group(3))
    return None, None


def kernel_table(lib, r, f2d, peak, shape, nsteps=2):
    """per-kernel accounting of `nsteps` steps run WITHOUT CUDA graphs (f2d_prof_begin /
    f2d_prof_report: a CUDA event behind every launch): share of the step, average duration,
    achieved algorithmic GB/s and fraction of the measured HBM peak for each kernel"""
    import ctypes
    g = f2d.model.ope.gmg
    lib.mg_set_graphs(g.h, 0)
    loop_body(f2d)                              # first ungraphed step: lazy set-up outside the accounting
    lib.prof_begin(r.stream)
    for _ in range(nsteps):
        loop_body(f2d)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.prof_report(buf, len(buf))
    lib.mg_set_graphs(g.h, 1)
    rows = []
    for ln in buf.value.decode().splitlines():
        name, cnt, us = ln.split("\t")
        rows.append([name, int(cnt), float(us)])
    total = sum(x[2] for x in rows) or 1.
    ny, nx = shape
    out = []
    for name, cnt, us in sorted(rows, key=lambda x: -x[2]):
        bpc, cells = kernel_bytes_per_cell(name)
        # f2d_* entry points without a tag work on the whole local grid
        if bpc is None and name in ("f2d_mask_orthogradient", "f2d_celltocorner"):
            bpc, cells = (32., ny*nx) if name == "f2d_mask_orthogradient" else (16., ny*nx)
        if bpc is not None and cells is None:
            cells = ny*nx
        avg = us/cnt
        gbs = bpc*cells/(avg*1e-6)/1e9 if bpc else None
        out.append({"kernel": name, "launches_per_step": cnt/float(nsteps), "share": us/total, "avg_us": avg,
                    "alg_bytes_per_cell": bpc, "cells": cells, "achieved_gbs": gbs,
                    "frac_of_peak": (gbs/peak if gbs else None)})
    return out, total/nsteps


def isolated_times(lib, r, mgh, table, torch, reps=10):
    """re-time the multigrid kernels of the table in isolation: `reps` back-to-back launches of
    the same operator on the hierarchy's own level arrays (f2d_mg_bench_op), CUDA events around
    the batch.  The in-step accounting charges a kernel with the wait for a host that issues
    launches slower than the device runs them (the step runs without graphs there); this does
    not.  Adds 'isolated_us' to the rows it can re-time and recomputes their roofline figures."""
    import ctypes
    import re
    shapes = {}
    ny, nx = ctypes.c_int(), ctypes.c_int()
    for lev in range(lib.mg_nlevels(mgh)):
        lib.mg_level_shape(mgh, lev, ctypes.byref(ny), ctypes.byref(nx))
        shapes[(nx.value-6, ny.value-6)] = lev
    tail0 = lib.mg_tail_level(mgh)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(kind, lev):
        try:
            lib.mg_bench_op(mgh, kind, lev, 2, r.stream)
        except Exception:
            return None
        torch.cuda.synchronize()
        e0.record()
        lib.mg_bench_op(mgh, kind, lev, reps, r.stream)
        e1.record()
        torch.cuda.synchronize()
        return 1e3*e0.elapsed_time(e1)/reps

    for row in table:
        name = row["kernel"]
        us = None
        m = re.match(r"k_smooth2<mode\d,input(\d)(,peer)?> (\d+)x(\d+)", name)
        if m and (int(m.group(3)), int(m.group(4))) in shapes:
            us = timed(int(m.group(1)), shapes[(int(m.group(3)), int(m.group(4)))])
        m = re.match(r"k_resid_restrict<mode\d(,peer)?> (\d+)x(\d+)", name)
        if m and (int(m.group(2)), int(m.group(3))) in shapes:
            us = timed(4, shapes[(int(m.group(2)), int(m.group(3)))])
        m = re.match(r"k_restrict(<peer>)? (\d+)x(\d+)", name)
        if m and (int(m.group(2)), int(m.group(3))) in shapes:
            us = timed(5, shapes[(int(m.group(2)), int(m.group(3)))])
        m = re.match(r"k_zsmooth_rr<mode1> (\d+)x(\d+)", name)
        if m and (int(m.group(1)), int(m.group(2))) in shapes:
            us = timed(9, shapes[(int(m.group(1)), int(m.group(2)))])
        if name.startswith("k_resid_sumsq"):
            us = timed(6, 0)
        m = re.match(r"k_mg_[cp]?tail<program(\d)>", name)
        if m and tail0 >= 0:
            us = timed(7 if m.group(1) in "01" else 8, tail0)
        if us:
            row["isolated_us"] = us
            if row["alg_bytes_per_cell"]:
                row["achieved_gbs"] = row["alg_bytes_per_cell"]*row["cells"]/(us*1e-6)/1e9
    return table


def dram_traffic_table():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
    capture (tools/ncu_dram_table.py writes the file); {} when absent"""
    p = os.path.join(REPO, "profiles", "r02_kernel_dram_bytes.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return {}
    return {}


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
def cpu_arm(n, tracers, steps, warmup, what=None, config="turb"):
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
    if config == "turb":
        f2d = build_case(api, n, tracers, tempfile.mkdtemp())
        cells = n*n
    else:
        f2d = build_config(api, config, tempfile.mkdtemp())
        cells = CONFIGS[config]["nx"]*CONFIGS[config]["ny"]
        what = what or ("%d steps (after %d warm-up) of the same workload, oracle C port with OpenMP on all host "
                        "threads" % (steps, warmup))
    f2d.model.diagnostics(f2d.model.var, 0.)
    log("[cpu] set-up %.1f s on %d threads" % (time.time()-t0, cores))
    for _ in range(warmup):
        loop_body(f2d)
    t0 = time.time()
    for _ in range(steps):
        loop_body(f2d)
    dt = time.time()-t0
    return {"value": cells*steps/dt, "unit": "cell-updates/s", "cores": cores, "kind": "port",
            "sample": what or ("%d steps (after %d warm-up) of the same %dx%d Euler freedecay workload, oracle C port "
                               "with OpenMP on all host threads" % (steps, warmup, n, n)),
            "ms_per_step": 1e3*dt/steps,
            # the reference's own end-of-run figure (core/fluid2d.py:329-338): wall time per
            # iteration per grid point, times the number of cores
            "rescaled_time_core_s_per_cell_update": cores*dt/(steps*cells)}


def resolve_workload(args, world):
    """(n, strong): N = 1 -> 4096^2 (S1); N > 1 -> 16384^2 global, strong scaling (S5) unless --weak"""
    strong = bool(args.strong) or (world > 1 and not args.weak and not args.replicas)
    n = args.n if args.n else (16384 if strong else 4096)
    return n, strong


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    if args.config != "turb":
        steps = max(1, min(args.steps, 40))
        warm = max(1, min(args.warmup, 5))
        r = cpu_arm(0, CONFIGS[args.config]["tracers"], steps, warm, None, args.config)
        line = {"impl": "reference", "metric": "cell_updates_per_s", "value": r["value"], "unit": r["unit"],
                "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": CONFIGS[args.config]["workload"]},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample",
                                                   "rescaled_time_core_s_per_cell_update")},
                "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    n, strong = resolve_workload(args, max(world, args.gpus))
    # bounded sample: the 16384^2 field is a periodic tiling of one 4096^2 freedecay tile, and
    # the host arm runs that tile (per-cell work identical; 86 GB of host arrays are not needed)
    ns = min(n, 4096)
    steps = max(1, min(args.steps, 40))
    warm = max(1, min(args.warmup, 5))
    what = None
    if ns != n:
        what = ("%d steps (after %d warm-up) of one %dx%d tile of the %dx%d workload (the field is that tile "
                "repeated), oracle C port with OpenMP on all host threads" % (steps, warm, ns, ns, n, n))
    r = cpu_arm(ns, args.tracers, steps, warm, what)
    line = {"impl": "reference", "metric": "cell_updates_per_s", "value": r["value"], "unit": r["unit"],
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(n, args.tracers, strong, 1 if strong else max(world, args.gpus))},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample",
                                               "rescaled_time_core_s_per_cell_update")},
            "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def timed_steps(f2d, steps, torch, barrier, lib=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nites = []
    barrier()
    if lib is not None:
        lib.launch_count_reset()
    wall0 = time.time()
    e0.record()
    for _ in range(steps):
        loop_body(f2d)
        nites.append(f2d.model.ope.last_solve[0])
    e1.record()
    barrier()
    return e0.elapsed_time(e1), nites, wall0, time.time()


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
    T = args.tracers
    config = args.config
    n, strong = resolve_workload(args, world)
    t0 = time.time()
    slabs = world > 1 and not args.replicas
    if strong and (args.replicas or n % world):
        raise SystemExit("strong scaling: n must be a multiple of the number of GPUs (and no --replicas)")
    if config != "turb":
        if world > 1:
            raise SystemExit("--config rb / vk are single-GPU lines")
        T = CONFIGS[config]["tracers"]
        n, rows = CONFIGS[config]["nx"], CONFIGS[config]["ny"]
        total_cells = n*rows
        f2d = build_config(api, config, tempfile.mkdtemp())
    else:
        rows = n//world if strong else n          # rows of this rank's slab
        total_cells = n*n if strong else world*n*n
        f2d = build_case(api, n, T, tempfile.mkdtemp(), world if slabs else 1, strong)
    model = f2d.model
    model.diagnostics(model.var, 0.)
    torch.cuda.synchronize()
    mg_slab_levels = getattr(model.ope.gmg, "slab_levels", 0)
    log("[gpu %d] set-up %.1f s" % (rank, time.time()-t0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    for _ in range(args.warmup):
        loop_body(f2d)
    if sampler:
        sampler.wait_ready()
    # ---- timed region: K steps, device timers, inputs (134 MB fields) exceed the 126 MB L2
    ms, nites, wall0, wall1 = timed_steps(f2d, args.steps, torch, barrier, lib)
    launches = int(lib.launch_count())
    clocks = sampler.window(wall0, wall1) if sampler else None
    ms = maxranks(ms)
    ms_step = ms/args.steps
    value = total_cells*args.steps/(ms*1e-3)
    n_F = float(np.mean(nites))

    # ---- per-kernel accounting of the same step without graphs (every rank: the slabs run in lock step)
    peak, peak_src = measured_peaks()
    table, us_nograph = kernel_table(lib, r, f2d, peak, (rows, n))
    dram = dram_traffic_table()
    for row in table:
        row["dram_bytes_per_launch"] = dram.get(row["kernel"].replace(",peer", "").replace(" +stage", ""))

    # ---- the level-0 smoother and the V-cycle timed alone
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
    # the multigrid kernels of the table re-timed alone; shares from launches x best available time
    isolated_times(lib, r, mgh, table, torch)
    for row in table:
        row["us_per_step"] = row["launches_per_step"]*row.get("isolated_us", row["avg_us"])
    tot = sum(row["us_per_step"] for row in table) or 1.
    for row in table:
        row["share"] = row["us_per_step"]/tot
        row["frac_of_peak"] = row["achieved_gbs"]/peak if row["achieved_gbs"] else None
    table.sort(key=lambda row: -row["share"])
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
    smooth_gbs = SMOOTH_BYTES_PER_CELL*rows*n/(smooth_ms*1e-3)/1e9
    # ---- V-cycle (metric part 2): one Vcycle(0) through its CUDA graph
    lib.mg_vcycle(mgh, 0, r.stream)
    torch.cuda.synchronize()
    s0.record()
    for _ in range(10):
        lib.mg_vcycle(mgh, 0, r.stream)
    s1.record()
    torch.cuda.synchronize()
    vcycle_ms = s0.elapsed_time(s1)/10
    del x0, b0

    # ---- end to end: host buffers in, host buffers out, every step
    ds = model.var.dstate
    _ = model.var.state            # materialise the pinned mirror
    for _ in range(2):
        ds._host_written(None)
        loop_body(f2d)
        _ = model.var.state
    barrier()
    h2d0, d2h0 = ds.h2d_bytes, ds.d2h_bytes
    ke = min(args.steps, 5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ke):
        ds._host_written(None)     # the caller's (pinned) host state is the step's input ...
        loop_body(f2d)             # ... uploaded here (H2D of every field) ...
        _ = model.var.state        # ... and the new state is read back (D2H of every field)
    e1.record()
    barrier()
    e2e_ms = maxranks(e0.elapsed_time(e1))
    e2e = {"value": total_cells*ke/(e2e_ms*1e-3), "unit": "cell-updates/s",
           "h2d_bytes_per_step": (ds.h2d_bytes-h2d0)//ke, "d2h_bytes_per_step": (ds.d2h_bytes-d2h0)//ke,
           "ms_per_step": e2e_ms/ke, "steps": ke}

    # ---- N = 1 only: SURVEY.md 8d case S5 on ONE GPU (16384^2, 86 GB), the denominator of the
    # parallel efficiency of the N > 1 lines (which run 16384^2 split in N slabs)
    s5 = None
    if world == 1 and not strong and not args.no_s5 and n == 4096 and config == "turb":
        try:
            del f2d, model, ds
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            t0 = time.time()
            g = build_case(api, 16384, T, tempfile.mkdtemp(), 1, True)
            g.model.diagnostics(g.model.var, 0.)
            for _ in range(3):
                loop_body(g)
            ms5, nit5, _, _ = timed_steps(g, 5, torch, barrier)
            s5 = {"workload": workload_string(16384, T, True), "n_gpus": 1, "steps": 5, "warmup": 3,
                  "ms_per_step": ms5/5, "value": 16384.*16384.*5/(ms5*1e-3), "unit": "cell-updates/s",
                  "n_F_mean": float(np.mean(nit5)), "setup_s": time.time()-t0}
            del g
            gc.collect()
            torch.cuda.empty_cache()
        except Exception as ex:   # a report next to the line, never a reason to lose the line
            s5 = {"failed": repr(ex)}

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
            if config != "turb":
                cpu = cpu_arm(0, T, 5, 1, None, config)
            else:
                cpu = cpu_arm(n, T, 3, 1)
                small = cpu_arm(1024, T, 20, 3)
                cpu["also_1024"] = {k: small[k] for k in ("value", "ms_per_step", "rescaled_time_core_s_per_cell_update")}
        except Exception as ex:   # the baseline is a report, never a reason to lose the GPU line
            cpu = {"value": None, "unit": "cell-updates/s", "cores": 0, "kind": "port", "sample": "failed: %r" % ex}
    balg = b_alg_config(config, T, n_F)
    # dominant kernel = the largest share of the step in the accounting above, among the
    # bandwidth-bound kernels (the latency-bound tail is listed in the table without a roofline)
    dom = next((row for row in table if row["achieved_gbs"]), None)
    # DRAM bytes of one whole step as ncu counted them (every kernel of an ungraphed 4096^2 step,
    # tools/ncu_step_dram.py -> profiles/r02_step_dram_bytes.json); only for the case it was taken on
    step_dram = None
    if config == "turb" and world == 1 and n == 4096 and T == 1:
        try:
            step_dram = json.load(open(os.path.join(REPO, "profiles", "r02_step_dram_bytes.json")))["step_total_bytes"]
        except Exception:
            step_dram = None
    line = {
        "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": CONFIGS[config]["workload"] if config != "turb" else workload_string(n, T, strong, world),
                   "grid": [n, rows*world if strong or world == 1 else rows], "tracers": T, "n_F_mean": n_F,
                   "parallelism": "single GPU" if world == 1 else (
                       "%d y-slabs of %dx%d (global %dx%d), halo rows by peer stores over NVLink fused into the "
                       "producing kernels, coarse levels gathered" % (world, n, rows, n, rows*world)
                       if slabs else "%d independent replicas" % world),
                   "mg_slab_levels": mg_slab_levels,
                   "cache": "working set %.1f GB >> 126 MB L2 (no flush needed)" % (40*(n+6)*(rows+6)*8/1e9)},
        "roofline": {"bound": "hbm",
                     "kernel": dom["kernel"] if dom else None,
                     "share_of_step": dom["share"] if dom else None,
                     "achieved": dom["achieved_gbs"] if dom else None, "peak": peak, "unit": "GB/s",
                     "frac": dom["frac_of_peak"] if dom else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture
                     # committed under profiles/ (r02_kernel_dram_bytes.json); null without a capture
                     "traffic": dom["dram_bytes_per_launch"] if dom else None,
                     "peak_source": peak_src,
                     "ms_per_launch": dom.get("isolated_us", dom["avg_us"])*1e-3 if dom else None,
                     "algorithmic_bytes_per_cell": dom["alg_bytes_per_cell"] if dom else None,
                     "how": "launch counts from CUDA events around every launch of 2 steps run without CUDA "
                            "graphs (f2d_prof_begin/report); the multigrid kernels then timed alone, 10 launches "
                            "back to back between two CUDA events (f2d_mg_bench_op); dominant = largest "
                            "launches x time; achieved = algorithmic bytes per launch / that time",
                     "level0_smoother_alone": {"kernel": "k_smooth2<mode1,input0> (Grid.smooth, level 0)",
                                               "ms_per_launch": smooth_ms, "achieved": smooth_gbs,
                                               "frac": smooth_gbs/peak, "algorithmic_bytes_per_cell": SMOOTH_BYTES_PER_CELL}},
        "kernels": [row for row in table if row["share"] >= 0.004],
        "step_us_without_graphs": us_nograph,
        "step_hbm": {"b_alg_bytes_per_cell": balg, "achieved_gbs": balg*value/world/1e9,
                     "frac_of_peak": balg*value/world/1e9/peak,
                     "note": "contract bytes (SURVEY.md 8d): fused kernels and the field-skipping RK maps move "
                             "less than this; step_dram_bytes is what ncu saw",
                     "step_dram_bytes": step_dram},
        "vcycle_ms": vcycle_ms,
        "vcycle_frac_of_peak": VCYCLE_BYTES_PER_CELL*rows*n/(vcycle_ms*1e-3)/1e9/peak,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if s5 is not None:
        line["s5_one_gpu"] = s5
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=0, help="grid size (default: 4096 on one GPU, 16384 global on several)")
    ap.add_argument("--tracers", type=int, default=1)
    ap.add_argument("--config", default="turb", choices=["turb", "rb", "vk"],
                    help="turb: Twodim_turbulence (the headline); rb / vk: RayleighBenard 2048x1024, VonKarman 4096x1024")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-s5", action="store_true", help="N=1: skip the 16384^2 single-GPU run")
    ap.add_argument("--replicas", action="store_true", help="N>1: independent replicas instead of slabs")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: --n is the GLOBAL grid (default 16384), split in --gpus y-slabs "
                         "(the default when N > 1)")
    ap.add_argument("--weak", action="store_true", help="N>1: weak scaling, --n (4096) squared cells per GPU")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        reference_main(args)
    else:
        gpu_main(args)


if __name__ == "__main__":
    main()
