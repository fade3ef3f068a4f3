"""Worker of tests/test_slab_emulated.py: two CPU processes (gloo), the product's host layer on
the emulated C ABI with the y-slab communicator (tests/emu_device.py).  Builds freedecay with
npy = world size, advances it, gathers the global fields and compares them with the same
global problem run on ONE rank in the same process (the CPU twin of tests/slab_worker.py).

    torchrun --nproc-per-node 2 tests/slab_emu_worker.py out.json nx ny nsteps
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import emu_device  # noqa: E402
import cases  # noqa: E402


from slab_cases import BUILDERS  # noqa: E402


def glue(emu, stack):
    """[k][local rows][nx] on every rank -> [k][global rows][nx]"""
    parts = [np.stack(emu.lib._gather(stack[k])) for k in range(stack.shape[0])]
    return np.stack([np.concatenate([p[0][:3]]+[q[3:-3] for q in p]+[p[-1][-3:]], axis=0) for p in parts])


def history_per_rank(api, out, nx, ny, nsteps, world):
    """Fluid2d.loop() on slabs: EVERY rank writes its own history file, record by record, while
    the loop runs (output.py:24-31,77-95); only the diagnostics file belongs to rank 0"""
    import torch.distributed as dist
    import output
    with contextlib.redirect_stdout(io.StringIO()):
        f2d = BUILDERS["freedecay"](api, tempfile.mkdtemp(), nx, ny, world)
        o = f2d.output
        o.freq_his = o.freq_diag = 0.
        o.tnexthis = o.tnextdiag = 0.
        f2d.exacthistime = False
        f2d.loop(nsteps=nsteps, joinhis=False)        # not even the end-of-run dump: the records are on disk already
    rank = dist.get_rank()
    his = output.load_records(o.hisfile)
    state = np.array(f2d.model.var.state, copy=True)
    k = f2d.model.var.index("vorticity")
    report = {"rank": rank, "hisfile": os.path.basename(o.hisfile), "exists": os.path.exists(o.hisfile),
              "nrec": int(len(his["t"])), "shape": list(his["vorticity"].shape),
              "last_is_my_slab": bool(np.array_equal(his["vorticity"][-1], state[k][3:-3, 3:-3].astype(np.float32))),
              "diag_exists": os.path.exists(o.diagfile)}
    everyone = [None]*world
    dist.all_gather_object(everyone, report)
    if rank == 0:
        json.dump(everyone, open(out, "w"))
    dist.barrier()
    dist.destroy_process_group()


def main():
    out, nx, ny, nsteps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    case = sys.argv[5] if len(sys.argv) > 5 else "freedecay"
    world = int(os.environ["WORLD_SIZE"])
    api, emu = emu_device.install()
    if case == "freedecay_his":
        return history_per_rank(api, out, nx, ny, nsteps, world)
    build = BUILDERS[case]
    with contextlib.redirect_stdout(io.StringIO()):
        f2d = build(api, tempfile.mkdtemp(), nx, ny, world)
        res = cases.run_steps(f2d, (nsteps,))[nsteps]
        flx = cases.run_fluxes(f2d) if getattr(f2d, "diag_fluxes", False) else None
    import torch.distributed as dist
    rank = dist.get_rank()
    glob = glue(emu, np.array(res[0]))
    gflx = glue(emu, np.array(flx)) if flx is not None else None
    report = {"rank": rank, "kt": f2d.kt, "t": res[1], "dt": res[2], "diags": res[3],
              "solve": list(f2d.model.ope.last_solve), "slab_levels": f2d.model.ope.gmg.slab_levels}
    if rank == 0:
        # the same global problem on one rank (a fresh single-rank emulator; no process group use)
        emu_device.uninstall()
        api1, emu1 = emu_device.install()
        with contextlib.redirect_stdout(io.StringIO()):
            one = build(api1, tempfile.mkdtemp(), nx, ny, 1)
            r1 = cases.run_steps(one, (nsteps,))[nsteps]
            flx1 = cases.run_fluxes(one) if flx is not None else None
        ref = np.array(r1[0])
        names = list(one.model.var.varname_list)
        report["fields_equal"] = {nm: bool(np.array_equal(glob[k][3:-3], ref[k][3:-3])) for k, nm in enumerate(names)}
        report["maxdiff"] = {nm: float(np.abs(glob[k][3:-3]-ref[k][3:-3]).max()) for k, nm in enumerate(names)}
        if flx is not None:
            report["fluxes_equal"] = bool(np.array_equal(gflx[:, 3:-3], np.array(flx1)[:, 3:-3]))
            report["fluxes_maxdiff"] = float(np.abs(gflx[:, 3:-3]-np.array(flx1)[:, 3:-3]).max())
        report["one"] = {"t": r1[1], "dt": r1[2], "diags": r1[3], "solve": list(one.model.ope.last_solve)}
    everyone = [None]*world
    dist.all_gather_object(everyone, report)
    if rank == 0:
        json.dump(everyone, open(out, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
