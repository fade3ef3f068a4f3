"""Worker of the y-slab tests: launched by torchrun, one process per GPU.  Builds the
freedecay case with npy = world size, advances it, gathers the global fields on rank 0
and compares them with a single-GPU run of the same global problem (ref .npz).

    torchrun --nproc-per-node 2 tests/slab_worker.py ref.npz out.json nx ny nsteps [case]
"""
import json
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import fluid2d_b200  # noqa: E402
import cases  # noqa: E402


def gather_global(state, nh, world):
    """[nvar, nyl, nx] local slabs -> [nvar, ny+2nh, nx] global array (rank order = south to north)"""
    t = torch.from_numpy(np.ascontiguousarray(state)).cuda()
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    parts = [p.cpu().numpy() for p in parts]
    rows = [parts[0][:, :nh, :]] + [p[:, nh:-nh, :] for p in parts] + [parts[-1][:, -nh:, :]]
    return np.concatenate(rows, axis=1)


def main():
    refpath, outpath, nx, ny, nsteps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    api = fluid2d_b200.api()
    world = int(os.environ["WORLD_SIZE"])
    so = sys.stdout
    sys.stdout = sys.stderr
    case = sys.argv[6] if len(sys.argv) > 6 else "freedecay"
    if case == "freedecay":
        f2d = cases.freedecay(api, tempfile.mkdtemp(), nx, ny=ny, npy=world)
    else:
        sys.path.insert(0, os.path.join(REPO, "tests"))
        from slab_cases import BUILDERS
        f2d = BUILDERS[case](api, tempfile.mkdtemp(), nx, ny, world)
    model = f2d.model
    rank = dist.get_rank()
    ref = np.load(refpath)
    names = list(model.var.varname_list)
    report = {"slab_levels": model.ope.gmg.slab_levels, "nlevs": model.ope.gmg.nlevs, "errors": {}}

    def compare(tag, gstate):
        if "oracle_"+tag in ref.files:      # the CPU oracle's run of the same global problem
            g = ref["oracle_"+tag]
            for k, nm in enumerate(names):
                n = np.linalg.norm(g[k][3:-3])
                e = np.linalg.norm(gstate[k][3:-3]-g[k][3:-3])/(n if n > 0 else 1.)
                report.setdefault("oracle_errors", {})["%s:%s" % (tag, nm)] = float(e)
        g = ref[tag]
        for k, nm in enumerate(names):
            n = np.linalg.norm(g[k][3:-3])
            e = np.linalg.norm(gstate[k][3:-3]-g[k][3:-3])/(n if n > 0 else 1.)
            report["errors"]["%s:%s" % (tag, nm)] = float(e)
            if e > 1e-10 and tag == "state0":
                d = np.abs(gstate[k]-g[k])
                report.setdefault("detail", {})[nm] = {
                    "rowmax": [float(x) for x in d.max(axis=1)[::8]], "mean_diff": float((gstate[k]-g[k])[3:-3, 3:-3].mean()),
                    "ref_absmax": float(np.abs(g[k]).max())}

    report["solve0"] = list(model.ope.last_solve)
    report["modes"] = [g.matrix_mode for g in model.ope.gmg.grid]
    from runtime import rt
    ep = torch.tensor([rt().lib.comm_epoch(rt().comm)], dtype=torch.int64, device="cuda")
    eps = [torch.zeros_like(ep) for _ in range(world)]
    dist.all_gather(eps, ep)
    report["epochs_after_setup"] = [int(e.item()) for e in eps]
    g0 = gather_global(np.array(model.var.state), 3, world)
    if rank == 0:
        compare("state0", g0)
    res = cases.run_steps(f2d, (1, nsteps))
    for k in (1, nsteps):
        gk = gather_global(res[k][0], 3, world)
        if rank == 0:
            compare("state%d" % k, gk)
            report["dt%d" % k] = res[k][2]
            report["dt%d_ref" % k] = float(ref["dt%d" % k])
    sys.stdout = so
    if rank == 0:
        json.dump(report, open(outpath, "w"))
    dist.barrier()


if __name__ == "__main__":
    main()
