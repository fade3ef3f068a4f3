"""Step time of the other BASELINE.json configurations at their full sizes (not bench lines:
`bench.py` carries the headline; these are the masked-kernel paths):
  RayleighBenard  Boussinesq 2048x1024, xchannel, no-slip, diffusion, forcing
  VonKarman       Euler 4096x1024, xchannel, disc obstacle + islands, no-slip, diffusion, sponge
    python tools/bench_configs.py [steps]
"""
import json
import os
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
import torch  # noqa: E402
import fluid2d_b200  # noqa: E402
import cases  # noqa: E402
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
api = fluid2d_b200.api()
so = sys.stdout
sys.stdout = sys.stderr
out = []
for name, build, cells in (("RayleighBenard 2048x1024", lambda d: cases.rb(api, d, 2048), 2048*1024),
                           ("VonKarman 4096x1024", lambda d: cases.karman(api, d, 1024, ratio=4), 4096*1024)):
    f2d = build(tempfile.mkdtemp())
    f2d.model.diagnostics(f2d.model.var, 0.)
    for _ in range(5):
        bench.loop_body(f2d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        bench.loop_body(f2d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/steps
    gmg = f2d.model.ope.gmg
    out.append({"config": name, "ms_per_step": ms, "cell_updates_per_s": cells/(ms*1e-3),
                "maxspeed": float(f2d.model.diags["maxspeed"]), "ke": float(f2d.model.diags["ke"]),
                "matrix_modes": [g.matrix_mode for g in gmg.grid]})
    del f2d
    torch.cuda.empty_cache()
sys.stdout = so
for o in out:
    print(json.dumps(o))
