"""Worker of tests/test_contraction_headroom.py: the product's host layer on the emulated C ABI
whose kernels come from a variant of oracle/f2d_oracle.c compiled WITH FMA contraction
(F2D_ORACLE_LIB); prints, per case and step count, the worst relative L2 distance of a field
from the fixture."""
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

out = {}
for name in sys.argv[1:]:
    api, emu = emu_device.install()
    gold = np.load(os.path.join(HERE, "golden", name+".npz"))
    with contextlib.redirect_stdout(io.StringIO()):
        f2d = cases.CASES[name](api, tempfile.mkdtemp())
        res = cases.run_steps(f2d)
    names = list(f2d.model.var.varname_list)
    for nstep, (state, t, dt, diags) in sorted(res.items()):
        g = gold["state%i" % nstep]
        out["%s:%d" % (name, nstep)] = max(
            float(np.linalg.norm(state[k]-g[k])/max(np.linalg.norm(g[k]), 1e-300)) for k in range(len(names)))
    emu_device.uninstall()
print(json.dumps(out))
