"""Host-layer orchestration is frozen: for every scenario of tests/golden/cases.py (the seven
experiment set-ups, one run per time scheme, the two diag_fluxes cases) and for a short
Fluid2d.loop() with output, the sequence of C-ABI calls the Python host layer makes -- entry
point, which field of which state buffer every pointer refers to, every scalar argument --
must equal the trace recorded from the GPU-verified host layer
(tests/golden/host_traces.json.gz, regenerate with `python tests/test_host_trace.py --write`
only together with a green `pytest -m gpu` run).  Runs on the CPU through tests/mock_device.py;
no arithmetic is involved."""
import gzip
import io
import json
import os
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
FIXTURE = os.path.join(HERE, "golden", "host_traces.json.gz")

import cases  # noqa: E402


def advection_case(api, datadir):
    """the `advection` model (core/advection.py): a tracer in the flow of a prescribed psi"""
    import numpy as np
    param = api.Param('default.xml')
    param.modelname = 'advection'
    cases._common(param, 'adv_32', datadir)
    param.nx = param.ny = 32
    param.geometry = 'closed'
    param.order = 5
    param.timestepping = 'RK3_SSP'
    param.diffusion = True
    param.Kdiff = 1e-3
    param.var_to_save = ['tracer']
    grid = api.Grid(param)
    f2d = api.Fluid2d(param, grid)
    model = f2d.model
    trac = model.var.get('tracer')
    trac[:] = np.exp(-((grid.xr-0.5)**2+(grid.yr-0.5)**2)/0.05)*grid.msk
    model.set_psi_from_tracer()
    trac[:] = np.round(grid.xr*4) % 2
    return f2d


EXTRA = {"advection_32": advection_case}


def scenarios():
    names = sorted(cases.CASES)
    return names + sorted(EXTRA) + ["loop:freedecay_32_flx", "loop:rb_64"]


def record(name):
    import mock_device
    api, fake = mock_device.install()
    so = sys.stdout
    sys.stdout = io.StringIO()
    try:
        loop = name.startswith("loop:")
        case = name.split(":")[-1]
        f2d = (cases.CASES.get(case) or EXTRA[case])(api, tempfile.mkdtemp())
        if loop:
            f2d.output.freq_his = f2d.output.freq_diag = 0.
            f2d.output.tnexthis = f2d.output.tnextdiag = 0.
            f2d.exacthistime = False
            f2d.loop(nsteps=2)
        else:
            if getattr(f2d, "diag_fluxes", False):
                cases.run_fluxes(f2d)
            cases.run_steps(f2d, (3,))
            if getattr(f2d, "diag_fluxes", False):
                cases.run_fluxes(f2d)
    finally:
        sys.stdout = so
        mock_device.uninstall()
    # (mg_destroy comes from Gmg.__del__, i.e. whenever the garbage collector runs)
    calls = [c for c in list(fake.lib.calls) if c[0] != "mg_destroy"]
    # the integral diagnostics the host derived from the (mock) reductions, and the clock
    diags = sorted((k, float("%.12g" % float(np.ravel(v)[0]))) for k, v in f2d.model.diags.items())
    calls.append(["<diags>", diags, float("%.12g" % f2d.t), float("%.12g" % f2d.dt), f2d.kt])
    return calls


def load():
    with gzip.open(FIXTURE, "rt") as f:
        return json.load(f)


@pytest.mark.parametrize("name", scenarios())
def test_host_layer_sends_the_frozen_call_sequence(name):
    ref = load()[name]
    got = json.loads(json.dumps(record(name)))
    assert len(got) == len(ref), "%s: %d calls, the frozen trace has %d" % (name, len(got), len(ref))
    for k, (a, b) in enumerate(zip(got, ref)):
        assert repr(a) == repr(b), "%s: call #%d differs:\n  now    %r\n  frozen %r" % (name, k, a, b)


if __name__ == "__main__":
    if "--write" in sys.argv:
        out = {name: record(name) for name in scenarios()}
        with gzip.open(FIXTURE, "wt") as f:
            json.dump(out, f)
        print("wrote %s: %d scenarios, %d calls, %d bytes" % (
            FIXTURE, len(out), sum(len(v) for v in out.values()), os.path.getsize(FIXTURE)))
