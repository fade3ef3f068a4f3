"""The reference's experiment scripts, run as they are (tools/experiment_compat.py): once on
the reference's own Python over the oracle kernels, once on the product's host layer over the
emulated C ABI; same final state, clock and diagnostics, bit for bit.  Three of the BASELINE
scripts here (two loop iterations); the table of all 33 is profiles/r01_experiment_compat.txt.
Build container only (the scripts live in /root/reference/experiments)."""
import os
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
import experiment_compat as ec  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.isdir(ec.EXPERIMENTS), reason="/root/reference is not present")


@pytest.mark.parametrize("script", ["Vortex/vortex.py", "RayleighBenard/rayleigh_benard.py",
                                    "VonKarman/karman_street.py"])
def test_reference_script_runs_unchanged_and_agrees(script):
    path = os.path.join(ec.EXPERIMENTS, script)
    tmp = tempfile.mkdtemp()
    os.environ.setdefault("OMP_NUM_THREADS", "2")
    outs = {impl: os.path.join(tmp, impl+".npz") for impl in ("reference", "product")}
    with ThreadPoolExecutor(2) as ex:
        res = list(ex.map(lambda impl: ec.run_one(impl, path, outs[impl], 2, 600.), outs))
    assert [r[0] for r in res] == ["ok", "ok"], res
    verdict, _worst = ec.compare(outs["reference"], outs["product"])
    assert verdict == "bit-identical", verdict
