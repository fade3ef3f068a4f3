"""The committed fixtures (reference Python on the C restatement of the kernels) against the
reference's Python on the reference's own Fortran SOURCE, executed through
oracle/fortran_source.py: initial state (a full multigrid solve), multigrid masks and
matrices, the state after one step, dt and every diagnostic -- bit for bit.  Run in a
subprocess (the reference's flat module names must not meet the product's in one
interpreter).  Build container only; one small case here (pure-Python loops are slow), the
others in profiles/r01_fixtures_on_fortran_source.txt."""
import os
import subprocess
import sys

import pytest

from oracle import fortran_source as F

pytestmark = pytest.mark.skipif(not F.available(), reason="/root/reference is not present on this machine")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", ["freedecay_32_o3_notracer"])
def test_fixture_is_what_the_reference_computes_on_its_own_fortran(name):
    p = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_golden.py"), "--kernels",
                        "fortran_source", "--check", "--steps", "1", name],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:]
    assert "IDENTICAL" in p.stdout
