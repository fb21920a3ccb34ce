"""The offline kernel tools (tools/) keep working: static SASS cost model, SASS comparison, ncu phase budget.
They need the CUDA toolkit's cuobjdump / nvdisasm / ncu but no GPU."""
import os
import shutil
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("nvdisasm") is None,
                                reason="needs the CUDA toolkit binaries")


def _run(*args):
    return subprocess.run([sys.executable, *args], cwd=ROOT, capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(shutil.which("ncu") is None, reason="needs ncu")
def test_ncu_regions_on_a_saved_report():
    rep = os.path.join(ROOT, "gpurun_out", "prof_k1f_final.ncu-rep")
    if not os.path.exists(rep):
        pytest.skip("no saved ncu report in gpurun_out/ (scratch, not committed)")
    res = _run("tools/ncu_regions.py", rep)
    assert res.returncode == 0, res.stderr
    assert "MAIN LOOP" in res.stdout and "per-read metadata" in res.stdout
