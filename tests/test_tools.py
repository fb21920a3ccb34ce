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


def test_sass_regions_attributes_the_default_kernel():
    res = _run("tools/sass_regions.py", "pileup_tiled_kernelILb1ELb0")
    assert res.returncode == 0, res.stderr
    out = res.stdout
    assert "SASS instructions:" in out and "MAIN LOOP" in out and "flush_window" in out


def test_sass_same_accepts_identical_builds(tmp_path):
    dump = tmp_path / "a.sass"
    lib = os.path.join(ROOT, "kindel_b200", "_lib", "libkindel_b200.so")
    dump.write_text(subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout)
    res = _run("tools/sass_same.py", str(dump), str(dump))
    assert res.returncode == 0 and "DIFFERENT" not in res.stdout and res.stdout.count("same") >= 20


@pytest.mark.skipif(shutil.which("ncu") is None, reason="needs ncu")
def test_ncu_regions_on_a_saved_report():
    rep = os.path.join(ROOT, "gpurun_out", "prof_k1f_final.ncu-rep")
    if not os.path.exists(rep):
        pytest.skip("no saved ncu report in gpurun_out/ (scratch, not committed)")
    res = _run("tools/ncu_regions.py", rep)
    assert res.returncode == 0, res.stderr
    assert "MAIN LOOP" in res.stdout and "per-read metadata" in res.stdout
