"""CPU suite: the pinned staging ring's copying threads (csrc/hr_staging.h) on pitched pictures — the lattice-rows-first
upload of pageable planes — without a GPU: tests/staging_harness.cpp stands a memcpy in for the copy engine."""
import pathlib
import shutil
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_copy_crew_gathers_and_scatters_pitched_blocks(tmp_path):
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("g++ not present")
    exe = tmp_path / "staging_harness"
    subprocess.run([cxx, "-O2", "-std=c++17", "-pthread", "-I", str(ROOT / "mpv-frame-interpolator_b200" / "csrc"), "-o", str(exe),
                    str(ROOT / "tests" / "staging_harness.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr
