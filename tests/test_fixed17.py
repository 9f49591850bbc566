"""CPU test of the dump writer's number formatting (pem_spgemm_b200/csrc/fixed17.h): compiled with the host compiler
and compared with std::to_chars(fixed, 17) — the digits the reference's `std::fixed << std::setprecision(17)` prints
(/root/reference/spgemm.cu:1529) — on several million probes."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fixed17_matches_to_chars(tmp_path):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not cxx:
        pytest.skip("no host C++ compiler")
    exe = str(tmp_path / "fixed17_check")
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-I", os.path.join(ROOT, "pem_spgemm_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "fixed17_check.cpp"), "-o", exe])
    out = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "mismatches 0" in out.stdout
