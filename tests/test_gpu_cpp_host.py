"""Runs the C++ host-mirror test driver (tests/cpp/test_reference_kats.cpp): the reference's
own product tests written against include/sprs_b200.hpp, the compiled-language stand-in
for the Rust wrapper crate."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_reference_kats():
    exe = os.path.join(ROOT, "tests", "cpp", "test_reference_kats")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe)])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK ")
