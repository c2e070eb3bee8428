"""The NodeSelect-shaped C++ adaptor (cranesched_b200/adaptor) driven with
reference-style objects (hostnames, gres names, device paths) must reproduce the
oracle fed with the same scenario as raw C-ABI tables."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run(lib_path, exe):
    libdir, libname = os.path.split(lib_path)
    out = os.path.join(ROOT, "tests", "_emu", exe)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["g++", "-O1", "-g", "-std=c++17", os.path.join(ROOT, "tests", "adaptor", "test_adaptor.cpp"),
           os.path.join(ROOT, "cranesched_b200", "adaptor", "crane_adaptor.cpp"),
           "-L" + libdir, "-l:" + libname, "-L" + os.path.join(ROOT, "oracle"), "-l:libcrane_oracle.so",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-pthread", "-o", out]
    subprocess.check_call(cmd)
    return subprocess.run([out], capture_output=True, text=True, timeout=600)


def test_adaptor_on_emulated_kernels(oracle, emu_lib):
    r = _build_and_run(emu_lib, "test_adaptor_emu")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout and "adaptor qos:" in r.stdout


@pytest.mark.gpu
def test_adaptor_on_gpu(oracle, gpu_lib):
    r = _build_and_run(gpu_lib, "test_adaptor_gpu")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout and "adaptor qos:" in r.stdout
