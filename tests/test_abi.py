"""The C-ABI library loads and exports every symbol include/crane_sched.h
declares (no compute without a GPU), and fails loudly without a device."""
import ctypes as C
import os
import re

import pytest

from cranesched_b200 import abi
from cranesched_b200.scheduler import EXPORTS, LIB_PATH, CraneSchedError, GpuScheduler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "crane_sched.h")).read()
    return sorted(set(re.findall(r"\b(crane_sched_[a-z_]+)\s*\(", src)))


def test_header_and_wrapper_agree():
    assert _declared() == sorted(EXPORTS)


def test_product_library_exports_all_symbols(gpu_lib):
    lib = C.CDLL(gpu_lib)
    for name in _declared():
        assert hasattr(lib, name), name


def test_struct_sizes_match_header():
    assert abi.RES_IN_NODE.itemsize == 72
    assert abi.RES_VIEW.itemsize == 56
    assert C.sizeof(abi.SchedConfig) == 56
    assert C.sizeof(abi.TimingC) == 36  # + qos_ms


def test_no_cpu_fallback_without_device(gpu_lib):
    """Without a CUDA device the product path must refuse to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(CraneSchedError) as e:
        GpuScheduler(abi.Config(), 0, gpu_lib)
    assert e.value.code == abi.ENODEV


def test_product_does_not_reference_oracle():
    """Nothing under cranesched_b200/ may import or link oracle/."""
    pkg = os.path.join(ROOT, "cranesched_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "crane_oracle" not in txt and "pyoracle" not in txt and "from oracle" not in txt, f
