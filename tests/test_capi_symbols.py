"""The C-ABI library loads and exports every symbol include/fvmgpu.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from fvm_b200 import build, capi


def header_symbols():
    text = open(os.path.join(ROOT, "include", "fvmgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fvmgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) >= 40
    assert sorted(capi.SIGNATURES) == syms


def test_product_library_exports_every_symbol():
    path = build.build_lib()  # nvcc cross-compiles sm_100a without a GPU
    dll = ctypes.CDLL(path)
    for s in header_symbols():
        assert hasattr(dll, s), s
    assert dll.fvmgpu_version() == 100


def test_library_holds_sm100a_code_only():
    import subprocess
    path = build.build_lib()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product must refuse loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = capi.Lib(build.build_lib())
    with pytest.raises(capi.FvmGpuError, match="CUDA device"):
        lib.init(0)
    with pytest.raises(capi.FvmGpuError, match="not initialised"):
        lib.call("fvmgpu_flush_l2")


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(capi.FvmGpuError, match="no CPU fallback"):
        capi.Lib(str(tmp_path / "libfvmgpu.so"))


def test_product_package_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fvm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)
                assert "hostsim.so" not in src or f == "build.py", f
