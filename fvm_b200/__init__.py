"""fvm_b200 -- B200-native assembly + AMG/BCGStab hot path for MEMOSA-FVM (btanasoi/fvm).

Layout: csrc/ (CUDA kernels + C ABI, built into libfvmgpu.so), capi.py (ctypes binding),
models.py (host-side mirror of the reference's Python model API), meshgen.py (synthetic meshes and
host mesh metrics), build.py (nvcc build). The CUDA library is mandatory: nothing here computes on
the CPU.
"""
from .capi import FvmGpuError  # noqa: F401

__all__ = ["capi", "models", "meshgen", "build", "FvmGpuError"]
