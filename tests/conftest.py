import contextlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def hostsim_lib():
    """TEST-ONLY single-threaded simulator of the kernel functors (csrc/common.cuh, FVMGPU_HOSTSIM).
    It exercises the host logic and the index logic of the kernels on the GPU-less box; it is not the
    product and is never loaded by fvm_b200."""
    from fvm_b200 import build, capi
    lib = capi.Lib(build.build_hostsim())
    lib.init(0)
    return lib


@pytest.fixture(scope="session")
def gpu_lib():
    """The product: fvm_b200/libfvmgpu.so on cuda:0 (through the C ABI)."""
    from fvm_b200 import capi
    return capi.default_lib()


def _lib_params():
    return [pytest.param("gpu", marks=pytest.mark.gpu), pytest.param("hostsim")]


@pytest.fixture(params=_lib_params())
def devlib(request):
    if request.param == "gpu":
        return request.getfixturevalue("gpu_lib")
    return request.getfixturevalue("hostsim_lib")


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref: the reference's own C++ hot path (travels to the GPU box prebuilt)."""
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref/libfvmref.so not built")
    return refapi


def oracle_aggregator():
    """Address of the oracle's restatement of the reference's sequential agglomeration (oracle/fvm_oracle.c:
    fvmo_create_coarsening, signature fvmgpu_aggregate_fn) for the library's verification hook."""
    import ctypes as C
    from oracle import port
    return C.cast(port.lib().fvmo_create_coarsening, C.c_void_p).value


@contextlib.contextmanager
def reference_order_mode(lib):
    """Reference-order verification mode of `lib` (include/fvmgpu.h: fvmgpu_debug_set_aggregator): aggregates from the
    oracle's sequential agglomeration, smoothing in natural row order."""
    lib.set_aggregator(oracle_aggregator())
    try:
        yield
    finally:
        lib.set_aggregator(None)


@pytest.fixture
def reference_order(hostsim_lib):
    with reference_order_mode(hostsim_lib):
        yield


@pytest.fixture
def reference_order_dev(devlib):
    """the same for tests parametrised over the device library and the host simulator"""
    with reference_order_mode(devlib):
        yield


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))
