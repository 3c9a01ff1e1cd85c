"""Drop-in check on the reference's side of the boundary: `integration/fvm_gpu_adaptor.h` compiled
against the UNMODIFIED reference headers/objects (oracle/Makefile `adaptor` target ->
oracle/_ref/adaptor_test, built in the dev container where /root/reference is mounted; the binary
travels to the GPU box). The reference's own ThermalModel<double>::advance runs with
`options.linearSolver = GpuAMG / GpuBCGStab` and is compared with the reference AMG (1e-8 rel L2);
`GpuScalarLinearizer` is compared entry by entry with the reference's linearize + initSolve (1e-12)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "adaptor_test")


@pytest.mark.gpu
@pytest.mark.parametrize("n", [24, 57])
def test_reference_thermal_model_with_gpu_solver_and_linearizer(n):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/adaptor_test not built (needs /root/reference at build time)")
    r = subprocess.run([BIN, str(n)], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "adaptor_test: OK" in r.stdout


def test_adaptor_glue_against_the_host_simulator():
    """The same binary linked with the TEST-ONLY host simulator of the library (oracle/Makefile adaptor_hostsim): the
    reference-side glue -- GpuAMG / GpuBCGStab as LinearSolvers of the reference's ThermalModel, GpuScalarLinearizer,
    the device-resident outer iteration and GpuFlowModel driven by the reference's own FlowModel object -- runs on the
    GPU-less box against the reference classes compiled in place."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not mounted")
    from fvm_b200 import build
    build.build_hostsim()
    mk = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "adaptor_hostsim"], capture_output=True, text=True)
    assert mk.returncode == 0, mk.stdout + mk.stderr
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "adaptor_test_hostsim"), "16"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "adaptor_test: OK" in r.stdout, r.stdout + r.stderr
    assert r.stdout.count(" OK") >= 6
