"""bench.py's reporting logic on the GPU-less box: the kernels' host simulator stands in for the library
(monkeypatched; the real bench loads fvm_b200/libfvmgpu.so and needs a B200), a tiny mesh for the
workload. Checks that the one JSON line carries every key of the measurement contract."""
import io
import json
import types
from contextlib import redirect_stdout

import pytest


@pytest.mark.parametrize("mesh", ["hex", "tet"])
def test_bench_json_line_has_the_contract_keys(hostsim_lib, monkeypatch, mesh):
    import bench
    from fvm_b200 import capi
    monkeypatch.setattr(capi, "default_lib", lambda: hostsim_lib)
    monkeypatch.setattr(bench, "MESH", mesh)
    monkeypatch.setattr(bench, "KRYLOV", False)
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, n=6, no_profile=False, no_cpu_baseline=True, ref_n=0,
                                 parity_size=8)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_ours(args)
    line = [l for l in buf.getvalue().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "step_ms", "solve_split_ms",
              "solve_hbm", "amg_cycles"):
        assert k in d, k
    assert d["metric"] == "fp64_cell_updates_per_s" and d["unit"] == "cell-updates/s" and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert ("hex" in d["config"]["workload"]) == (mesh == "hex") and "%" not in d["config"]["workload"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0 and d["amg_cycles"] > 0 and len(d["step_ms"]) == 2
    assert 0 < d["solve_hbm"]["bytes_per_cycle"]
    # parity block: this run's own small problem against the oracle (assembly 1e-12, solution 1e-8)
    oc = d["parity"]["oracle_check"]
    assert oc["pass"] and oc["assembly_max_rel_diff"]["diag"] <= 1e-12 and oc["solution_rel_l2"] <= 1e-8, oc
    if mesh == "hex":
        assert d["parity"]["workload_vs_exact_solution_rel_l2"] < 1e-6


@pytest.mark.parametrize("workload", ["cavity", "electric-tet"])
def test_secondary_workloads_print_the_same_contract(hostsim_lib, monkeypatch, workload):
    """BASELINE configs[2] (cavity, FlowModel SIMPLE) and configs[4] (ElectricModel on tets) through bench.py's workload
    runners at toy size in the host simulator: contract keys, cpu_baseline from the reference's own model, and the
    parity block green against oracle/_ref."""
    import bench
    import bench_workloads as W
    from fvm_b200 import capi
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref not built")
    monkeypatch.setattr(capi, "default_lib", lambda: hostsim_lib)
    monkeypatch.setattr(bench, "WORKLOAD", workload)
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, n=20 if workload == "cavity" else 5, no_profile=False,
                                 no_cpu_baseline=False, ref_n=12 if workload == "cavity" else 4,
                                 parity_size=12 if workload == "cavity" else 4)
    buf = io.StringIO()
    with redirect_stdout(buf):
        (W.run_cavity if workload == "cavity" else W.run_electric)(args)
    d = json.loads([l for l in buf.getvalue().splitlines() if l.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "parity", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "fp64_cell_updates_per_s" and d["dtype"] == "f64" and d["gpu_launches"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 0
    assert d["parity"]["oracle_check"]["pass"], d["parity"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
