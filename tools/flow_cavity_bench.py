"""Lid-driven cavity (BASELINE.json configs[2]: FlowModel, SIMPLE momentum + pressure-correction AMG
on a synthetic n x n quad mesh, default 2048^2) through the public API (fvm_b200.models.FlowModelA).
Prints one JSON line: cell-updates/s over the timed SIMPLE iterations and the per-phase split.

  python tools/flow_cavity_bench.py [--size 2048] [--iters 20] [--warmup 3] [--mu 0.01]
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from fvm_b200 import capi, meshgen as G, models as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mu", type=float, default=0.01)
    a = ap.parse_args()
    lib = capi.default_lib()
    t0 = time.time()
    raw = G.quad_mesh(a.size, a.size)
    mesh = M.Mesh(raw)
    geom = M.GeomFields("geom")
    M.MeshMetricsCalculatorA(geom, [mesh], lib=lib).init()
    ff = M.FlowFields("flow")
    fm = M.FlowModelA(geom, ff, [mesh], lib=lib)
    fm.getBCMap()[4]["specifiedXVelocity"] = 1.0
    fm.getVCMap()[mesh.getID()]["viscosity"] = a.mu
    fm.getOptions().momentumTolerance = 1e-30
    fm.getOptions().continuityTolerance = 1e-30
    fm.init()
    setup_s = time.time() - t0
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        fm.advance(a.warmup)
    lib.synchronize()
    n0 = len(fm.timings)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        fm.advance(a.iters)
    lib.synchronize()
    wall = time.perf_counter() - t0
    tm = fm.timings[n0:]
    phases = {k: float(np.mean([t[k] for t in tm])) for k in
              ("momentum_assemble_ms", "momentum_solve_ms", "continuity_assemble_ms", "continuity_solve_ms")}
    dev_ms = sum(phases.values())
    out = {"workload": "lid-driven cavity, %dx%d quads (%d cells), mu=%g, SIMPLE with the reference's default "
                       "solvers (AMG rel 1e-1 / 20 cycles for momentum and pressure correction)" % (a.size, a.size, raw.n_cells, a.mu),
           "iterations": a.iters, "cell_updates_per_s_device": raw.n_cells / (dev_ms * 1e-3),
           "cell_updates_per_s_e2e": raw.n_cells * a.iters / wall, "ms_per_iteration_device": dev_ms,
           "ms_per_iteration_e2e": wall / a.iters * 1e3, "phase_ms": phases,
           "pressure_cycles_per_iteration": float(np.mean([t["pressure_iterations"] for t in tm])),
           "momentum_cycles_per_iteration": [float(np.mean([t["momentum_iterations"][k] for t in tm])) for k in range(3)],
           "last_norms": {"momentum": [float(v) for v in tm[-1]["momentum_norm"]], "continuity": float(tm[-1]["continuity_norm"])},
           "mesh_setup_s": setup_s}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
