"""Repeat the same AMG solve several times under different kernel-path switches and count distinct
results (debug aid for run-to-run reproducibility)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from fvm_b200 import capi as X, meshgen as G  # noqa: E402


def systems(lib):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "mm226.npz")))
    yield "mm226", lambda: X.DeviceSystem(lib, raw=(int(g["n"]), 0, g["row"], g["col"], g["diag"], g["off"], g["b"])), None
    raw = G.hex_mesh(20, 22, 24, jitter=0.1, seed=4)
    geo = G.metrics(raw)
    row, col = G.connectivity(raw)
    dm = X.DeviceMesh(lib, 3, raw.n_cells, raw.n_total, raw.face_cells, row, col, raw.group_offset, raw.group_count,
                      raw.group_id, raw.group_kind)
    dm.set_geometry(geo["face_area"], geo["face_area_mag"], geo["cell_centroid"], geo["cell_volume"],
                    ib_type=np.full(raw.n_total, -1, np.int32))

    def mk():
        ds = X.DeviceSystem(lib, dm)
        ds.fill_field(X.FIELD_X, 300.0)
        ds.set_bc(5, X.BC_DIRICHLET, [300.0]); ds.set_bc(6, X.BC_DIRICHLET, [400.0])
        for gid in (1, 2, 3, 4):
            ds.set_bc(gid, X.BC_NEUMANN, [1.0])
        ds.assemble()
        return ds
    yield "hexj", mk, dm


def main():
    lib = X.default_lib()
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    for name, mk, _ in systems(lib):
        for label, env in (("default", {}), ("no_fused", {"FVMGPU_NO_FUSED": "1"}), ("tail_only", {"FVMGPU_COOP_ROWS": "0"})):
            for k in ("FVMGPU_NO_FUSED", "FVMGPU_COOP_ROWS"):
                os.environ.pop(k, None)
            os.environ.update(env)
            seen, cols = {}, set()
            for _ in range(reps):
                ds = mk()
                o = lib.default_amg_opts()
                o.nMaxIterations, o.relativeTolerance = 40, 1e-30
                amg = X.DeviceAMG(lib, o)
                amg.solve(ds)
                x = ds.get_field(X.FIELD_DELTA)
                cols.add(tuple(amg.levels()["colours"]))
                seen[x.tobytes()] = seen.get(x.tobytes(), 0) + 1
                amg.close(); ds.close()
            print("%-6s %-9s distinct results %d %s colourings %d" % (name, label, len(seen), sorted(seen.values()), len(cols)), flush=True)


if __name__ == "__main__":
    main()
