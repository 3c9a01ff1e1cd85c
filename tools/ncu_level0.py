"""Driver for the `ncu --set full` capture of the dominant kernel: the level-0 multicolour
Gauss-Seidel colour pass (k_rows<GsRows>) and the fused residual + 1-norm (k_reduce1<ResidualRows>)
on the 256^3 thermal system. maxCoarseLevels = 0 keeps the hierarchy at level 0 only, so every
GsRows launch in this process is a level-0 launch (in bench.py they are interleaved with ~120
coarse-level launches per cycle).

  python tools/ncu_level0.py [cells_per_side]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from fvm_b200 import capi as X, meshgen as G  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    lib = X.default_lib()
    raw = G.hex_mesh(n, n, n)
    geo = G.metrics(raw)
    row, col = G.connectivity(raw)
    dm = X.DeviceMesh(lib, 3, raw.n_cells, raw.n_total, raw.face_cells, row, col, raw.group_offset, raw.group_count,
                      raw.group_id, raw.group_kind)
    dm.set_geometry(geo["face_area"], geo["face_area_mag"], geo["cell_centroid"], geo["cell_volume"],
                    ib_type=np.full(raw.n_total, -1, np.int32))
    ds = X.DeviceSystem(lib, dm)
    ds.fill_field(X.FIELD_X, 300.0)
    ds.set_bc(5, X.BC_DIRICHLET, [300.0])
    ds.set_bc(6, X.BC_DIRICHLET, [400.0])
    for g in (1, 2, 3, 4):
        ds.set_bc(g, X.BC_NEUMANN, [0.0])
    ds.assemble()
    o = lib.default_amg_opts()
    # optional second argument: number of coarse levels to keep (1: the level-0 restriction / prolongation kernels
    # InjectRows / CorrectRows appear as well)
    # (with coarse levels the iteration limit is >= 64, so that the 16-bit column copy is built as in a long solve)
    coarse = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    o.maxCoarseLevels, o.nMaxIterations, o.relativeTolerance = coarse, (70 if coarse else 6), 1e-30
    amg = X.DeviceAMG(lib, o)
    r0, r, it = amg.solve(ds)
    lv = amg.levels()
    print("%d cycles, residual %g -> %g, colours %s, column bytes per entry %s" % (it, r0, r, lv["colours"], lv["col_bytes"]))


if __name__ == "__main__":
    main()
