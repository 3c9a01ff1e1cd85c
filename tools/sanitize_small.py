"""Small end-to-end exercise of the library for compute-sanitizer (initcheck / memcheck): a raw AMG solve,
a thermal assemble + solve + update on a jittered hex mesh, BCGStab, and a few SIMPLE iterations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from fvm_b200 import capi as X, meshgen as G  # noqa: E402

lib = X.default_lib()
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "mm226.npz")))
for _ in range(2):
    ds = X.DeviceSystem(lib, raw=(int(g["n"]), 0, g["row"], g["col"], g["diag"], g["off"], g["b"]))
    amg = X.DeviceAMG(lib)
    print("mm226", amg.solve(ds))
    amg.close(); ds.close()
raw = G.hex_mesh(14, 12, 10, jitter=0.1, seed=4)
row, col = G.connectivity(raw)
dm = X.DeviceMesh(lib, 3, raw.n_cells, raw.n_total, raw.face_cells, row, col, raw.group_offset, raw.group_count,
                  raw.group_id, raw.group_kind)
dm.compute_geometry(raw.nodes, raw.face_node_count, raw.face_nodes)
ds = X.DeviceSystem(lib, dm)
ds.fill_field(X.FIELD_X, 300.0)
ds.set_bc(5, X.BC_DIRICHLET, [300.0]); ds.set_bc(6, X.BC_DIRICHLET, [400.0])
for gid in (1, 2, 3, 4):
    ds.set_bc(gid, X.BC_NEUMANN, [1.0])
ds.assemble()
amg = X.DeviceAMG(lib)
print("thermal", amg.solve(ds))
ds.post_solve_update()
ds.assemble()
print("bcgstab", amg.bcgstab(ds, 50, 1e-10, 1e-50))
ds.post_solve_update()
fl = X.DeviceFlow(lib, dm)
fl.fill_field(X.FLOW_VISCOSITY, 0.05)
for gid in (1, 2, 3, 4, 5):
    fl.set_bc(gid, X.FLOWBC_NOSLIP_WALL, [0, 0, 0])
fl.set_bc(6, X.FLOWBC_NOSLIP_WALL, [1.0, 0, 0])
fl.init()
o = fl.opts()
om = lib.default_amg_opts(); om.relativeTolerance, om.nMaxIterations, om.verbosity = 1e-1, 20, 0
am, ap = X.DeviceAMG(lib, om), X.DeviceAMG(lib, om)
for it in range(3):
    fl.assemble_momentum(o); print("mom", fl.solve_momentum(am)[0]); am.cleanup()
    fl.assemble_continuity(o); print("cont", fl.solve_continuity(ap, o)); ap.cleanup()
print("done")
