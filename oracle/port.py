"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/libfvmoracle.so -- our independent plain-C restatement (fvm_oracle.c) of
the reference's sequential algorithm for the hot path. Parity status: pinned by
tests/test_oracle_port.py against the reference's golden vectors and against oracle/_ref.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfvmoracle.so")

_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


class _Mesh(C.Structure):
    _fields_ = [("dim", C.c_int), ("nSelf", C.c_int), ("nTotal", C.c_int), ("nFaces", C.c_int), ("nGroups", C.c_int),
                ("faceCells", _ip), ("row", _ip), ("col", _ip), ("pairToCol", _ip), ("groupOffset", _ip),
                ("groupCount", _ip), ("groupKind", _ip), ("faceArea", _dp), ("faceAreaMag", _dp),
                ("cellCentroid", _dp), ("cellVolume", _dp)]


class _Bc(C.Structure):
    _fields_ = [("kind", C.c_int), ("p", C.c_double * 4)]


class _Opts(C.Structure):
    _fields_ = [("diffusion", C.c_int), ("convection", C.c_int), ("source", C.c_int), ("time_order", C.c_int),
                ("dt", C.c_double), ("underrelax", C.c_double), ("apply_bcs", C.c_int), ("eliminate_boundary", C.c_int)]


class AmgOpts(C.Structure):
    _fields_ = [("nMaxIterations", C.c_int), ("verbosity", C.c_int), ("relativeTolerance", C.c_double),
                ("absoluteTolerance", C.c_double), ("maxCoarseLevels", C.c_int), ("nPreSweeps", C.c_int),
                ("nPostSweeps", C.c_int), ("coarseGroupSize", C.c_int), ("weightRatioThreshold", C.c_double),
                ("cycleType", C.c_int), ("smootherType", C.c_int)]


def amg_opts(nMaxIterations=100, relativeTolerance=1e-8, absoluteTolerance=1e-50, maxCoarseLevels=30,
             nPreSweeps=0, nPostSweeps=1, coarseGroupSize=2, weightRatioThreshold=0.65, cycleType=0,
             smootherType=0):
    """reference defaults (F/AMG.cpp:14-22, F/LinearSolver.h:15-20)"""
    return AmgOpts(nMaxIterations, 0, relativeTolerance, absoluteTolerance, maxCoarseLevels, nPreSweeps,
                   nPostSweeps, coarseGroupSize, weightRatioThreshold, cycleType, smootherType)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)
        _lib = C.CDLL(LIB_PATH)
        _lib.fvmo_solve.restype = C.c_int
    return _lib


def _i(a):
    return np.ascontiguousarray(a, np.int32)


def _d(a):
    return np.ascontiguousarray(a, np.float64)


def _pi(a):
    return a.ctypes.data_as(_ip)


def _pd(a):
    return a.ctypes.data_as(_dp) if a is not None else None


BC_KINDS = dict(dirichlet=0, neumann=1, extrapolation=2, convective=3, radiative=4, mixed=5, interface=6,
                dirichlet_or_outflow=7)


class PortMesh:
    """Mesh arrays (connectivity + geometry dicts as produced by meshgen / refapi / golden files)."""

    def __init__(self, g):
        self.keep = dict(fc=_i(g["face_cells"]).reshape(-1), row=_i(g["cc_row"]), col=_i(g["cc_col"]),
                         go=_i(g["group_offset"]), gc=_i(g["group_count"]), gk=_i(g["group_kind"]),
                         fa=_d(g["face_area"]).reshape(-1), fam=_d(g["face_area_mag"]),
                         cc=_d(g["cell_centroid"]).reshape(-1), cv=_d(g["cell_volume"]))
        k = self.keep
        self.n_self, self.n_total = int(g["n_self"]), int(g["n_total"])
        self.n_faces = len(k["fc"]) // 2
        self.nnz = int(k["row"][-1])
        k["p2c"] = np.zeros(2 * self.n_faces, np.int32)
        lib().fvmo_pair_to_col(self.n_faces, _pi(k["fc"]), _pi(k["row"]), _pi(k["col"]), _pi(k["p2c"]))
        self.group_id = _i(g["group_id"])
        self.c = _Mesh(int(g["dim"]), self.n_self, self.n_total, self.n_faces, len(k["go"]), _pi(k["fc"]),
                       _pi(k["row"]), _pi(k["col"]), _pi(k["p2c"]), _pi(k["go"]), _pi(k["gc"]), _pi(k["gk"]),
                       _pd(k["fa"]), _pd(k["fam"]), _pd(k["cc"]), _pd(k["cv"]))
        self.weights = np.zeros(3 * self.nnz)
        lib().fvmo_ls_weights(C.byref(self.c), _pd(self.weights))

    @property
    def pair_to_col(self):
        return self.keep["p2c"].reshape(-1, 2)

    def gradient(self, x):
        x = _d(x)
        g = np.zeros(3 * self.n_total)
        lib().fvmo_gradient(C.byref(self.c), _pd(self.weights), _pd(x), _pd(g))
        return g.reshape(-1, 3)

    def assemble(self, x, bcs, diffusivity=None, source=None, face_flux=None, x_n1=None, x_n2=None,
                 density=None, cont_resid=None, diffusion=1, convection=0, time_order=0, dt=0.0,
                 underrelax=0.0, apply_bcs=1, eliminate_boundary=1):
        """bcs: {group id: (kind name, [params])}. Returns dict; x is copied (Dirichlet BCs edit it)."""
        x = _d(x).copy()
        grad = self.gradient(x).reshape(-1)
        k = _d(diffusivity) if diffusivity is not None else np.ones(self.n_total)
        s = _d(source) if source is not None else np.zeros(self.n_total)
        table = (_Bc * len(self.group_id))()
        for gi, gid in enumerate(self.group_id):
            table[gi].kind = -1
            if self.keep["gk"][gi] == 2:
                table[gi].kind = 6
            if int(gid) in bcs and self.keep["gk"][gi] != 0:
                kind, p = bcs[int(gid)]
                table[gi].kind = BC_KINDS[kind]
                for j, v in enumerate(list(p)[:4]):
                    table[gi].p[j] = float(v)
        o = _Opts(diffusion, convection, 1, time_order, dt, underrelax, apply_bcs, eliminate_boundary)
        nt, nf = self.n_total, self.n_faces
        out = dict(diag=np.zeros(nt), offdiag=np.zeros(self.nnz), b=np.zeros(nt), is_boundary=np.zeros(nt, np.int32),
                   bflux=np.zeros(nf), rflux=np.zeros(nf), coeffL=np.zeros(nf), coeffR=np.zeros(nf), x=x,
                   gradient=grad.reshape(-1, 3))
        opt = [None if a is None else _d(a) for a in (face_flux, x_n1, x_n2, density, cont_resid)]
        lib().fvmo_thermal_assemble(C.byref(self.c), C.byref(o), table, _pd(x), _pd(grad), _pd(k), _pd(s),
                                    _pd(opt[0]), _pd(opt[1]), _pd(opt[2]), _pd(opt[3]), _pd(opt[4]),
                                    _pd(out["diag"]), _pd(out["offdiag"]), _pd(out["b"]), _pi(out["is_boundary"]),
                                    _pd(out["bflux"]), _pd(out["rflux"]), _pd(out["coeffL"]), _pd(out["coeffR"]))
        return out

    def post_solve(self, a, delta):
        delta = _d(delta).copy()
        x = a["x"].copy()
        bflux = a["bflux"].copy()
        lib().fvmo_post_solve(C.byref(self.c), _pd(a["diag"]), _pd(a["offdiag"]), _pd(a["b"]), _pi(a["is_boundary"]),
                              _pd(delta), _pd(x), _pd(bflux), _pd(a["rflux"]), _pd(a["coeffL"]), _pd(a["coeffR"]))
        return x, bflux


def solve(n_self, row, col, diag, off, b, opts=None, n_ghost=0, is_boundary=None, bcgstab=False, x0=None):
    """Reference-algorithm AMG (or BCGStab + AMG) on CSR with separate diagonal, r = b + A x."""
    opts = opts or amg_opts()
    nt = n_self + n_ghost
    row, col, diag, off, b = _i(row), _i(col), _d(diag), _d(off), _d(b)
    x = np.zeros(nt) if x0 is None else _d(x0).copy()
    hist = np.zeros(1 << 14)
    nh = C.c_int(0)
    lv = np.zeros(64, np.int32)
    isb = None if is_boundary is None else _i(is_boundary)
    iters = lib().fvmo_solve(n_self, n_ghost, _pi(row), _pi(col), _pd(diag), _pd(off), _pd(b),
                             _pi(isb) if isb is not None else None, C.byref(opts), int(bool(bcgstab)), _pd(x),
                             _pd(hist), len(hist), C.byref(nh), _pi(lv))
    levels = []
    for v in lv:
        if v < 0:
            break
        levels.append(int(v))
    return dict(x=x, iters=iters, history=hist[:nh.value].copy(), levels=levels)


def thermal_reference(raw, conn, geo, k, bcs, x0=300.0, tol=1e-12):
    """One ThermalModel outer iteration by the port: assemble twice like smoke() does on the device
    (second assembly sees the Dirichlet values in the ghosts), solve, update."""
    g = dict(dim=raw.dim, n_self=raw.n_cells, n_total=raw.n_total)
    g.update(conn)
    g.update(geo)
    pm = PortMesh(g)
    all_bcs = {int(gid): ("neumann", [0.0]) for gid in g["group_id"][1:]}
    all_bcs.update({gid: (kind, [v]) for gid, (kind, v) in bcs.items()})
    x = np.full(raw.n_total, float(x0))
    a = pm.assemble(x, all_bcs, diffusivity=k)
    r = solve(pm.n_self, g["cc_row"], g["cc_col"], a["diag"], a["offdiag"], a["b"],
              amg_opts(nMaxIterations=5000, relativeTolerance=tol), n_ghost=pm.n_total - pm.n_self,
              is_boundary=a["is_boundary"])
    xs, _ = pm.post_solve(a, r["x"])
    return dict(diag=a["diag"], off=a["offdiag"], b=a["b"], x=xs)
