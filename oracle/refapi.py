"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_ref/libfvmref.so: the reference's own C++ hot path (compiled in
place from /root/reference by oracle/Makefile) behind the glue in oracle/ref_api.cpp.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product package (fvm_b200) never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfvmref.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class SolverCfg(C.Structure):
    """Public tunables of the reference solvers (F/AMG.h:74-81, F/LinearSolver.h:15-20)."""

    _fields_ = [
        ("kind", C.c_int),  # 0 AMG, 1 BCGStab + AMG preconditioner
        ("nMaxIterations", C.c_int),
        ("verbosity", C.c_int),
        ("relativeTolerance", C.c_double),
        ("absoluteTolerance", C.c_double),
        ("maxCoarseLevels", C.c_int),
        ("nPreSweeps", C.c_int),
        ("nPostSweeps", C.c_int),
        ("coarseGroupSize", C.c_int),
        ("weightRatioThreshold", C.c_double),
        ("cycleType", C.c_int),
        ("smootherType", C.c_int),
    ]


def solver_cfg(kind=0, nMaxIterations=100, verbosity=1, relativeTolerance=1e-8,
               absoluteTolerance=1e-50, maxCoarseLevels=30, nPreSweeps=0, nPostSweeps=1,
               coarseGroupSize=2, weightRatioThreshold=0.65, cycleType=0, smootherType=0):
    """Defaults are the reference's (F/AMG.cpp:14-22, F/LinearSolver.h:15-20)."""
    return SolverCfg(kind, nMaxIterations, verbosity, relativeTolerance, absoluteTolerance,
                     maxCoarseLevels, nPreSweeps, nPostSweeps, coarseGroupSize,
                     weightRatioThreshold, cycleType, smootherType)


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.fvmref_last_error.restype = C.c_char_p
        L.fvmref_mesh_from_cas.restype = C.c_void_p
        L.fvmref_mesh_from_cas.argtypes = [C.c_char_p]
        L.fvmref_mesh_from_raw.restype = C.c_void_p
        L.fvmref_mesh_from_raw.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_int, _ip, _ip, _ip,
                                           C.c_int, _ip]
        L.fvmref_mesh_from_raw_typed.restype = C.c_void_p
        L.fvmref_mesh_from_raw_typed.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_int, _ip, _ip, _ip,
                                                 C.c_int, _ip, C.c_int, _ip, _ip]
        L.fvmref_mesh_set_cell_geometry.restype = C.c_int
        L.fvmref_mesh_set_cell_geometry.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp]
        L.fvmref_mesh_from_raw_sym.restype = C.c_void_p
        L.fvmref_mesh_from_raw_sym.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_int, _ip, _ip, _ip,
                                               C.c_int, _ip, C.c_int, _ip]
        L.fvmref_mesh_free.argtypes = [C.c_void_p]
        L.fvmref_mesh_sizes.argtypes = [C.c_void_p, _ip]
        L.fvmref_mesh_connectivity.argtypes = [C.c_void_p] + [_ip] * 8
        L.fvmref_mesh_cell_nodes.argtypes = [C.c_void_p, _ip, C.c_void_p]
        L.fvmref_mesh_node_coordinates.argtypes = [C.c_void_p, _dp]
        L.fvmref_mesh_geometry.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _ip]
        L.fvmref_vacancy_create.restype = C.c_void_p
        L.fvmref_vacancy_create.argtypes = [C.c_void_p]
        L.fvmref_vacancy_free.argtypes = [C.c_void_p]
        L.fvmref_vacancy_set_bc.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_double]
        L.fvmref_vacancy_set_vc.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.fvmref_vacancy_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.fvmref_vacancy_set_solver.argtypes = [C.c_void_p, C.POINTER(SolverCfg)]
        L.fvmref_vacancy_init.argtypes = [C.c_void_p]
        L.fvmref_vacancy_field.restype = C.POINTER(C.c_double)
        L.fvmref_vacancy_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.fvmref_vacancy_advance.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.fvmref_vacancy_update_time.argtypes = [C.c_void_p]
        L.fvmref_vacancy_flux_integral.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.fvmref_species_create.restype = C.c_void_p
        L.fvmref_species_create.argtypes = [C.c_void_p, C.c_int]
        L.fvmref_species_free.argtypes = [C.c_void_p]
        L.fvmref_species_set_bc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_double]
        L.fvmref_species_set_vc.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_double]
        L.fvmref_species_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.fvmref_species_set_solver.argtypes = [C.c_void_p, C.POINTER(SolverCfg)]
        L.fvmref_species_init.argtypes = [C.c_void_p]
        L.fvmref_species_field.restype = C.POINTER(C.c_double)
        L.fvmref_species_field.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_int)]
        L.fvmref_species_advance.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.fvmref_species_update_time.argtypes = [C.c_void_p]
        L.fvmref_species_query.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.fvmref_thermal_create.restype = C.c_void_p
        L.fvmref_thermal_create.argtypes = [C.c_void_p]
        L.fvmref_thermal_free.argtypes = [C.c_void_p]
        L.fvmref_thermal_set_bc.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_double]
        L.fvmref_thermal_set_vc.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.fvmref_thermal_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.fvmref_thermal_set_solver.argtypes = [C.c_void_p, C.POINTER(SolverCfg)]
        L.fvmref_thermal_init.argtypes = [C.c_void_p]
        L.fvmref_thermal_field.restype = C.POINTER(C.c_double)
        L.fvmref_thermal_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.fvmref_thermal_heat_flux.argtypes = [C.c_void_p, C.c_int, _dp]
        L.fvmref_thermal_assemble.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _ip,
                                              C.POINTER(C.c_double)]
        L.fvmref_thermal_advance.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int,
                                             C.POINTER(C.c_double)]
        L.fvmref_thermal_advance_timed.argtypes = [C.c_void_p, _dp, C.c_char_p, C.c_int]
        L.fvmref_linsolve.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp,
                                      C.POINTER(SolverCfg), _dp, C.POINTER(C.c_double),
                                      C.POINTER(C.c_int), _ip, C.c_char_p, C.c_int,
                                      C.POINTER(C.c_double)]
        L.fvmref_flow_create.restype = C.c_void_p
        L.fvmref_flow_create.argtypes = [C.c_void_p]
        L.fvmref_flow_free.argtypes = [C.c_void_p]
        L.fvmref_flow_set_bc.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_double]
        L.fvmref_flow_set_vc.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.fvmref_flow_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.fvmref_flow_set_solver.argtypes = [C.c_void_p, C.c_int, C.POINTER(SolverCfg)]
        L.fvmref_flow_init.argtypes = [C.c_void_p]
        L.fvmref_flow_update_time.argtypes = [C.c_void_p]
        L.fvmref_flow_field.restype = C.POINTER(C.c_double)
        L.fvmref_flow_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.fvmref_flow_momentum_system.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.fvmref_flow_solve_momentum.argtypes = [C.c_void_p, _dp]
        L.fvmref_flow_continuity_system.argtypes = [C.c_void_p, _dp, _dp, _dp, _ip]
        L.fvmref_flow_solve_continuity.argtypes = [C.c_void_p, _dp]
        L.fvmref_flow_advance.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_double)]
        L.fvmref_electric_create.restype = C.c_void_p
        L.fvmref_electric_create.argtypes = [C.c_void_p]
        L.fvmref_electric_free.argtypes = [C.c_void_p]
        L.fvmref_electric_set_bc.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_double]
        L.fvmref_electric_set.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_double]
        L.fvmref_electric_set_solver.argtypes = [C.c_void_p, C.c_int, C.POINTER(SolverCfg)]
        L.fvmref_electric_init.argtypes = [C.c_void_p]
        L.fvmref_electric_field.restype = C.POINTER(C.c_double)
        L.fvmref_electric_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.fvmref_electric_potential_system.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.fvmref_electric_advance.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.fvmref_electric_update_time.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc is None or rc == -1:
        raise RuntimeError("reference: " + lib().fvmref_last_error().decode())


class RefMesh:
    """A reference `Mesh` + `GeomFields` after `MeshMetricsCalculator::init`."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference: " + lib().fvmref_last_error().decode())
        self.h = handle
        s = np.zeros(8, np.int32)
        _check(lib().fvmref_mesh_sizes(self.h, s))
        (self.dim, self.n_self, self.n_total, self.n_faces, self.nnz, self.n_groups,
         self.n_nodes, self.mesh_id) = [int(v) for v in s]
        self._conn = None
        self._geom = None

    @classmethod
    def from_cas(cls, path):
        return cls(lib().fvmref_mesh_from_cas(path.encode()))

    @classmethod
    def from_raw(cls, dim, n_cells, nodes, face_cells, face_nodes, face_node_count, face_group_size,
                 symmetry_groups=(), dielectric_groups=()):
        if len(dielectric_groups):
            nodes = np.ascontiguousarray(nodes, np.float64).reshape(-1, 3)
            fc = np.ascontiguousarray(face_cells, np.int32).reshape(-1)
            fn = np.ascontiguousarray(face_nodes, np.int32).reshape(-1)
            fnc = np.ascontiguousarray(face_node_count, np.int32)
            fgs = np.ascontiguousarray(face_group_size, np.int32)
            ids = np.ascontiguousarray(list(symmetry_groups) + list(dielectric_groups), np.int32)
            codes = np.ascontiguousarray([3] * len(symmetry_groups) + [4] * len(dielectric_groups), np.int32)
            return cls(lib().fvmref_mesh_from_raw_typed(dim, n_cells, len(nodes), nodes, len(fnc), fc, fn, fnc,
                                                        len(fgs), fgs, len(ids), ids, codes))
        if len(symmetry_groups):
            nodes = np.ascontiguousarray(nodes, np.float64).reshape(-1, 3)
            fc = np.ascontiguousarray(face_cells, np.int32).reshape(-1)
            fn = np.ascontiguousarray(face_nodes, np.int32).reshape(-1)
            fnc = np.ascontiguousarray(face_node_count, np.int32)
            fgs = np.ascontiguousarray(face_group_size, np.int32)
            sym = np.ascontiguousarray(list(symmetry_groups), np.int32)
            return cls(lib().fvmref_mesh_from_raw_sym(dim, n_cells, len(nodes), nodes, len(fnc), fc, fn, fnc,
                                                      len(fgs), fgs, len(sym), sym))
        nodes = np.ascontiguousarray(nodes, np.float64).reshape(-1, 3)
        fc = np.ascontiguousarray(face_cells, np.int32).reshape(-1)
        fn = np.ascontiguousarray(face_nodes, np.int32).reshape(-1)
        fnc = np.ascontiguousarray(face_node_count, np.int32)
        fgs = np.ascontiguousarray(face_group_size, np.int32)
        return cls(lib().fvmref_mesh_from_raw(dim, n_cells, len(nodes), nodes, len(fnc), fc, fn, fnc,
                                              len(fgs), fgs))

    def connectivity(self):
        if self._conn is None:
            fc = np.zeros(2 * self.n_faces, np.int32)
            row = np.zeros(self.n_total + 1, np.int32)
            col = np.zeros(self.nnz, np.int32)
            p2c = np.zeros(2 * self.n_faces, np.int32)
            go = np.zeros(self.n_groups, np.int32)
            gc = np.zeros(self.n_groups, np.int32)
            gi = np.zeros(self.n_groups, np.int32)
            gk = np.zeros(self.n_groups, np.int32)
            _check(lib().fvmref_mesh_connectivity(self.h, fc, row, col, p2c, go, gc, gi, gk))
            self._conn = dict(face_cells=fc.reshape(-1, 2), cc_row=row, cc_col=col,
                              pair_to_col=p2c.reshape(-1, 2), group_offset=go, group_count=gc,
                              group_id=gi, group_kind=gk)
        return self._conn

    def cell_nodes(self):
        """Mesh::getCellNodes() of the self cells as (row, col): the reference's canonical node order per cell type."""
        row = np.zeros(self.n_self + 1, np.int32)
        _check(lib().fvmref_mesh_cell_nodes(self.h, row, None))
        col = np.zeros(int(row[-1]), np.int32)
        _check(lib().fvmref_mesh_cell_nodes(self.h, row, col.ctypes.data_as(C.c_void_p)))
        return row, col

    def node_coordinates(self):
        xyz = np.zeros(3 * self.n_nodes)
        _check(lib().fvmref_mesh_node_coordinates(self.h, xyz))
        return xyz.reshape(-1, 3)

    def geometry(self):
        if self._geom is None:
            fa = np.zeros(3 * self.n_faces)
            fam = np.zeros(self.n_faces)
            fx = np.zeros(3 * self.n_faces)
            cx = np.zeros(3 * self.n_total)
            cv = np.zeros(self.n_total)
            ib = np.zeros(self.n_total, np.int32)
            _check(lib().fvmref_mesh_geometry(self.h, fa, fam, fx, cx, cv, ib))
            self._geom = dict(face_area=fa.reshape(-1, 3), face_area_mag=fam,
                              face_centroid=fx.reshape(-1, 3), cell_centroid=cx.reshape(-1, 3),
                              cell_volume=cv, ib_type=ib)
        return self._geom

    def set_cell_geometry(self, first, centroid, volume):
        """overwrite centroid / volume of `len(volume)` cells from `first` on in the reference's GeomFields"""
        c = np.ascontiguousarray(centroid, np.float64).reshape(-1)
        v = np.ascontiguousarray(volume, np.float64)
        _check(lib().fvmref_mesh_set_cell_geometry(self.h, int(first), len(v), c, v))
        self._geom = None

    def close(self):
        if self.h:
            lib().fvmref_mesh_free(self.h)
            self.h = None


class RefThermal:
    """The reference `ThermalModel<double>` on a RefMesh (F/ThermalModel.h:28-52)."""

    def __init__(self, mesh):
        self.mesh = mesh
        self.h = lib().fvmref_thermal_create(mesh.h)
        if not self.h:
            raise RuntimeError("reference: " + lib().fvmref_last_error().decode())

    def set_bc(self, gid, bc_type="", **vars_):
        _check(lib().fvmref_thermal_set_bc(self.h, gid, bc_type.encode(), b"", 0.0))
        for k, v in vars_.items():
            _check(lib().fvmref_thermal_set_bc(self.h, gid, b"", k.encode(), float(v)))

    def set_vc(self, name, value):
        _check(lib().fvmref_thermal_set_vc(self.h, name.encode(), float(value)))

    def set_option(self, name, value):
        _check(lib().fvmref_thermal_set_option(self.h, name.encode(), float(value)))

    def set_solver(self, cfg):
        self._cfg = cfg
        _check(lib().fvmref_thermal_set_solver(self.h, C.byref(cfg)))

    def init(self):
        _check(lib().fvmref_thermal_init(self.h))

    def field(self, name):
        """numpy VIEW of the reference's host Array (writes go to the model)."""
        n = C.c_int(0)
        p = lib().fvmref_thermal_field(self.h, name.encode(), C.byref(n))
        if not p:
            raise RuntimeError("reference: " + lib().fvmref_last_error().decode())
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def heat_flux(self, gid, count):
        out = np.zeros(count)
        _check(lib().fvmref_thermal_heat_flux(self.h, gid, out))
        return out

    def assemble(self, stage=1):
        m = self.mesh
        diag = np.zeros(m.n_total)
        off = np.zeros(m.nnz)
        b = np.zeros(m.n_total)
        x = np.zeros(m.n_total)
        isb = np.zeros(m.n_total, np.int32)
        sec = C.c_double(0)
        _check(lib().fvmref_thermal_assemble(self.h, stage, diag, off, b, x, isb, C.byref(sec)))
        return dict(diag=diag, offdiag=off, b=b, x=x, is_boundary=isb, seconds=sec.value)

    def advance(self, niter=1):
        buf = C.create_string_buffer(1 << 16)
        sec = C.c_double(0)
        _check(lib().fvmref_thermal_advance(self.h, niter, buf, len(buf), C.byref(sec)))
        return buf.value.decode(), sec.value

    def advance_timed(self):
        t = np.zeros(8)
        buf = C.create_string_buffer(1 << 16)
        _check(lib().fvmref_thermal_advance_timed(self.h, t, buf, len(buf)))
        return dict(assemble_s=t[0], solve_s=t[1], update_s=t[2], cycles=int(t[3]), rnorm0=t[4],
                    text=buf.value.decode())

    def close(self):
        if self.h:
            lib().fvmref_thermal_free(self.h)
            self.h = None


class RefVacancy:
    """The reference `VacancyModel<double>` on a RefMesh (F/VacancyModel.h:18-55)."""

    def __init__(self, mesh):
        self.mesh = mesh
        self.h = lib().fvmref_vacancy_create(mesh.h)
        if not self.h:
            raise RuntimeError("reference: " + lib().fvmref_last_error().decode())

    def set_bc(self, gid, bc_type="", **vars_):
        _check(lib().fvmref_vacancy_set_bc(self.h, gid, bc_type.encode(), b"", 0.0))
        for k, v in vars_.items():
            _check(lib().fvmref_vacancy_set_bc(self.h, gid, b"", k.encode(), float(v)))

    def set_vc(self, name, value):
        _check(lib().fvmref_vacancy_set_vc(self.h, name.encode(), float(value)))

    def set_option(self, name, value):
        _check(lib().fvmref_vacancy_set_option(self.h, name.encode(), float(value)))

    def set_solver(self, cfg):
        self._cfg = cfg
        _check(lib().fvmref_vacancy_set_solver(self.h, C.byref(cfg)))

    def init(self):
        _check(lib().fvmref_vacancy_init(self.h))

    def field(self, name):
        n = C.c_int(0)
        p = lib().fvmref_vacancy_field(self.h, name.encode(), C.byref(n))
        if not p:
            raise RuntimeError("reference: " + lib().fvmref_last_error().decode())
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def advance(self, niter=1):
        buf = C.create_string_buffer(1 << 16)
        _check(lib().fvmref_vacancy_advance(self.h, niter, buf, len(buf)))
        return buf.value.decode()

    def update_time(self):
        _check(lib().fvmref_vacancy_update_time(self.h))

    def flux_integral(self, gid):
        out = C.c_double(0)
        _check(lib().fvmref_vacancy_flux_integral(self.h, gid, C.byref(out)))
        return out.value

    def close(self):
        if self.h:
            lib().fvmref_vacancy_free(self.h)
            self.h = None


class RefSpecies:
    """The reference `SpeciesModel<double>` on a RefMesh (F/SpeciesModel.h:17-55)."""

    def __init__(self, mesh, n_species):
        self.mesh = mesh
        self.h = lib().fvmref_species_create(mesh.h, n_species)
        if not self.h:
            raise RuntimeError("reference: " + lib().fvmref_last_error().decode())

    def set_bc(self, species, gid, bc_type="", **vars_):
        _check(lib().fvmref_species_set_bc(self.h, species, gid, bc_type.encode(), b"", 0.0))
        for k, v in vars_.items():
            _check(lib().fvmref_species_set_bc(self.h, species, gid, b"", k.encode(), float(v)))

    def set_vc(self, species, name, value):
        _check(lib().fvmref_species_set_vc(self.h, species, name.encode(), float(value)))

    def set_option(self, name, value):
        _check(lib().fvmref_species_set_option(self.h, name.encode(), float(value)))

    def set_solver(self, cfg):
        self._cfg = cfg
        _check(lib().fvmref_species_set_solver(self.h, C.byref(cfg)))

    def init(self):
        _check(lib().fvmref_species_init(self.h))

    def field(self, species, name):
        n = C.c_int(0)
        p = lib().fvmref_species_field(self.h, species, name.encode(), C.byref(n))
        if not p:
            raise RuntimeError("reference: " + lib().fvmref_last_error().decode())
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def advance(self, niter=1):
        buf = C.create_string_buffer(1 << 16)
        _check(lib().fvmref_species_advance(self.h, niter, buf, len(buf)))
        return buf.value.decode()

    def update_time(self):
        _check(lib().fvmref_species_update_time(self.h))

    def _query(self, species, what, arg=0):
        out = C.c_double(0)
        _check(lib().fvmref_species_query(self.h, species, what, arg, C.byref(out)))
        return out.value

    def mass_flux_integral(self, species, gid):
        return self._query(species, 0, gid)

    def average_mass_fraction(self, species):
        return self._query(species, 1)

    def mass_fraction_residual(self, species):
        return self._query(species, 2)

    def close(self):
        if self.h:
            lib().fvmref_species_free(self.h)
            self.h = None


def linsolve(n_self, row, col, diag, offdiag, b, cfg, n_ghost=0):
    """Run the reference AMG / BCGStab on a raw CSR-with-separate-diagonal system
    (sign convention r = b + A x, F/CRMatrix.h:407-426). Returns dict."""
    n = n_self + n_ghost
    x = np.zeros(n)
    rn0 = C.c_double(0)
    it = C.c_int(0)
    lv = np.zeros(64, np.int32)
    buf = C.create_string_buffer(1 << 16)
    sec = C.c_double(0)
    _check(lib().fvmref_linsolve(n_self, n_ghost, np.ascontiguousarray(row, np.int32),
                                 np.ascontiguousarray(col, np.int32),
                                 np.ascontiguousarray(diag, np.float64),
                                 np.ascontiguousarray(offdiag, np.float64),
                                 np.ascontiguousarray(b, np.float64), C.byref(cfg), x,
                                 C.byref(rn0), C.byref(it), lv, buf, len(buf), C.byref(sec)))
    levels = []
    for v in lv:
        if v < 0:
            break
        levels.append(int(v))
    return dict(x=x, rnorm0=rn0.value, iters=it.value, levels=levels, text=buf.value.decode(),
                seconds=sec.value)


class RefFlow:
    """The reference `FlowModel<double>` (SIMPLE) on a RefMesh (F/FlowModel.h:17-95)."""

    def __init__(self, mesh):
        self.mesh = mesh
        self.h = lib().fvmref_flow_create(mesh.h)
        if not self.h:
            raise RuntimeError("reference: " + lib().fvmref_last_error().decode())

    def set_bc(self, gid, bc_type="", **vars_):
        _check(lib().fvmref_flow_set_bc(self.h, gid, bc_type.encode(), b"", 0.0))
        for k, v in vars_.items():
            _check(lib().fvmref_flow_set_bc(self.h, gid, b"", k.encode(), float(v)))

    def set_vc(self, name, value):
        _check(lib().fvmref_flow_set_vc(self.h, name.encode(), float(value)))

    def set_option(self, name, value):
        _check(lib().fvmref_flow_set_option(self.h, name.encode(), float(value)))

    def set_solver(self, which, cfg):
        """which: 0 momentum, 1 pressure correction"""
        setattr(self, "_cfg%d" % which, cfg)
        _check(lib().fvmref_flow_set_solver(self.h, which, C.byref(cfg)))

    def init(self):
        _check(lib().fvmref_flow_init(self.h))

    def field(self, name):
        """numpy VIEW of the reference's host Array (vectors AoS: 3 or 9 doubles per cell)."""
        n = C.c_int(0)
        p = lib().fvmref_flow_field(self.h, name.encode(), C.byref(n))
        if not p:
            raise RuntimeError("reference: " + lib().fvmref_last_error().decode())
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def update_time(self):
        _check(lib().fvmref_flow_update_time(self.h))

    def momentum_system(self):
        m = self.mesh
        diag, off, b = np.zeros(3 * m.n_total), np.zeros(m.nnz), np.zeros(3 * m.n_total)
        _check(lib().fvmref_flow_momentum_system(self.h, diag, off, b))
        return dict(diag=diag.reshape(-1, 3), offdiag=off, b=b.reshape(-1, 3))

    def solve_momentum(self):
        r = np.zeros(3)
        _check(lib().fvmref_flow_solve_momentum(self.h, r))
        return r

    def continuity_system(self):
        m = self.mesh
        diag, off, b = np.zeros(m.n_total), np.zeros(m.nnz), np.zeros(m.n_total)
        isb = np.zeros(m.n_total, np.int32)
        _check(lib().fvmref_flow_continuity_system(self.h, diag, off, b, isb))
        return dict(diag=diag, offdiag=off, b=b, is_boundary=isb)

    def solve_continuity(self):
        r = np.zeros(1)
        _check(lib().fvmref_flow_solve_continuity(self.h, r))
        return float(r[0])

    def advance(self, niter=1):
        buf = C.create_string_buffer(1 << 18)
        sec = C.c_double(0)
        rc = lib().fvmref_flow_advance(self.h, niter, buf, len(buf), C.byref(sec))
        _check(rc)
        return bool(rc), buf.value.decode(), sec.value

    def close(self):
        if self.h:
            lib().fvmref_flow_free(self.h)
            self.h = None


class RefElectric:
    """The reference `ElectricModel<double>` on a RefMesh (F/ElectricModel.h)."""

    def __init__(self, mesh):
        self.mesh = mesh
        self.h = lib().fvmref_electric_create(mesh.h)
        if not self.h:
            raise RuntimeError("reference: " + lib().fvmref_last_error().decode())

    def set_bc(self, gid, bc_type="", **vars_):
        _check(lib().fvmref_electric_set_bc(self.h, gid, bc_type.encode(), b"", 0.0))
        for k, v in vars_.items():
            _check(lib().fvmref_electric_set_bc(self.h, gid, b"", k.encode(), float(v)))

    def set_vc(self, name, value):
        _check(lib().fvmref_electric_set(self.h, 0, name.encode(), float(value)))

    def set_option(self, name, value):
        _check(lib().fvmref_electric_set(self.h, 1, name.encode(), float(value)))

    def set_constant(self, name, value):
        _check(lib().fvmref_electric_set(self.h, 2, name.encode(), float(value)))

    def set_solver(self, which, cfg):
        """which: 0 electrostatics, 1 charge transport"""
        setattr(self, "_cfg%d" % which, cfg)
        _check(lib().fvmref_electric_set_solver(self.h, which, C.byref(cfg)))

    def init(self):
        _check(lib().fvmref_electric_init(self.h))

    def field(self, name):
        n = C.c_int(0)
        p = lib().fvmref_electric_field(self.h, name.encode(), C.byref(n))
        if not p:
            raise RuntimeError("reference: " + lib().fvmref_last_error().decode())
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def potential_system(self):
        m = self.mesh
        diag, off, b = np.zeros(m.n_total), np.zeros(m.nnz), np.zeros(m.n_total)
        _check(lib().fvmref_electric_potential_system(self.h, diag, off, b))
        return dict(diag=diag, offdiag=off, b=b)

    def advance(self, niter=1):
        buf = C.create_string_buffer(1 << 18)
        rc = lib().fvmref_electric_advance(self.h, niter, buf, len(buf))
        _check(rc)
        return bool(rc), buf.value.decode()

    def update_time(self):
        _check(lib().fvmref_electric_update_time(self.h))

    def close(self):
        if self.h:
            lib().fvmref_electric_free(self.h)
            self.h = None
