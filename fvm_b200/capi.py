"""ctypes binding of libfvmgpu.so (include/fvmgpu.h) -- the thin layer every Python caller uses.

There is no CPU fallback: if the CUDA library is missing or no device is present the calls
raise `FvmGpuError` (the analogue of the reference's CException -> RuntimeError, F/baseExt.i:49-59).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfvmgpu.so")


class FvmGpuError(RuntimeError):
    pass


# enums of fvmgpu.h
GROUP_INTERIOR, GROUP_BOUNDARY, GROUP_INTERFACE, GROUP_SYMMETRY, GROUP_DIELECTRIC_INTERFACE = 0, 1, 2, 3, 4
(BC_DIRICHLET, BC_NEUMANN, BC_EXTRAPOLATION, BC_CONVECTIVE, BC_RADIATIVE, BC_MIXED, BC_INTERFACE,
 BC_DIRICHLET_OR_OUTFLOW, BC_DIELECTRIC_INTERFACE) = range(9)
(FIELD_X, FIELD_DIFFUSIVITY, FIELD_SOURCE, FIELD_FACE_FLUX, FIELD_X_N1, FIELD_X_N2, FIELD_DENSITY,
 FIELD_CONT_RESID, FIELD_GRADIENT, FIELD_BFLUX, FIELD_DELTA, FIELD_B, FIELD_BFLUX_BOUNDARY) = range(13)
CYCLE_V, CYCLE_W, CYCLE_F = 0, 1, 2
SMOOTHER_GAUSS_SEIDEL, SMOOTHER_JACOBI = 0, 1


class AssembleOpts(C.Structure):
    _fields_ = [("diffusion", C.c_int), ("convection", C.c_int), ("source", C.c_int),
                ("time_order", C.c_int), ("dt", C.c_double), ("underrelax", C.c_double),
                ("apply_bcs", C.c_int), ("eliminate_boundary", C.c_int), ("interface_thickness", C.c_double)]


class AmgOpts(C.Structure):
    _fields_ = [("nMaxIterations", C.c_int), ("verbosity", C.c_int),
                ("relativeTolerance", C.c_double), ("absoluteTolerance", C.c_double),
                ("maxCoarseLevels", C.c_int), ("nPreSweeps", C.c_int), ("nPostSweeps", C.c_int),
                ("coarseGroupSize", C.c_int), ("weightRatioThreshold", C.c_double),
                ("cycleType", C.c_int), ("smootherType", C.c_int)]


class FlowOpts(C.Structure):
    """fvmgpu_flow_opts"""

    _fields_ = [("momentumURF", C.c_double), ("pressureURF", C.c_double), ("transient", C.c_int),
                ("time_order", C.c_int), ("dt", C.c_double), ("correctVelocity", C.c_int),
                ("operatingPressure", C.c_double), ("operatingTemperature", C.c_double),
                ("molecularWeight", C.c_double), ("incompressible", C.c_int)]


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_dpn = C.c_void_p  # nullable double*
_ipn = C.c_void_p  # nullable int*

# name -> (restype, argtypes); every symbol include/fvmgpu.h declares
SIGNATURES = {
    "fvmgpu_amg_default_opts": (None, [C.POINTER(AmgOpts)]),
    "fvmgpu_init": (C.c_int, [C.c_int]),
    "fvmgpu_shutdown": (C.c_int, []),
    "fvmgpu_last_error": (C.c_char_p, []),
    "fvmgpu_version": (C.c_int, []),
    "fvmgpu_device_info": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "fvmgpu_synchronize": (C.c_int, []),
    "fvmgpu_timer_start": (C.c_int, [C.c_int]),
    "fvmgpu_timer_stop": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "fvmgpu_counters": (C.c_int, [C.POINTER(C.c_longlong)] * 3),
    "fvmgpu_flush_l2": (C.c_int, []),
    "fvmgpu_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_ulonglong]),
    "fvmgpu_host_free": (C.c_int, [C.c_void_p]),
    "fvmgpu_profile_begin": (C.c_int, []),
    "fvmgpu_profile_end": (C.c_int, [C.c_int, C.c_char_p, C.c_int,
                                     np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"),
                                     np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"),
                                     np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS"),
                                     C.POINTER(C.c_int)]),
    "fvmgpu_mesh_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _ip,
                                     C.c_int, _ip, _ip, _ip, _ip]),
    "fvmgpu_mesh_set_geometry": (C.c_int, [_vp, _dp, _dp, _dpn, _dp, _dp, _ipn]),
    "fvmgpu_mesh_compute_geometry": (C.c_int, [_vp, C.c_int, _dp, _ip, _ip, _dpn, _dpn, _dpn, _dpn, _dpn]),
    "fvmgpu_mesh_set_halo": (C.c_int, [_vp, C.c_int, _ip, _ip, _ip, _ip, _ip]),
    "fvmgpu_mesh_destroy": (C.c_int, [_vp]),
    "fvmgpu_mesh_download_pair_to_col": (C.c_int, [_vp, _ip]),
    "fvmgpu_mesh_download_gradient_weights": (C.c_int, [_vp, _dp]),
    "fvmgpu_system_create": (C.c_int, [C.POINTER(_vp), _vp]),
    "fvmgpu_system_create_raw": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp]),
    "fvmgpu_system_destroy": (C.c_int, [_vp]),
    "fvmgpu_system_set_field": (C.c_int, [_vp, C.c_int, _dp, C.c_longlong]),
    "fvmgpu_system_fill_field": (C.c_int, [_vp, C.c_int, C.c_double]),
    "fvmgpu_system_get_field": (C.c_int, [_vp, C.c_int, _dp, C.c_longlong]),
    "fvmgpu_system_set_bc": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int, _dpn]),
    "fvmgpu_compute_gradient": (C.c_int, [_vp]),
    "fvmgpu_assemble": (C.c_int, [_vp, C.POINTER(AssembleOpts)]),
    "fvmgpu_download_system": (C.c_int, [_vp, _dp, _dp, _dp, _ipn]),
    "fvmgpu_amg_create": (C.c_int, [C.POINTER(_vp), C.POINTER(AmgOpts)]),
    "fvmgpu_amg_set_opts": (C.c_int, [_vp, C.POINTER(AmgOpts)]),
    "fvmgpu_amg_solve": (C.c_int, [_vp, _vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "fvmgpu_amg_smooth": (C.c_int, [_vp, _vp]),
    "fvmgpu_amg_cleanup": (C.c_int, [_vp]),
    "fvmgpu_amg_destroy": (C.c_int, [_vp]),
    "fvmgpu_amg_levels": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int),
                                    np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"),
                                    np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"), _ip]),
    "fvmgpu_amg_level_order": (C.c_int, [_vp, C.c_int, C.c_longlong, _ip, C.POINTER(C.c_int),
                                         np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")]),
    "fvmgpu_debug_set_aggregator": (C.c_int, [_vp, _vp]),
    "fvmgpu_debug_tail_trace": (C.c_int, [C.c_int, np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS"),
                                          _ip, C.POINTER(C.c_int)]),
    "fvmgpu_amg_level_col_bytes": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double)]),
    "fvmgpu_amg_last_timing": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "fvmgpu_solver_history": (C.c_int, [_vp, C.c_int, _dp, C.POINTER(C.c_int)]),
    "fvmgpu_bcgstab_solve": (C.c_int, [_vp, _vp, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "fvmgpu_bcgstab_ilu0_solve": (C.c_int, [_vp, _vp, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double),
                                            C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "fvmgpu_ilu0_solve": (C.c_int, [_vp, _vp, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "fvmgpu_cg_solve": (C.c_int, [_vp, _vp, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double),
                                  C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "fvmgpu_jacobi_solve": (C.c_int, [_vp, _vp, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "fvmgpu_post_solve_update": (C.c_int, [_vp]),
    "fvmgpu_electric_field": (C.c_int, [_vp, _dpn]),
    "fvmgpu_electric_drift_flux": (C.c_int, [_vp, _vp, C.c_double, C.c_double, C.c_int, _ip, _dpn]),
    "fvmgpu_flow_create": (C.c_int, [C.POINTER(_vp), _vp]),
    "fvmgpu_flow_destroy": (C.c_int, [_vp]),
    "fvmgpu_flow_set_field": (C.c_int, [_vp, C.c_int, _dp, C.c_longlong]),
    "fvmgpu_flow_fill_field": (C.c_int, [_vp, C.c_int, C.c_double]),
    "fvmgpu_flow_get_field": (C.c_int, [_vp, C.c_int, _dp, C.c_longlong]),
    "fvmgpu_flow_set_bc": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int]),
    "fvmgpu_flow_init": (C.c_int, [_vp]),
    "fvmgpu_flow_set_reference_cell": (C.c_int, [_vp, C.c_int]),
    "fvmgpu_flow_assemble_momentum": (C.c_int, [_vp, C.POINTER(FlowOpts)]),
    "fvmgpu_flow_download_momentum": (C.c_int, [_vp, _dp, _dp, _dp]),
    "fvmgpu_flow_solve_momentum": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_double, C.c_double, _dp, _ip]),
    "fvmgpu_flow_assemble_continuity": (C.c_int, [_vp, C.POINTER(FlowOpts)]),
    "fvmgpu_flow_download_continuity": (C.c_int, [_vp, _dp, _dp, _dp, _ipn]),
    "fvmgpu_flow_solve_continuity": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_double, C.c_double,
                                               C.POINTER(FlowOpts), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "fvmgpu_comm_unique_id": (C.c_int, [C.c_char_p]),
    "fvmgpu_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_char_p]),
    "fvmgpu_comm_destroy": (C.c_int, []),
    "fvmgpu_comm_counters": (C.c_int, [C.POINTER(C.c_longlong)]),
    "fvmgpu_system_halo_exchange": (C.c_int, [_vp, C.c_int]),
}


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _nullable(a, dtype):
    if a is None:
        return None, None
    arr = np.ascontiguousarray(a, dtype=dtype)
    return arr, arr.ctypes.data_as(C.c_void_p)


def _demangle(name):
    """'N6fvmgpu6GsRowsE' -> 'GsRows' (Itanium nested name of a functor in namespace fvmgpu)."""
    import re
    m = re.match(r"^N6fvmgpu(\d+)", name)
    if m:
        k = int(m.group(1))
        s = name[len("N6fvmgpu") + len(m.group(1)):]
        return s[:k]
    return name


class Lib:
    """A loaded libfvmgpu.so with typed entry points; `call()` raises FvmGpuError on failure."""

    def __init__(self, path=None):
        path = path or LIB_PATH
        if not os.path.exists(path):
            raise FvmGpuError(
                "libfvmgpu.so not built (%s). Build it with `python -m fvm_b200.build`; "
                "there is no CPU fallback." % path)
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args
        self._initialised = False

    def call(self, name, *args):
        rc = getattr(self.dll, name)(*args)
        if rc != 0:
            raise FvmGpuError(self.dll.fvmgpu_last_error().decode())

    def init(self, device=0):
        self.call("fvmgpu_init", device)
        self._initialised = True

    def device_info(self):
        name = C.create_string_buffer(256)
        sm = C.c_int(0)
        mem = C.c_double(0)
        self.call("fvmgpu_device_info", name, 256, C.byref(sm), C.byref(mem))
        return name.value.decode(), sm.value, mem.value

    def counters(self):
        a, b, c = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        self.dll.fvmgpu_counters(C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def timer_start(self, slot=0):
        self.call("fvmgpu_timer_start", slot)

    def timer_stop(self, slot=0):
        ms = C.c_double(0)
        self.call("fvmgpu_timer_stop", slot, C.byref(ms))
        return ms.value

    def synchronize(self):
        self.call("fvmgpu_synchronize")

    def flush_l2(self):
        self.call("fvmgpu_flush_l2")

    def pinned_empty(self, shape, dtype=np.float64):
        """numpy array in page-locked host memory (fvmgpu_host_alloc), freed with the array: field arrays allocated
        this way go to and from the device at PCIe speed."""
        import weakref
        shape = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p(0)
        self.call("fvmgpu_host_alloc", C.byref(p), max(nbytes, 1))
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        weakref.finalize(buf, self.dll.fvmgpu_host_free, p)
        return arr

    def pinned_full(self, shape, value, dtype=np.float64):
        a = self.pinned_empty(shape, dtype)
        a[...] = value
        return a

    def profile_begin(self):
        self.call("fvmgpu_profile_begin")

    def profile_end(self, cap=4096):
        """-> list of dicts(name, rows, launches, ms) aggregated per (kernel class, rows)."""
        stride = 96
        names = C.create_string_buffer(cap * stride)
        rows = np.zeros(cap, np.int64)
        launches = np.zeros(cap, np.int64)
        ms = np.zeros(cap)
        n = C.c_int(0)
        self.call("fvmgpu_profile_end", cap, names, stride, rows, launches, ms, C.byref(n))
        out = []
        raw = names.raw
        for i in range(min(n.value, cap)):
            nm = raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode()
            level = -1
            if "@L" in nm:
                nm, lv = nm.rsplit("@L", 1)
                level = int(lv)
            out.append(dict(name=_demangle(nm), rows=int(rows[i]), launches=int(launches[i]), ms=float(ms[i]),
                            level=level))
        return out

    def tail_trace(self, cap=8192):
        """(times_ns, tags) of the last fused coarse-level kernel launch (needs FVMGPU_TAIL_TRACE=1)."""
        t = np.zeros(cap, np.uint64)
        g = np.zeros(cap, np.int32)
        n = C.c_int(0)
        self.call("fvmgpu_debug_tail_trace", cap, t, g, C.byref(n))
        return t[:n.value].copy(), g[:n.value].copy()

    def set_aggregator(self, fn_ptr, user=None):
        """Verification hook (include/fvmgpu.h: fvmgpu_debug_set_aggregator): fn_ptr is the address of a C function
        with the fvmgpu_aggregate_fn signature (or None to restore the library's own parallel pairing)."""
        self.call("fvmgpu_debug_set_aggregator", C.c_void_p(fn_ptr) if fn_ptr else None, user)

    # ---- multi-GPU
    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        self.call("fvmgpu_comm_unique_id", buf)
        return buf.raw

    def comm_init(self, nranks, rank, unique_id=None):
        self.call("fvmgpu_comm_init", int(nranks), int(rank), unique_id if unique_id is not None else b"\0" * 128)
        self.nranks, self.rank = int(nranks), int(rank)

    def comm_destroy(self):
        self.call("fvmgpu_comm_destroy")

    def comm_collectives(self):
        n = C.c_longlong(0)
        self.call("fvmgpu_comm_counters", C.byref(n))
        return n.value

    def default_amg_opts(self):
        o = AmgOpts()
        self.dll.fvmgpu_amg_default_opts(C.byref(o))
        return o


(FLOW_VELOCITY, FLOW_PRESSURE, FLOW_DENSITY, FLOW_VISCOSITY, FLOW_MASS_FLUX, FLOW_FACE_PRESSURE,
 FLOW_PRESSURE_GRADIENT, FLOW_VELOCITY_GRADIENT, FLOW_CONT_RESID, FLOW_MOM_AP, FLOW_PREV_VELOCITY,
 FLOW_VELOCITY_N1, FLOW_VELOCITY_N2) = range(13)
FLOWBC_NOSLIP_WALL = 0
FLOWBC_SYMMETRY = 1
FLOWBC_VELOCITY = 2
FLOWBC_PRESSURE = 3
FLOWBC_SLIP_JUMP = 4
_FLOW_WIDTH = {FLOW_VELOCITY: 3, FLOW_PRESSURE_GRADIENT: 3, FLOW_VELOCITY_GRADIENT: 9, FLOW_MOM_AP: 3,
               FLOW_PREV_VELOCITY: 3, FLOW_VELOCITY_N1: 3, FLOW_VELOCITY_N2: 3}
_FLOW_FACE = (FLOW_MASS_FLUX, FLOW_FACE_PRESSURE)


class DeviceFlow:
    """Device mirror of FlowFields + the momentum / pressure-correction systems of one mesh."""

    def __init__(self, lib, mesh):
        self.lib, self.mesh = lib, mesh
        self.h = _vp()
        lib.call("fvmgpu_flow_create", C.byref(self.h), mesh.h)

    def _len(self, field):
        base = self.mesh.n_faces if field in _FLOW_FACE else self.mesh.n_total
        return base * _FLOW_WIDTH.get(field, 1)

    def set_field(self, field, values):
        a = _f64(values).reshape(-1)
        self.lib.call("fvmgpu_flow_set_field", self.h, int(field), a, a.size)

    def fill_field(self, field, value):
        self.lib.call("fvmgpu_flow_fill_field", self.h, int(field), float(value))

    def get_field(self, field):
        out = np.zeros(self._len(field))
        self.lib.call("fvmgpu_flow_get_field", self.h, int(field), out, out.size)
        w = _FLOW_WIDTH.get(field, 1)
        return out.reshape(-1, w) if w > 1 else out

    def set_bc(self, group_id, kind, params):
        p = _f64(list(params) + [0.0] * max(0, 4 - len(params)))
        self.lib.call("fvmgpu_flow_set_bc", self.h, int(group_id), int(kind), p, len(p))

    def set_reference_cell(self, local_cell):
        """Mesh parts only: local index of the globally lowest cell on the rank that owns it, -1 elsewhere."""
        self.lib.call("fvmgpu_flow_set_reference_cell", self.h, int(local_cell))

    def init(self):
        self.lib.call("fvmgpu_flow_init", self.h)

    @staticmethod
    def opts(momentumURF=0.7, pressureURF=0.3, transient=0, time_order=1, dt=0.1, correctVelocity=1,
             operatingPressure=101325.0, operatingTemperature=300.0, molecularWeight=28.966, incompressible=1):
        return FlowOpts(momentumURF, pressureURF, int(transient), int(time_order), dt, int(correctVelocity),
                        float(operatingPressure), float(operatingTemperature), float(molecularWeight),
                        int(incompressible))

    def assemble_momentum(self, o):
        self.lib.call("fvmgpu_flow_assemble_momentum", self.h, C.byref(o))

    def download_momentum(self):
        nt, nnz = self.mesh.n_total, self.mesh.nnz
        d, off, b = np.zeros(3 * nt), np.zeros(max(nnz, 1)), np.zeros(3 * nt)
        self.lib.call("fvmgpu_flow_download_momentum", self.h, d, off, b)
        return dict(diag=d.reshape(-1, 3), offdiag=off[:nnz], b=b.reshape(-1, 3))

    def solve_momentum(self, amg, bcgstab=None):
        """bcgstab: None or (nMaxIterations, relTol, absTol) of a BCGStab wrapped around `amg`."""
        r0, it = np.zeros(3), np.zeros(3, np.int32)
        b = bcgstab or (0, 0.0, 0.0)
        self.lib.call("fvmgpu_flow_solve_momentum", self.h, amg.h, (b[3] if len(b) > 3 else 1) if bcgstab else 0, int(b[0]), float(b[1]),
                      float(b[2]), r0, it)
        return r0, it

    def assemble_continuity(self, o):
        self.lib.call("fvmgpu_flow_assemble_continuity", self.h, C.byref(o))

    def download_continuity(self):
        nt, nnz = self.mesh.n_total, self.mesh.nnz
        d, off, b, isb = np.zeros(nt), np.zeros(max(nnz, 1)), np.zeros(nt), np.zeros(nt, np.int32)
        self.lib.call("fvmgpu_flow_download_continuity", self.h, d, off, b, isb.ctypes.data_as(C.c_void_p))
        return dict(diag=d, offdiag=off[:nnz], b=b, is_boundary=isb)

    def solve_continuity(self, amg, o, bcgstab=None):
        r0, it = C.c_double(0), C.c_int(0)
        b = bcgstab or (0, 0.0, 0.0)
        self.lib.call("fvmgpu_flow_solve_continuity", self.h, amg.h, (b[3] if len(b) > 3 else 1) if bcgstab else 0, int(b[0]), float(b[1]),
                      float(b[2]), C.byref(o), C.byref(r0), C.byref(it))
        return r0.value, it.value

    def close(self):
        if self.h:
            self.lib.call("fvmgpu_flow_destroy", self.h)
            self.h = None


_default = None


def init_comm_from_torch(lib):
    """One NCCL communicator for the library, bootstrapped through the process group torchrun set
    up: rank 0 creates the unique id, torch.distributed broadcasts its 128 bytes."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(lib.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    lib.comm_init(world, rank, bytes(buf.cpu().numpy().tobytes()))


def default_lib():
    """The product library (fvm_b200/libfvmgpu.so), initialised on device LOCAL_RANK (or 0)."""
    global _default
    if _default is None:
        lib = Lib()
        lib.init(int(os.environ.get("LOCAL_RANK", "0")))
        _default = lib
    return _default


class DeviceMesh:
    """Device mirror of Mesh + StorageSite + CRConnectivity + GeomFields for one mesh."""

    def __init__(self, lib, dim, n_self, n_total, face_cells, cc_row, cc_col, group_offset, group_count,
                 group_id, group_kind):
        self.lib = lib
        self.dim, self.n_self, self.n_total = int(dim), int(n_self), int(n_total)
        fc = _i32(face_cells).reshape(-1)
        self.n_faces = len(fc) // 2
        self.n_interior_faces = int(group_count[0])
        self.nnz = int(cc_row[-1])
        self.group_offset = _i32(group_offset)
        self.group_count = _i32(group_count)
        self.group_id = _i32(group_id)
        self.group_kind = _i32(group_kind)
        self.h = _vp()
        lib.call("fvmgpu_mesh_create", C.byref(self.h), self.dim, self.n_self, self.n_total, self.n_faces,
                 fc, _i32(cc_row), _i32(cc_col), len(self.group_offset), self.group_offset,
                 self.group_count, self.group_id, self.group_kind)

    def set_geometry(self, face_area, face_area_mag, cell_centroid, cell_volume, face_centroid=None,
                     ib_type=None):
        _a, fcp = _nullable(face_centroid, np.float64)
        _b, ibp = _nullable(ib_type, np.int32)
        self.lib.call("fvmgpu_mesh_set_geometry", self.h, _f64(face_area).reshape(-1), _f64(face_area_mag),
                      fcp, _f64(cell_centroid).reshape(-1), _f64(cell_volume), ibp)

    def compute_geometry(self, nodes, face_node_count, face_nodes):
        """MeshMetricsCalculator on the device; returns the GeomFields arrays (host copies)."""
        nodes = _f64(nodes).reshape(-1)
        off = np.zeros(self.n_faces + 1, np.int32)
        np.cumsum(_i32(face_node_count), out=off[1:])
        fa, fam, fc = np.zeros((self.n_faces, 3)), np.zeros(self.n_faces), np.zeros((self.n_faces, 3))
        cc, cv = np.zeros((self.n_total, 3)), np.zeros(self.n_total)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.lib.call("fvmgpu_mesh_compute_geometry", self.h, len(nodes) // 3, nodes, off, _i32(face_nodes).reshape(-1),
                      p(fa), p(fam), p(fc), p(cc), p(cv))
        return dict(face_area=fa, face_area_mag=fam, face_centroid=fc, cell_centroid=cc, cell_volume=cv)

    def set_halo(self, peers, scatter_off, scatter_idx, gather_off, gather_idx):
        """StorageSite scatter/gather maps per neighbour rank (F/StorageSite.h:58-84)."""
        z = np.zeros(1, np.int32)
        self.lib.call("fvmgpu_mesh_set_halo", self.h, len(peers), _i32(peers) if len(peers) else z,
                      _i32(scatter_off), _i32(scatter_idx) if len(scatter_idx) else z, _i32(gather_off),
                      _i32(gather_idx) if len(gather_idx) else z)

    def pair_to_col(self):
        out = np.zeros(2 * self.n_faces, np.int32)
        self.lib.call("fvmgpu_mesh_download_pair_to_col", self.h, out)
        return out.reshape(-1, 2)

    def gradient_weights(self):
        out = np.zeros(3 * self.nnz)
        self.lib.call("fvmgpu_mesh_download_gradient_weights", self.h, out)
        return out.reshape(-1, 3)

    def close(self):
        if self.h:
            self.lib.call("fvmgpu_mesh_destroy", self.h)
            self.h = None


class DeviceSystem:
    """Device mirror of LinearSystem + CRMatrix<T,T,T> (+ boundary flux rows) on a DeviceMesh."""

    def __init__(self, lib, mesh=None, raw=None):
        self.lib = lib
        self.mesh = mesh
        self.h = _vp()
        if mesh is not None:
            lib.call("fvmgpu_system_create", C.byref(self.h), mesh.h)
            self.n_self, self.n_total, self.nnz = mesh.n_self, mesh.n_total, mesh.nnz
        else:
            n_self, n_ghost, row, col, diag, off, b = raw
            self.n_self, self.n_total = int(n_self), int(n_self + n_ghost)
            self.nnz = int(row[-1])
            lib.call("fvmgpu_system_create_raw", C.byref(self.h), int(n_self), int(n_ghost), _i32(row),
                     _i32(col) if self.nnz else np.zeros(1, np.int32), _f64(diag),
                     _f64(off) if self.nnz else np.zeros(1), _f64(b))

    def set_field(self, field, values):
        v = _f64(values).reshape(-1)
        self.lib.call("fvmgpu_system_set_field", self.h, field, v, v.size)

    def fill_field(self, field, value):
        self.lib.call("fvmgpu_system_fill_field", self.h, field, float(value))

    def get_field(self, field, out=None):
        """`out`: a C-contiguous float64 array of the right size to receive the values in place (no temporary)"""
        if field == FIELD_GRADIENT:
            n = 3 * self.n_total
        elif field in (FIELD_FACE_FLUX, FIELD_BFLUX):
            n = self.mesh.n_faces
        elif field == FIELD_BFLUX_BOUNDARY:
            n = self.mesh.n_faces - self.mesh.n_interior_faces
        else:
            n = self.n_total
        if out is None or out.dtype != np.float64 or not out.flags.c_contiguous or out.size != n:
            res = np.empty(n)
            self.lib.call("fvmgpu_system_get_field", self.h, field, res, n)
            if out is not None:
                out.reshape(-1)[:] = res
                return out
            return res.reshape(-1, 3) if field == FIELD_GRADIENT else res
        self.lib.call("fvmgpu_system_get_field", self.h, field, out.reshape(-1), n)
        return out

    def set_bc(self, group_id, kind, params=(), per_face=None):
        p = _f64(list(params) + [0.0] * (4 - len(params)))
        _a, pf = _nullable(per_face, np.float64)
        self.lib.call("fvmgpu_system_set_bc", self.h, int(group_id), int(kind), p, 4, pf)

    def compute_gradient(self):
        self.lib.call("fvmgpu_compute_gradient", self.h)

    def assemble(self, diffusion=1, convection=0, source=1, time_order=0, dt=0.0, underrelax=0.0,
                 apply_bcs=1, eliminate_boundary=1, interface_thickness=0.0):
        o = AssembleOpts(diffusion, convection, source, time_order, dt, underrelax, apply_bcs,
                         eliminate_boundary, interface_thickness)
        self.lib.call("fvmgpu_assemble", self.h, C.byref(o))

    def download(self):
        diag = np.zeros(self.n_total)
        off = np.zeros(max(self.nnz, 1))
        b = np.zeros(self.n_total)
        isb = np.zeros(self.n_total, np.int32)
        self.lib.call("fvmgpu_download_system", self.h, diag, off, b, isb.ctypes.data_as(C.c_void_p))
        return dict(diag=diag, offdiag=off[:self.nnz], b=b, is_boundary=isb)

    def post_solve_update(self):
        self.lib.call("fvmgpu_post_solve_update", self.h)

    def electric_field(self, want=True):
        """updateElectricField: E = -grad(x) for every cell -> [n_total, 3]"""
        out = np.zeros(3 * self.n_total) if want else None
        self.lib.call("fvmgpu_electric_field", self.h, out.ctypes.data_as(C.c_void_p) if want else None)
        return out.reshape(-1, 3) if want else None

    def drift_flux_into(self, charge_system, mobility, vsat, symmetry_group_ids=(), want_velocity=True):
        """updateElectronVelocity + updateConvectionFlux; the face flux goes to charge_system's FACE_FLUX"""
        ids = _i32(list(symmetry_group_ids) or [0])
        out = np.zeros(3 * self.n_total) if want_velocity else None
        self.lib.call("fvmgpu_electric_drift_flux", self.h, charge_system.h, float(mobility), float(vsat),
                      len(symmetry_group_ids), ids, out.ctypes.data_as(C.c_void_p) if want_velocity else None)
        return out.reshape(-1, 3) if want_velocity else None

    def halo_exchange(self, field):
        self.lib.call("fvmgpu_system_halo_exchange", self.h, int(field))

    def close(self):
        if self.h:
            self.lib.call("fvmgpu_system_destroy", self.h)
            self.h = None


class DeviceAMG:
    """Device AMG hierarchy + cycle driver; also the preconditioner of BCGStab."""

    def __init__(self, lib, opts=None):
        self.lib = lib
        self.opts = opts or lib.default_amg_opts()
        self.h = _vp()
        lib.call("fvmgpu_amg_create", C.byref(self.h), C.byref(self.opts))

    def set_opts(self, opts):
        self.opts = opts
        self.lib.call("fvmgpu_amg_set_opts", self.h, C.byref(opts))

    def solve(self, system):
        r0, r, it = C.c_double(0), C.c_double(0), C.c_int(0)
        self.lib.call("fvmgpu_amg_solve", self.h, system.h, C.byref(r0), C.byref(r), C.byref(it))
        return r0.value, r.value, it.value

    def smooth(self, system):
        self.lib.call("fvmgpu_amg_smooth", self.h, system.h)

    def bcgstab(self, system, n_max_iterations, relative_tolerance, absolute_tolerance):
        r0, r, it = C.c_double(0), C.c_double(0), C.c_int(0)
        self.lib.call("fvmgpu_bcgstab_solve", self.h, system.h, int(n_max_iterations),
                      float(relative_tolerance), float(absolute_tolerance), C.byref(r0), C.byref(r),
                      C.byref(it))
        return r0.value, r.value, it.value

    def cg(self, system, n_max_iterations, relative_tolerance, absolute_tolerance):
        """CG preconditioned by one cycle of this AMG (F/CG.cpp:24-140)"""
        r0, r, it = C.c_double(0), C.c_double(0), C.c_int(0)
        self.lib.call("fvmgpu_cg_solve", self.h, system.h, int(n_max_iterations), float(relative_tolerance),
                      float(absolute_tolerance), C.byref(r0), C.byref(r), C.byref(it))
        return r0.value, r.value, it.value

    def jacobi(self, system, n_max_iterations, relative_tolerance, absolute_tolerance):
        """JacobiSolver::solve (F/JacobiSolver.cpp:46-95)"""
        r0, r, it = C.c_double(0), C.c_double(0), C.c_int(0)
        self.lib.call("fvmgpu_jacobi_solve", self.h, system.h, int(n_max_iterations), float(relative_tolerance),
                      float(absolute_tolerance), C.byref(r0), C.byref(r), C.byref(it))
        return r0.value, r.value, it.value

    def bcgstab_ilu0(self, ds, n_max_iterations, relative_tolerance, absolute_tolerance=1e-50):
        r0, r, it = C.c_double(0), C.c_double(0), C.c_int(0)
        self.lib.call("fvmgpu_bcgstab_ilu0_solve", self.h, ds.h, int(n_max_iterations), float(relative_tolerance),
                      float(absolute_tolerance), C.byref(r0), C.byref(r), C.byref(it))
        return r0.value, r.value, it.value

    def ilu0(self, ds, n_max_iterations, relative_tolerance, absolute_tolerance=1e-50):
        r0, r, it, lv = C.c_double(0), C.c_double(0), C.c_int(0), C.c_int(0)
        self.lib.call("fvmgpu_ilu0_solve", self.h, ds.h, int(n_max_iterations), float(relative_tolerance),
                      float(absolute_tolerance), C.byref(r0), C.byref(r), C.byref(it), C.byref(lv))
        self.ilu_levels = lv.value
        return r0.value, r.value, it.value

    def levels(self):
        n = C.c_int(0)
        sizes = np.zeros(64, np.int64)
        nnzs = np.zeros(64, np.int64)
        cols = np.zeros(64, np.int32)
        self.lib.call("fvmgpu_amg_levels", self.h, 64, C.byref(n), sizes, nnzs, cols)
        k = n.value
        cb = np.zeros(64)
        self.lib.call("fvmgpu_amg_level_col_bytes", self.h, 64, cb.ctypes.data_as(C.POINTER(C.c_double)))
        return dict(sizes=sizes[:k].tolist(), nnz=nnzs[:k].tolist(), colours=cols[:k].tolist(), col_bytes=cb[:k].tolist())

    def level_order(self, level=0):
        """(nat, colourStart): nat[r] = source row of level-row r; colour class c = level-rows
        colourStart[c]:colourStart[c+1]."""
        n = self.levels()["sizes"][level]
        nat = np.zeros(max(n, 1), np.int32)
        nc = C.c_int(0)
        cs = np.zeros(65, np.int64)
        self.lib.call("fvmgpu_amg_level_order", self.h, level, n, nat, C.byref(nc), cs)
        return nat[:n], cs[:nc.value + 1].copy()

    def last_timing(self):
        a, b = C.c_double(0), C.c_double(0)
        self.lib.call("fvmgpu_amg_last_timing", self.h, C.byref(a), C.byref(b))
        return dict(setup_ms=a.value, cycles_ms=b.value)

    def history(self):
        n = C.c_int(0)
        out = np.zeros(1 << 14)
        self.lib.call("fvmgpu_solver_history", self.h, len(out), out, C.byref(n))
        return out[:min(n.value, len(out))].copy()

    def cleanup(self):
        self.lib.call("fvmgpu_amg_cleanup", self.h)

    def close(self):
        if self.h:
            self.lib.call("fvmgpu_amg_destroy", self.h)
            self.h = None
