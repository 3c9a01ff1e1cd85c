"""Host-side mirror of the reference's Python model API for the accelerated path.

The reference exposes `ThermalModelA`, `AMG`, `BCGStab`, BC/VC/option dictionaries through SWIG
(F/ThermalModel.i:22-48, F/AMG.i:1-34, F/FloatVarDict.i:22-55); scripts drive them as in
T/THERMAL_MATRIX/testThermalParallel.py. The classes below keep those names, fields, defaults and
call order, and run every numerical step of `advance()` on the GPU through the C ABI
(include/fvmgpu.h). Host numpy arrays stay the source of truth at the API boundary, like the
reference's `Field[site].asNumPyArray()` views: `advance()` uploads the model's input fields,
runs assembly + solve + update on the device and copies the solution back.

No numerical work happens in Python and there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import capi
from . import meshgen


class CException(RuntimeError):
    """Same name as the reference's exception type (F/CException.h:16-21)."""


# ----------------------------------------------------------------------------- sites / mesh
class StorageSite:
    """F/StorageSite.h:18-112 (counts only; scatter/gather maps live on the Mesh)."""

    def __init__(self, self_count, n_ghost=0, offset=0):
        self._self, self._count, self._offset = int(self_count), int(self_count + n_ghost), int(offset)

    def getCount(self):
        return self._count

    def getSelfCount(self):
        return self._self

    def getOffset(self):
        return self._offset


class FaceGroup:
    """F/Mesh.h:28-43"""

    def __init__(self, count, offset, gid, group_type):
        self.site = StorageSite(count, 0, offset)
        self.id = int(gid)
        self.groupType = group_type


class Mesh:
    """Host mesh: what the importers / partitioner hand to the models (F/Mesh.h), built here from
    the raw constructor arrays (F/Mesh.h:93-99). Geometry comes from MeshMetricsCalculatorA."""

    _last_id = 0

    def __init__(self, raw, group_types=None):
        if group_types is None and "group_types" in raw:   # a partitioned mesh (fvm_b200.partition)
            group_types = raw.group_types
        self.raw = raw
        self.dim = raw.dim
        self._id = Mesh._last_id
        Mesh._last_id += 1
        self._cells = StorageSite(raw.n_cells, raw.n_total - raw.n_cells)
        self._faces = StorageSite(raw.n_faces)
        self.cc_row, self.cc_col = meshgen.connectivity(raw)
        self._groups = []
        for g in range(len(raw.group_offset)):
            gt = "interior" if g == 0 else (group_types[g] if group_types else "wall")
            self._groups.append(FaceGroup(raw.group_count[g], raw.group_offset[g], raw.group_id[g], gt))
        self.device = None  # capi.DeviceMesh, created by MeshMetricsCalculatorA.init()

    def getID(self):
        return self._id

    def getDimension(self):
        return self.dim

    def getCells(self):
        return self._cells

    def getFaces(self):
        return self._faces

    def getAllFaceGroups(self):
        return list(self._groups)

    def getBoundaryFaceGroups(self):
        return [g for g in self._groups if g.groupType not in ("interior", "interface")]

    def group_kinds(self):
        kinds = []
        for g in self._groups:
            kinds.append({"interior": capi.GROUP_INTERIOR, "interface": capi.GROUP_INTERFACE,
                          "symmetry": capi.GROUP_SYMMETRY,
                          "dielectric interface": capi.GROUP_DIELECTRIC_INTERFACE}.get(g.groupType, capi.GROUP_BOUNDARY))
        return np.array(kinds, np.int32)


class Field(dict):
    """name + {site: ndarray} (F/Field.h); `field[site]` is the host array itself."""

    def __init__(self, name):
        super().__init__()
        self.name = name


class GeomFields:
    """F/GeomFields.h"""

    def __init__(self, base_name):
        for n in ("coordinate", "area", "areaMag", "volume", "ibType"):
            setattr(self, n, Field(base_name + "." + n))


class MeshMetricsCalculatorA:
    """F/MeshMetricsCalculator.h:31-34; init() fills GeomFields and uploads the mesh to the device."""

    def __init__(self, geom_fields, meshes, lib=None):
        self.geom, self.meshes, self.lib = geom_fields, meshes, lib

    def init(self):
        lib = self.lib or capi.default_lib()
        for m in self.meshes:
            cells, faces = m.getCells(), m.getFaces()
            raw = m.raw
            dm = capi.DeviceMesh(lib, m.dim, raw.n_cells, raw.n_total, raw.face_cells, m.cc_row, m.cc_col,
                                 raw.group_offset, raw.group_count, raw.group_id, m.group_kinds())
            ib = np.full(cells.getCount(), -1, np.int32)  # IBTYPE_FLUID
            if "geometry" in raw:
                # a partitioned mesh carries its geometry (interface ghosts = the remote cells' metrics)
                mt = raw.geometry
                dm.set_geometry(mt["face_area"], mt["face_area_mag"], mt["cell_centroid"], mt["cell_volume"],
                                face_centroid=mt["face_centroid"], ib_type=ib)
            else:
                # face areas / centroids, cell centroids / volumes computed on the device from the nodes
                mt = dm.compute_geometry(raw.nodes, raw.face_node_count, raw.face_nodes)
            self.geom.area[faces] = mt["face_area"]
            self.geom.areaMag[faces] = mt["face_area_mag"]
            self.geom.coordinate[faces] = mt["face_centroid"]
            self.geom.coordinate[cells] = mt["cell_centroid"]
            self.geom.volume[cells] = mt["cell_volume"]
            self.geom.ibType[cells] = ib
            if "halo" in raw:  # StorageSite scatter/gather maps of a partitioned mesh
                h = raw.halo
                dm.set_halo(h["peers"], h["scatter_off"], h["scatter_idx"], h["gather_off"], h["gather_idx"])
            m.device = dm


def upload_mesh(lib, m, geom):
    """Create the device mirror of mesh `m` from host connectivity + GeomFields arrays."""
    raw = m.raw
    cells, faces = m.getCells(), m.getFaces()
    dm = capi.DeviceMesh(lib, m.dim, raw.n_cells, raw.n_total, raw.face_cells, m.cc_row, m.cc_col,
                         raw.group_offset, raw.group_count, raw.group_id, m.group_kinds())
    dm.set_geometry(geom.area[faces], geom.areaMag[faces], geom.coordinate[cells], geom.volume[cells],
                    face_centroid=geom.coordinate[faces], ib_type=geom.ibType[cells])
    if "halo" in raw:  # StorageSite scatter/gather maps of a partitioned mesh
        h = raw.halo
        dm.set_halo(h["peers"], h["scatter_off"], h["scatter_idx"], h["gather_off"], h["gather_idx"])
    m.device = dm
    return dm


# ----------------------------------------------------------------------------- dictionaries
class FloatVarDict(dict):
    """F/FloatVarDict.h:44-98 + the SWIG helpers setVar/getVar (F/FloatVarDict.i:22-55).
    A value is a float or a per-site numpy array (the reference's Field-valued FloatVal)."""

    def defineVar(self, name, default):
        dict.__setitem__(self, name, default)

    def setVar(self, name, value):
        if name not in self:
            raise CException("uknown var " + name)  # (sic) F/FloatVarDict.h:66
        dict.__setitem__(self, name, value)

    def getVar(self, name):
        if name not in self:
            raise CException("uknown var " + name)
        return dict.__getitem__(self, name)

    def __getitem__(self, name):
        return self.getVar(name)

    def __setitem__(self, name, value):
        self.setVar(name, value)


class ThermalBC(FloatVarDict):
    """F/ThermalBC.h:8-21"""

    def __init__(self):
        super().__init__()
        self.defineVar("specifiedTemperature", 300.0)
        self.defineVar("specifiedHeatFlux", 0.0)
        self.defineVar("convectiveCoefficient", 0.0)
        self.defineVar("farFieldTemperature", 300.0)
        self.defineVar("surfaceEmissivity", 1.0)
        self.bcType = ""


class ThermalVC(FloatVarDict):
    """F/ThermalBC.h:23-34"""

    def __init__(self):
        super().__init__()
        self.defineVar("thermalConductivity", 1.0)
        self.defineVar("density", 1.0)
        self.defineVar("specificHeat", 1.0)
        self.vcType = ""


class ThermalModelOptions(FloatVarDict):
    """F/ThermalBC.h:36-69"""

    def __init__(self):
        super().__init__()
        self.defineVar("initialTemperature", 300.0)
        self.defineVar("timeStep", 1e-7)
        self.relativeTolerance = 1e-8
        self.absoluteTolerance = 1e-16
        self.linearSolver = None
        self.useCentralDifference = False
        self.transient = False
        self.timeDiscretizationOrder = 1

    def getLinearSolver(self):
        if self.linearSolver is None:  # F/ThermalBC.h:56-67
            ls = AMG()
            ls.relativeTolerance = 1e-1
            ls.nMaxIterations = 20
            ls.verbosity = 0
            self.linearSolver = ls
        return self.linearSolver


class ThermalFields:
    """F/ThermalFields.h"""

    def __init__(self, base_name):
        for n in ("temperature", "temperatureN1", "temperatureN2", "specificHeat", "conductivity",
                  "heatFlux", "temperatureGradient", "convectionFlux", "source", "zero", "one"):
            setattr(self, n, Field(base_name + "." + n))


# ----------------------------------------------------------------------------- solvers
class LinearSolver:
    """F/LinearSolver.h:11-31"""

    def __init__(self):
        self.nMaxIterations = 100
        self.verbosity = 2
        self.relativeTolerance = 1e-8
        self.absoluteTolerance = 1e-50


class AMG(LinearSolver):
    """F/AMG.h:24-112; tunables and defaults F/AMG.cpp:14-22. `solve` runs entirely on the GPU."""

    V_CYCLE, W_CYCLE, F_CYCLE = 0, 1, 2
    GAUSS_SEIDEL, JACOBI = 0, 1

    def __init__(self):
        super().__init__()
        self.maxCoarseLevels = 30
        self.nPreSweeps = 0
        self.nPostSweeps = 1
        self.coarseGroupSize = 2
        self.weightRatioThreshold = 0.65
        self.cycleType = AMG.V_CYCLE
        self.smootherType = AMG.GAUSS_SEIDEL
        self.scaleCorrections = True
        self._dev = None
        self._lib = None
        self._totalIterations = 0
        self.lastIterations = 0
        self.lastHistory = None

    def _opts(self):
        return capi.AmgOpts(self.nMaxIterations, self.verbosity, self.relativeTolerance,
                            self.absoluteTolerance, self.maxCoarseLevels, self.nPreSweeps,
                            self.nPostSweeps, self.coarseGroupSize, self.weightRatioThreshold,
                            self.cycleType, self.smootherType)

    def _device(self, lib):
        if self._dev is None or self._lib is not lib:
            self._dev = capi.DeviceAMG(lib, self._opts())
            self._lib = lib
        else:
            self._dev.set_opts(self._opts())
        return self._dev

    def solve(self, ls):
        """LinearSolver::solve: returns the INITIAL residual 1-norm (F/AMG.cpp:235,281)."""
        dev = self._device(ls.lib)
        r0, r, it = dev.solve(ls)
        self._totalIterations += it
        self.lastIterations = it
        self.lastResidual = r
        if self.verbosity > 0:  # parallel-build print format, F/AMG.cpp:239-271
            self.lastHistory = dev.history()
            print("0: [%s : %g]" % (ls.field_name, r0))
            if it > 0:
                print("%d: [%s : %g]" % (it, ls.field_name, r))
        return r0

    def smooth(self, ls):
        self._device(ls.lib).smooth(ls)

    def cleanup(self):
        if self._dev is not None:
            self._dev.cleanup()

    def getTotalIterations(self):
        return self._totalIterations

    def levels(self):
        return self._dev.levels() if self._dev is not None else None


class BCGStab(LinearSolver):
    """F/BCGStab.h; `preconditioner` is an AMG (one cycle per application, F/BCGStab.cpp:85-89) or an
    ILU0Solver (one ILU(0) solve per application, T/PARALLEL_CAVITY_ILU0)."""

    def __init__(self):
        super().__init__()
        self.preconditioner = None
        self._totalIterations = 0
        self.lastIterations = 0

    def solve(self, ls):
        if self.preconditioner is None:
            raise CException("BCGStab: no preconditioner set")
        dev = self.preconditioner._device(ls.lib)
        if isinstance(self.preconditioner, ILU0Solver):
            r0, r, it = dev.bcgstab_ilu0(ls, self.nMaxIterations, self.relativeTolerance, self.absoluteTolerance)
        else:
            r0, r, it = dev.bcgstab(ls, self.nMaxIterations, self.relativeTolerance, self.absoluteTolerance)
        self._totalIterations += it
        self.lastIterations = it
        self.lastResidual = r
        if self.verbosity > 0:
            print("0: [%s : %g]" % (ls.field_name, r0))
            print("%d: [%s : %g]" % (it, ls.field_name, r))
        return r0

    def smooth(self, ls):
        raise CException("cannot use BCGStab as preconditioner")  # F/BCGStab.cpp:172-176

    def cleanup(self):
        if self.preconditioner is not None:
            self.preconditioner.cleanup()

    def getTotalIterations(self):
        return self._totalIterations


class CG(LinearSolver):
    """F/CG.h, F/CG.cpp:24-140: conjugate gradients, `preconditioner` must be an AMG (one cycle per
    application). For symmetric systems (pure diffusion)."""

    def __init__(self):
        super().__init__()
        self.preconditioner = None
        self._totalIterations = 0
        self.lastIterations = 0

    def solve(self, ls):
        if self.preconditioner is None:
            raise CException("CG: no preconditioner set")
        dev = self.preconditioner._device(ls.lib)
        r0, r, it = dev.cg(ls, self.nMaxIterations, self.relativeTolerance, self.absoluteTolerance)
        self._totalIterations += it
        self.lastIterations = it
        self.lastResidual = r
        if self.verbosity > 0:
            print("0: [%s : %g]" % (ls.field_name, r0))
            print("%d: [%s : %g]" % (it, ls.field_name, r))
        return r0

    def smooth(self, ls):
        raise CException("cannot use CG as preconditioner")  # F/CG.cpp:142-146

    def cleanup(self):
        if self.preconditioner is not None:
            self.preconditioner.cleanup()

    def getTotalIterations(self):
        return self._totalIterations


class JacobiSolver(LinearSolver):
    """F/JacobiSolver.h, F/JacobiSolver.cpp:46-95: plain Jacobi iterations on the finest level."""

    def __init__(self):
        super().__init__()
        self._amg = AMG()
        self.lastIterations = 0

    def solve(self, ls):
        dev = self._amg._device(ls.lib)
        r0, r, it = dev.jacobi(ls, self.nMaxIterations, self.relativeTolerance, self.absoluteTolerance)
        self.lastIterations = it
        self.lastResidual = r
        if self.verbosity > 0:
            print("0: [%s : %g]" % (ls.field_name, r0))
            print("%d: [%s : %g]" % (it, ls.field_name, r))
        return r0

    def smooth(self, ls):
        raise CException("JacobiSolver.smooth: use AMG with smootherType = JACOBI as a preconditioner")

    def cleanup(self):
        self._amg.cleanup()


class ILU0Solver(LinearSolver):
    """F/ILU0Solver.h, F/ILU0Solver.cpp:46-93: delta = U^-1 L^-1 (-b) per sweep with the reference's ILU(0)
    (F/CRMatrix.h:1546-1715, factors bit-identical to the reference's). On its own every sweep recomputes the
    same delta -- exactly as in the reference -- so it is meant as BCGStab's preconditioner."""

    def __init__(self):
        super().__init__()
        self._amg = AMG()
        self.lastIterations = 0

    def _device(self, lib):
        return self._amg._device(lib)

    def solve(self, ls):
        dev = self._device(ls.lib)
        r0, r, it = dev.ilu0(ls, self.nMaxIterations, self.relativeTolerance, self.absoluteTolerance)
        self.lastIterations = it
        self.lastResidual = r
        if self.verbosity > 0:
            print("0: [%s : %g]" % (ls.field_name, r0))
            print("%d: [%s : %g]" % (it, ls.field_name, r))
        return r0

    def smooth(self, ls):
        self._device(ls.lib).ilu0(ls, 2, 0.0, 0.0)   # one sweep

    def cleanup(self):
        self._amg.cleanup()


class LinearSystem(capi.DeviceSystem):
    """Device LinearSystem (F/LinearSystem.h:11-64) tagged with the field name used in prints."""

    def __init__(self, lib, mesh=None, raw=None, field_name="x"):
        super().__init__(lib, mesh=mesh, raw=raw)
        self.field_name = field_name


# ----------------------------------------------------------------------------- ThermalModel
_BC_KINDS = {
    "SpecifiedTemperature": capi.BC_DIRICHLET,
    "SpecifiedHeatFlux": capi.BC_NEUMANN,
    "Symmetry": capi.BC_NEUMANN,
    "Convective": capi.BC_CONVECTIVE,
    "Radiative": capi.BC_RADIATIVE,
    "Mixed": capi.BC_MIXED,
}


class ThermalModelA:
    """ThermalModel<double> (F/ThermalModel.h:28-52, Impl in F/ThermalModel_impl.h).

    advance(niter) performs, per outer iteration and all on the device: gradient, the fused
    assembly (diffusion + convection + source [+ time derivative] + BCs + boundary elimination),
    the linear solve, postSolve/updateSolution -- the statement sequence of Impl::advance
    (F/ThermalModel_impl.h:424-456)."""

    def __init__(self, geom_fields, thermal_fields, meshes, lib=None):
        self.geom, self.fields, self.meshes = geom_fields, thermal_fields, meshes
        self.lib = lib
        self._bcMap, self._vcMap = {}, {}
        self._options = ThermalModelOptions()
        self._initialNorm = None
        self._niters = 0
        self._systems = {}
        self.timings = []
        for mesh in meshes:  # F/ThermalModel_impl.h:52-82
            vc = ThermalVC()
            vc.vcType = "flow"
            self._vcMap[mesh.getID()] = vc
            for fg in mesh.getBoundaryFaceGroups():
                bc = ThermalBC()
                self._bcMap[fg.id] = bc
                if fg.groupType in ("wall", "symmetry"):
                    bc.bcType = "SpecifiedHeatFlux"
                elif fg.groupType in ("velocity-inlet", "pressure-outlet"):
                    bc.bcType = "SpecifiedTemperature"
                else:
                    raise CException("ThermalModel: unknown face group type " + fg.groupType)

    def getBCMap(self):
        return self._bcMap

    def getVCMap(self):
        return self._vcMap

    def getBC(self, gid):
        return self._bcMap[gid]

    def getOptions(self):
        return self._options

    def init(self):  # F/ThermalModel_impl.h:84-172
        f, o = self.fields, self._options
        for mesh in self.meshes:
            cells, faces = mesh.getCells(), mesh.getFaces()
            n = cells.getCount()
            vc = self._vcMap[mesh.getID()]
            if mesh.device is None:
                raise CException("ThermalModel.init: mesh metrics not initialised (MeshMetricsCalculatorA.init)")
            hostlib = self.lib or mesh.device.lib
            # the fields that travel every advance() live in page-locked memory (the reference's Array<T> storage
            # allocated with fvmgpu_host_alloc): scripts see ordinary numpy arrays
            f.temperature[cells] = hostlib.pinned_full(n, float(o["initialTemperature"]))
            if o.transient:
                f.temperatureN1[cells] = f.temperature[cells].copy()
                if o.timeDiscretizationOrder > 1:
                    f.temperatureN2[cells] = f.temperature[cells].copy()
            f.conductivity[cells] = hostlib.pinned_full(n, float(vc["thermalConductivity"]))
            f.source[cells] = hostlib.pinned_full(n, 0.0)
            f.specificHeat[cells] = np.full(n, float(vc["density"]) * float(vc["specificHeat"]))
            f.temperatureGradient[cells] = np.zeros((n, 3))
            f.convectionFlux[faces] = np.zeros(faces.getCount())
            for fg in mesh.getBoundaryFaceGroups():
                f.heatFlux[fg.site] = np.zeros(fg.site.getCount())
            if mesh.device is None:
                raise CException("ThermalModel.init: mesh metrics not initialised (MeshMetricsCalculatorA.init)")
            lib = self.lib or mesh.device.lib
            self._systems[mesh.getID()] = LinearSystem(lib, mesh=mesh.device, field_name=f.temperature.name)
        self._niters = 0
        self._initialNorm = None

    # ---- device plumbing
    def _upload(self, mesh, ls):
        f, o = self.fields, self._options
        cells, faces = mesh.getCells(), mesh.getFaces()
        ls.set_field(capi.FIELD_X, f.temperature[cells])
        ls.set_field(capi.FIELD_DIFFUSIVITY, f.conductivity[cells])
        ls.set_field(capi.FIELD_SOURCE, f.source[cells])
        flux = f.convectionFlux[faces]
        self._convecting = bool(flux.any())   # (no temporary: the face array of a 256^3 mesh has 50 M entries)
        if self._convecting:
            ls.set_field(capi.FIELD_FACE_FLUX, flux)
        if o.transient:
            ls.set_field(capi.FIELD_X_N1, f.temperatureN1[cells])
            ls.set_field(capi.FIELD_DENSITY, f.specificHeat[cells])
            if o.timeDiscretizationOrder > 1:
                ls.set_field(capi.FIELD_X_N2, f.temperatureN2[cells])
        for fg in mesh.getBoundaryFaceGroups():
            bc = self._bcMap[fg.id]
            if bc.bcType not in _BC_KINDS:
                raise CException(bc.bcType + " not implemented for ThermalModel")
            kind = _BC_KINDS[bc.bcType]
            per_face = None

            def val(name):
                v = bc[name]
                return v

            if bc.bcType == "SpecifiedTemperature":
                v = val("specifiedTemperature")
                if self._convecting:
                    kind = capi.BC_DIRICHLET_OR_OUTFLOW
                params = [v]
            elif bc.bcType == "SpecifiedHeatFlux":
                params = [val("specifiedHeatFlux")]
            elif bc.bcType == "Symmetry":
                params = [0.0]
            elif bc.bcType == "Convective":
                params = [val("convectiveCoefficient"), val("farFieldTemperature")]
            elif bc.bcType == "Radiative":
                params = [val("surfaceEmissivity"), val("farFieldTemperature")]
            else:
                params = [val("convectiveCoefficient"), val("surfaceEmissivity"), val("farFieldTemperature")]
            if isinstance(params[0], np.ndarray):
                per_face = params[0]
                params[0] = 0.0
            ls.set_bc(fg.id, kind, [float(p) for p in params], per_face=per_face)

    def _download(self, mesh, ls):
        f = self.fields
        cells = mesh.getCells()
        ls.get_field(capi.FIELD_X, out=f.temperature[cells])          # straight into the field array
        bflux = ls.get_field(capi.FIELD_BFLUX_BOUNDARY)               # boundary faces only
        ni = mesh.getFaces().getCount() - len(bflux)
        for fg in mesh.getBoundaryFaceGroups():
            o = fg.site.getOffset() - ni
            f.heatFlux[fg.site][:] = bflux[o:o + fg.site.getCount()]

    def _assemble(self, ls):
        o = self._options
        ls.assemble(diffusion=1, convection=(2 if o.useCentralDifference else 1) if self._convecting else 0,
                    source=1, time_order=(o.timeDiscretizationOrder if o.transient else 0),
                    dt=float(o["timeStep"]) if o.transient else 0.0, underrelax=0.0, apply_bcs=1,
                    eliminate_boundary=1)

    def advance(self, niter):
        """Impl::advance, F/ThermalModel_impl.h:424-456 (single mesh per system)."""
        o = self._options
        solver = o.getLinearSolver()
        for mesh in self.meshes:
            ls = self._systems[mesh.getID()]
            lib = ls.lib
            self._upload(mesh, ls)
            for _ in range(niter):
                t = {}
                lib.timer_start(1)
                self._assemble(ls)                      # initLinearization+initAssembly+linearize+initSolve
                t["assemble_ms"] = lib.timer_stop(1)
                lib.timer_start(1)
                rnorm = solver.solve(ls)                # LinearSolver::solve
                t["solve_ms"] = lib.timer_stop(1)
                t["linear_iterations"] = solver.lastIterations
                if self._initialNorm is None:
                    self._initialNorm = rnorm
                ratio = rnorm / self._initialNorm if self._initialNorm != 0 else 0.0
                print("%d: [%s : %g]" % (self._niters, self.fields.temperature.name, rnorm))
                solver.cleanup()
                lib.timer_start(1)
                ls.post_solve_update()                  # postSolve + updateSolution
                t["update_ms"] = lib.timer_stop(1)
                t["rnorm"] = rnorm
                self.timings.append(t)
                self._niters += 1
                if rnorm < o.absoluteTolerance or ratio < o.relativeTolerance:
                    break
            self._download(mesh, ls)

    def dumpMatrix(self, file_base):
        """Impl::dumpMatrix, F/ThermalModel_impl.h:499-572 (MatrixMarket + rhs text files)."""
        for idx, mesh in enumerate(self.meshes):
            ls = self._systems[mesh.getID()]
            self._upload(mesh, ls)
            self._assemble(ls)
            d = ls.download()
            n = mesh.getCells().getSelfCount()
            row, col = mesh.cc_row, mesh.cc_col
            lines = []
            for i in range(n):
                lines.append("%d %d %f" % (i + 1, i + 1, d["diag"][i]))
                for jp in range(row[i], row[i + 1]):
                    if col[jp] < n:
                        lines.append("%d %d %f" % (i + 1, col[jp] + 1, d["offdiag"][jp]))
            with open("%s_mesh%d.mat" % (file_base, idx), "w") as fh:
                fh.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (n, n, len(lines)))
                fh.write("\n".join(lines) + "\n")
            with open(file_base + ".rhs", "w") as fh:
                fh.write("".join("%f\n" % (-v) for v in d["b"][:n]))

    def getHeatFluxIntegral(self, mesh, face_group_id):  # F/ThermalModel_impl.h:400-421
        for fg in mesh.getBoundaryFaceGroups():
            if fg.id == face_group_id:
                return float(np.sum(self.fields.heatFlux[fg.site]))
        raise CException("getHeatFluxIntegral: invalid faceGroupID")

    def updateTime(self):  # F/ThermalModel_impl.h:585-612
        f, o = self.fields, self._options
        for mesh in self.meshes:
            cells = mesh.getCells()
            if o.timeDiscretizationOrder > 1:
                f.temperatureN2[cells][:] = f.temperatureN1[cells]
            f.temperatureN1[cells][:] = f.temperature[cells]


# ----------------------------------------------------------------------------- SpeciesModel
class SpeciesBC(FloatVarDict):
    """F/SpeciesBC.h:8-19"""

    def __init__(self):
        super().__init__()
        self.defineVar("specifiedMassFraction", 0.0)
        self.defineVar("specifiedMassFlux", 0.0)
        self.bcType = ""


class SpeciesVC(FloatVarDict):
    """F/SpeciesBC.h:21-30"""

    def __init__(self):
        super().__init__()
        self.defineVar("massDiffusivity", 1e-9)
        self.defineVar("initialMassFraction", 1.0)
        self.vcType = ""


class SpeciesModelOptions(FloatVarDict):
    """F/SpeciesBC.h:32-74"""

    def __init__(self):
        super().__init__()
        for name, v in (("A_coeff", 1.0), ("B_coeff", 0.0), ("ButlerVolmerRRConstant", 5.0e-7),
                        ("ButlerVolmerAnodeShellMeshID", -1), ("ButlerVolmerCathodeShellMeshID", -1),
                        ("timeStep", 0.1), ("interfaceUnderRelax", 1.0)):
            self.defineVar(name, v)
        self.relativeTolerance = 1e-8
        self.absoluteTolerance = 1e-16
        self.linearSolver = None
        self.useCentralDifference = False
        self.transient = False
        self.ButlerVolmer = False
        self.timeDiscretizationOrder = 1

    def getLinearSolver(self):
        if self.linearSolver is None:  # F/SpeciesBC.h:61-72
            ls = AMG()
            ls.relativeTolerance = 1e-1
            ls.nMaxIterations = 20
            ls.verbosity = 0
            self.linearSolver = ls
        return self.linearSolver


class SpeciesFields:
    """F/SpeciesFields.h:11-30 (one set per species)"""

    def __init__(self, base_name):
        for n in ("massFraction", "massFlux", "diffusivity", "source", "convectionFlux", "massFractionN1",
                  "massFractionN2", "zero", "one", "elecPotential", "massFractionElectricModel"):
            setattr(self, n, Field(base_name + "." + n))


class SpeciesModelA:
    """SpeciesModel<double> (F/SpeciesModel.h:17-55, Impl in F/SpeciesModel_impl.h): nSpecies scalar transport
    equations (diffusion + convection + source [+ time derivative]), one after the other per outer iteration -- the
    same device path as ThermalModelA (gradient, fused assembly with BCs and boundary elimination, linear solve,
    postSolve / updateSolution), one LinearSystem per species. The shell-mesh couplings of the reference (interface
    jump, Butler-Volmer: F/SpeciesModel_impl.h:493-540) belong to its battery models and are rejected."""

    def __init__(self, geom_fields, meshes, nSpecies, lib=None):
        self.geom, self.meshes, self.lib, self._nSpecies = geom_fields, list(meshes), lib, int(nSpecies)
        self._options = SpeciesModelOptions()
        self._fields = [SpeciesFields("species") for _ in range(self._nSpecies)]
        self._bcMaps = [dict() for _ in range(self._nSpecies)]
        self._vcMaps = [dict() for _ in range(self._nSpecies)]
        self._initialNorm = [None] * self._nSpecies
        self._currentResidual = [None] * self._nSpecies
        self._niters = 0
        self._systems = {}
        self.timings = []
        for m in range(self._nSpecies):  # F/SpeciesModel_impl.h:55-95
            for mesh in self.meshes:
                vc = SpeciesVC()
                vc.vcType = "flow"
                self._vcMaps[m][mesh.getID()] = vc
                for fg in mesh.getBoundaryFaceGroups():
                    bc = SpeciesBC()
                    self._bcMaps[m][fg.id] = bc
                    if fg.groupType in ("wall", "symmetry"):
                        bc.bcType = "SpecifiedMassFlux"
                    elif fg.groupType in ("velocity-inlet", "pressure-outlet"):
                        bc.bcType = "SpecifiedMassFraction"
                    else:
                        raise CException("SpeciesModel: unknown face group type " + fg.groupType)

    def getSpeciesFields(self, speciesId):
        return self._fields[speciesId]

    def getBCMap(self, speciesId):
        return self._bcMaps[speciesId]

    def getVCMap(self, speciesId):
        return self._vcMaps[speciesId]

    def getBC(self, gid, speciesId):
        return self._bcMaps[speciesId][gid]

    def getOptions(self):
        return self._options

    def init(self):  # F/SpeciesModel_impl.h:97-330
        o = self._options
        if o.ButlerVolmer:
            raise CException("SpeciesModelA: the Butler-Volmer shell coupling is not built")
        for m in range(self._nSpecies):
            f = self._fields[m]
            for mesh in self.meshes:
                if mesh.device is None:
                    raise CException("SpeciesModel.init: mesh metrics not initialised (MeshMetricsCalculatorA.init)")
                cells, faces = mesh.getCells(), mesh.getFaces()
                n = cells.getCount()
                vc = self._vcMaps[m][mesh.getID()]
                f.massFraction[cells] = np.full(n, float(vc["initialMassFraction"]))
                if o.transient:
                    f.massFractionN1[cells] = f.massFraction[cells].copy()
                    if o.timeDiscretizationOrder > 1:
                        f.massFractionN2[cells] = f.massFraction[cells].copy()
                f.diffusivity[cells] = np.full(n, float(vc["massDiffusivity"]))
                f.source[cells] = np.zeros(n)
                f.zero[cells] = np.zeros(n)
                f.one[cells] = np.ones(n)
                f.convectionFlux[faces] = np.zeros(faces.getCount())
                for fg in mesh.getBoundaryFaceGroups():
                    f.massFlux[fg.site] = np.zeros(fg.site.getCount())
                lib = self.lib or mesh.device.lib
                self._systems[(m, mesh.getID())] = LinearSystem(lib, mesh=mesh.device, field_name=f.massFraction.name)
        self._niters = 0
        self._initialNorm = [None] * self._nSpecies
        self._currentResidual = [None] * self._nSpecies

    def _upload(self, m, mesh, ls):
        f, o = self._fields[m], self._options
        cells, faces = mesh.getCells(), mesh.getFaces()
        ls.set_field(capi.FIELD_X, f.massFraction[cells])
        ls.set_field(capi.FIELD_DIFFUSIVITY, f.diffusivity[cells])
        ls.set_field(capi.FIELD_SOURCE, f.source[cells])
        flux = f.convectionFlux[faces]
        convecting = bool(flux.any())
        if convecting:
            ls.set_field(capi.FIELD_FACE_FLUX, flux)
        if o.transient:
            ls.set_field(capi.FIELD_X_N1, f.massFractionN1[cells])
            ls.set_field(capi.FIELD_DENSITY, f.one[cells])
            if o.timeDiscretizationOrder > 1:
                ls.set_field(capi.FIELD_X_N2, f.massFractionN2[cells])
        for fg in mesh.getBoundaryFaceGroups():   # F/SpeciesModel_impl.h:549-606
            bc = self._bcMaps[m].get(fg.id)
            if bc is None:
                raise CException("SpeciesModel: Error in BC Map")
            if bc.bcType == "SpecifiedMassFraction":
                # the model always holds a convection flux: per face extrapolation where it leaves, else Dirichlet
                v = bc["specifiedMassFraction"]
                kind = capi.BC_DIRICHLET_OR_OUTFLOW if convecting else capi.BC_DIRICHLET
                if isinstance(v, np.ndarray):
                    ls.set_bc(fg.id, kind, [0.0], per_face=v)
                else:
                    ls.set_bc(fg.id, kind, [float(v)])
            elif bc.bcType == "SpecifiedMassFlux":
                ls.set_bc(fg.id, capi.BC_NEUMANN, [float(bc["specifiedMassFlux"])])
            elif bc.bcType == "Symmetry":
                ls.set_bc(fg.id, capi.BC_NEUMANN, [0.0])
            else:
                raise CException(bc.bcType + " not implemented for SpeciesModel")
        return convecting

    def advance(self, niter):
        """Impl::advance, F/SpeciesModel_impl.h:671-713: every outer iteration solves the species one after the other;
        the loop ends when all of them meet the tolerance."""
        o = self._options
        solver = o.getLinearSolver()
        if len(self.meshes) != 1:
            raise CException("SpeciesModelA: one mesh per model in this release")
        mesh = self.meshes[0]
        cells = mesh.getCells()
        for _ in range(niter):
            all_converged = True
            for m in range(self._nSpecies):
                f = self._fields[m]
                ls = self._systems[(m, mesh.getID())]
                convecting = self._upload(m, mesh, ls)
                ls.lib.timer_start(1)
                ls.assemble(diffusion=1, convection=(2 if o.useCentralDifference else 1) if convecting else 0, source=1,
                            time_order=(o.timeDiscretizationOrder if o.transient else 0),
                            dt=float(o["timeStep"]) if o.transient else 0.0, underrelax=0.0, apply_bcs=1,
                            eliminate_boundary=1)
                rnorm = solver.solve(ls)
                if self._initialNorm[m] is None:
                    self._initialNorm[m] = rnorm
                ratio = rnorm / self._initialNorm[m] if self._initialNorm[m] != 0 else 0.0
                print("Species Number: %d" % m)
                print("%d: [%s : %g]" % (self._niters, f.massFraction.name, rnorm))
                solver.cleanup()
                ls.post_solve_update()
                self.timings.append({"species": m, "ms": ls.lib.timer_stop(1), "rnorm": rnorm,
                                     "linear_iterations": getattr(solver, "lastIterations", 0)})
                f.massFraction[cells][:] = ls.get_field(capi.FIELD_X)
                bflux = ls.get_field(capi.FIELD_BFLUX)
                for fg in mesh.getBoundaryFaceGroups():
                    off = fg.site.getOffset()
                    f.massFlux[fg.site][:] = bflux[off:off + fg.site.getCount()]
                self._niters += 1
                self._currentResidual[m] = rnorm
                if not (rnorm < o.absoluteTolerance or ratio < o.relativeTolerance):
                    all_converged = False
            if all_converged:
                break

    def updateTime(self):  # F/SpeciesModel_impl.h:341-368
        o = self._options
        for f in self._fields:
            for mesh in self.meshes:
                cells = mesh.getCells()
                if o.timeDiscretizationOrder > 1:
                    f.massFractionN2[cells][:] = f.massFractionN1[cells]
                f.massFractionN1[cells][:] = f.massFraction[cells]

    def getMassFluxIntegral(self, mesh, faceGroupId, m):  # F/SpeciesModel_impl.h:614-652
        for fg in mesh.getBoundaryFaceGroups():
            if fg.id == faceGroupId:
                return float(np.sum(self._fields[m].massFlux[fg.site]))
        raise CException("getMassFluxIntegral: invalid faceGroupID")

    def getAverageMassFraction(self, mesh, m):  # :654-668
        cells = mesh.getCells()
        n = cells.getSelfCount()
        vol = np.asarray(self.geom.volume[cells])[:n]
        return float(np.sum(np.asarray(self._fields[m].massFraction[cells])[:n] * vol) / np.sum(vol))

    def getMassFractionResidual(self, speciesId):  # :739-748
        return self._currentResidual[speciesId]

    def printBCs(self):  # :715-737
        for m in range(self._nSpecies):
            print("Species Number :%d" % m)
            for gid in sorted(self._bcMaps[m]):
                bc = self._bcMaps[m][gid]
                print("Face Group %d:" % gid)
                print("    bc type " + bc.bcType)
                for k in sorted(bc):
                    print("   %s %g" % (k, bc[k]))


# ----------------------------------------------------------------------------- VacancyModel
class VacancyBC(FloatVarDict):
    """F/VacancyBC.h:8-20"""

    def __init__(self):
        super().__init__()
        self.defineVar("specifiedConcentration", 300.0)
        self.defineVar("specifiedVacaFlux", 0.0)
        self.defineVar("convectiveCoefficient", 0.0)
        self.defineVar("farFieldConcentration", 300.0)
        self.bcType = ""


class VacancyVC(FloatVarDict):
    """F/VacancyBC.h:22-34"""

    def __init__(self):
        super().__init__()
        self.defineVar("vacancyDiffusioncoefficient", 1.0)
        self.defineVar("density", 1.0)
        self.defineVar("specificVaca", 1.0)
        self.vcType = ""


class VacancyModelOptions(FloatVarDict):
    """F/VacancyBC.h:36-73"""

    def __init__(self):
        super().__init__()
        self.defineVar("initialConcentration", 300.0)
        self.defineVar("timeStep", 1e-7)
        self.relativeTolerance = 1e-8
        self.absoluteTolerance = 1e-16
        self.linearSolver = None
        self.useCentralDifference = False
        self.transient = False
        self.timeDiscretizationOrder = 1

    def getLinearSolver(self):
        if self.linearSolver is None:
            ls = AMG()
            ls.relativeTolerance = 1e-1
            ls.nMaxIterations = 20
            ls.verbosity = 0
            self.linearSolver = ls
        return self.linearSolver


class VacancyFields:
    """F/VacancyFields.h:11-30 (names as F/VacancyFields.cpp:7-21 spells them)"""

    def __init__(self, base_name):
        for n in ("concentration", "concentrationN1", "concentrationN2", "vacaFlux", "concentrationGradient",
                  "concentrationGradientVector", "diffusioncoefficient", "source", "convectionFlux", "specificVaca"):
            setattr(self, n, Field(base_name + "." + n))
        for n in ("plasticStrain", "zero", "one"):
            setattr(self, n, Field(base_name + n))


class VacancyModelA:
    """VacancyModel<double> (F/VacancyModel.h:18-55, Impl in F/VacancyModel_impl.h): the vacancy-concentration
    transport equation -- diffusion + convection + source [+ rho * specificVaca time derivative] with
    SpecifiedConcentration (per-face outflow rule) / SpecifiedVacaFlux / Symmetry / Convective boundaries -- on the
    scalar device path. computePlasticStrainRate (the gradient of the concentration gradient, a coupling to the
    structure models, :619-662) and the immersed-boundary hooks are not built."""

    def __init__(self, geom_fields, vacancy_fields, meshes, lib=None):
        self.geom, self.fields, self.meshes, self.lib = geom_fields, vacancy_fields, list(meshes), lib
        self._bcMap, self._vcMap = {}, {}
        self._options = VacancyModelOptions()
        self._initialNorm = None
        self._niters = 0
        self._systems = {}
        self.timings = []
        for mesh in self.meshes:  # F/VacancyModel_impl.h:57-86
            vc = VacancyVC()
            vc.vcType = "flow"
            self._vcMap[mesh.getID()] = vc
            for fg in mesh.getBoundaryFaceGroups():
                bc = VacancyBC()
                self._bcMap[fg.id] = bc
                if fg.groupType in ("wall", "symmetry"):
                    bc.bcType = "SpecifiedVacaFlux"
                elif fg.groupType in ("velocity-inlet", "pressure-outlet"):
                    bc.bcType = "SpecifiedConcentration"
                else:
                    raise CException("VacancyModel: unknown face group type " + fg.groupType)

    def getBCMap(self):
        return self._bcMap

    def getVCMap(self):
        return self._vcMap

    def getBC(self, gid):
        return self._bcMap[gid]

    def getOptions(self):
        return self._options

    def init(self):  # F/VacancyModel_impl.h:88-186
        f, o = self.fields, self._options
        for mesh in self.meshes:
            if mesh.device is None:
                raise CException("VacancyModel.init: mesh metrics not initialised (MeshMetricsCalculatorA.init)")
            cells, faces = mesh.getCells(), mesh.getFaces()
            n = cells.getCount()
            vc = self._vcMap[mesh.getID()]
            f.concentration[cells] = np.full(n, float(o["initialConcentration"]))
            if o.transient:
                f.concentrationN1[cells] = f.concentration[cells].copy()
                if o.timeDiscretizationOrder > 1:
                    f.concentrationN2[cells] = f.concentration[cells].copy()
            f.diffusioncoefficient[cells] = np.full(n, float(vc["vacancyDiffusioncoefficient"]))
            f.source[cells] = np.zeros(n)
            f.zero[cells] = np.zeros(n)
            f.one[cells] = np.ones(n)
            f.specificVaca[cells] = np.full(n, float(vc["density"]) * float(vc["specificVaca"]))
            f.concentrationGradient[cells] = np.zeros((n, 3))
            f.concentrationGradientVector[cells] = np.zeros((n, 3))
            f.plasticStrain[cells] = np.zeros((n, 3, 3))
            f.convectionFlux[faces] = np.zeros(faces.getCount())
            for fg in mesh.getBoundaryFaceGroups():
                f.vacaFlux[fg.site] = np.zeros(fg.site.getCount())
            lib = self.lib or mesh.device.lib
            self._systems[mesh.getID()] = LinearSystem(lib, mesh=mesh.device, field_name=f.concentration.name)
        self._niters = 0
        self._initialNorm = None

    def _upload(self, mesh, ls):
        f, o = self.fields, self._options
        cells, faces = mesh.getCells(), mesh.getFaces()
        ls.set_field(capi.FIELD_X, f.concentration[cells])
        ls.set_field(capi.FIELD_DIFFUSIVITY, f.diffusioncoefficient[cells])
        ls.set_field(capi.FIELD_SOURCE, f.source[cells])
        flux = f.convectionFlux[faces]
        convecting = bool(flux.any())
        if convecting:
            ls.set_field(capi.FIELD_FACE_FLUX, flux)
        if o.transient:
            ls.set_field(capi.FIELD_X_N1, f.concentrationN1[cells])
            ls.set_field(capi.FIELD_DENSITY, f.specificVaca[cells])
            if o.timeDiscretizationOrder > 1:
                ls.set_field(capi.FIELD_X_N2, f.concentrationN2[cells])

        def scalar_or_faces(v):
            return (0.0, v) if isinstance(v, np.ndarray) else (float(v), None)

        for fg in mesh.getBoundaryFaceGroups():   # F/VacancyModel_impl.h:316-381
            bc = self._bcMap[fg.id]
            if bc.bcType == "SpecifiedConcentration":
                v, pf = scalar_or_faces(bc["specifiedConcentration"])
                ls.set_bc(fg.id, capi.BC_DIRICHLET_OR_OUTFLOW if convecting else capi.BC_DIRICHLET, [v], per_face=pf)
            elif bc.bcType == "SpecifiedVacaFlux":
                v, pf = scalar_or_faces(bc["specifiedVacaFlux"])
                ls.set_bc(fg.id, capi.BC_NEUMANN, [v], per_face=pf)
            elif bc.bcType == "Symmetry":
                ls.set_bc(fg.id, capi.BC_NEUMANN, [0.0])
            elif bc.bcType == "Convective":
                ls.set_bc(fg.id, capi.BC_CONVECTIVE, [float(bc["convectiveCoefficient"]), float(bc["farFieldConcentration"])])
            else:
                raise CException(bc.bcType + " not implemented for VacancyModel")
        return convecting

    def advance(self, niter):
        """Impl::advance, F/VacancyModel_impl.h:424-456"""
        o = self._options
        solver = o.getLinearSolver()
        for mesh in self.meshes:
            ls = self._systems[mesh.getID()]
            cells = mesh.getCells()
            convecting = self._upload(mesh, ls)
            for _ in range(niter):
                ls.lib.timer_start(1)
                ls.assemble(diffusion=1, convection=(2 if o.useCentralDifference else 1) if convecting else 0, source=1,
                            time_order=(o.timeDiscretizationOrder if o.transient else 0),
                            dt=float(o["timeStep"]) if o.transient else 0.0, underrelax=0.0, apply_bcs=1,
                            eliminate_boundary=1)
                rnorm = solver.solve(ls)
                if self._initialNorm is None:
                    self._initialNorm = rnorm
                ratio = rnorm / self._initialNorm if self._initialNorm != 0 else 0.0
                print("%d: [%s : %g]" % (self._niters, self.fields.concentration.name, rnorm))
                solver.cleanup()
                ls.post_solve_update()
                self.timings.append({"ms": ls.lib.timer_stop(1), "rnorm": rnorm})
                self._niters += 1
                if rnorm < o.absoluteTolerance or ratio < o.relativeTolerance:
                    break
            self.fields.concentration[cells][:] = ls.get_field(capi.FIELD_X)
            bflux = ls.get_field(capi.FIELD_BFLUX)
            for fg in mesh.getBoundaryFaceGroups():
                off = fg.site.getOffset()
                self.fields.vacaFlux[fg.site][:] = bflux[off:off + fg.site.getCount()]

    def updateTime(self):  # F/VacancyModel_impl.h:472-494
        f, o = self.fields, self._options
        for mesh in self.meshes:
            cells = mesh.getCells()
            if o.timeDiscretizationOrder > 1:
                f.concentrationN2[cells][:] = f.concentrationN1[cells]
            f.concentrationN1[cells][:] = f.concentration[cells]

    def getVacaFluxIntegral(self, mesh, faceGroupId):  # :400-421
        for fg in mesh.getBoundaryFaceGroups():
            if fg.id == faceGroupId:
                return float(np.sum(self.fields.vacaFlux[fg.site]))
        raise CException("getVacaFluxIntegral: invalid faceGroupID")

    def computePlasticStrainRate(self):
        raise CException("VacancyModelA: computePlasticStrainRate is not built")

    def printBCs(self):  # :458-470
        for gid in sorted(self._bcMap):
            bc = self._bcMap[gid]
            print("Face Group %d:" % gid)
            print("    bc type " + bc.bcType)
            for k in sorted(bc):
                print("   %s %g" % (k, bc[k]))


# ----------------------------------------------------------------------------- FlowModel (SIMPLE)
class FlowBC(FloatVarDict):
    """F/FlowBC.h:9-21"""

    def __init__(self):
        super().__init__()
        self.defineVar("specifiedXVelocity", 0.0)
        self.defineVar("specifiedYVelocity", 0.0)
        self.defineVar("specifiedZVelocity", 0.0)
        self.defineVar("specifiedPressure", 0.0)
        self.defineVar("accomodationCoefficient", 1.0)
        self.bcType = ""


class FlowVC(FloatVarDict):
    """F/FlowBC.h:23-34"""

    def __init__(self):
        super().__init__()
        self.defineVar("viscosity", 1e-3)
        self.defineVar("density", 1.0)
        self.defineVar("eddyviscosity", 1e-5)
        self.defineVar("totalviscosity", 2e-3)
        self.vcType = ""


class FlowModelOptions(FloatVarDict):
    """F/FlowBC.h:37-115"""

    def __init__(self):
        super().__init__()
        for name, v in (("initialXVelocity", 0.0), ("initialYVelocity", 0.0), ("initialZVelocity", 0.0),
                        ("initialPressure", 0.0), ("momentumURF", 0.7), ("velocityURF", 1.0),
                        ("pressureURF", 0.3), ("timeStep", 0.1), ("operatingPressure", 101325.0),
                        ("operatingTemperature", 300.0), ("molecularWeight", 28.966)):
            self.defineVar(name, v)
        self.momentumTolerance = 1e-3
        self.continuityTolerance = 1e-3
        self.printNormalizedResiduals = True
        self.transient = False
        self.correctVelocity = True
        self.timeDiscretizationOrder = 1
        self.momentumLinearSolver = None
        self.pressureLinearSolver = None
        self.coupledLinearSolver = None
        self.incompressible = True
        self.turbulent = False

    @staticmethod
    def _default_solver():
        ls = AMG()
        ls.relativeTolerance = 1e-1
        ls.nMaxIterations = 20
        ls.verbosity = 0
        return ls

    def getMomentumLinearSolver(self):  # F/FlowBC.h:87-99
        if self.momentumLinearSolver is None:
            self.momentumLinearSolver = self._default_solver()
        return self.momentumLinearSolver

    def getPressureLinearSolver(self):  # F/FlowBC.h:101-113
        if self.pressureLinearSolver is None:
            self.pressureLinearSolver = self._default_solver()
        return self.pressureLinearSolver


class FlowFields:
    """F/FlowFields.h:15-41 (the fields the SIMPLE path reads or writes)"""

    def __init__(self, base_name):
        for n in ("velocity", "pressure", "massFlux", "velocityGradient", "pressureGradient", "momentumFlux",
                  "viscosity", "density", "continuityResidual", "velocityN1", "velocityN2"):
            setattr(self, n, Field(base_name + "." + n))


def _solver_args(solver):
    """(DeviceAMG-providing AMG mirror, bcgstab tuple or None) of a LinearSolver mirror."""
    if isinstance(solver, BCGStab):
        if solver.preconditioner is None:
            raise CException("BCGStab: no preconditioner set")
        kind = 2 if isinstance(solver.preconditioner, ILU0Solver) else 1   # fvmgpu_flow_solve_*'s bcgstab argument
        return solver.preconditioner, (solver.nMaxIterations, solver.relativeTolerance, solver.absoluteTolerance, kind)
    if isinstance(solver, JacobiSolver):
        return solver._amg, (solver.nMaxIterations, solver.relativeTolerance, solver.absoluteTolerance, 3)
    if isinstance(solver, CG):
        if solver.preconditioner is None:
            raise CException("CG: no preconditioner set")
        return solver.preconditioner, (solver.nMaxIterations, solver.relativeTolerance, solver.absoluteTolerance, 4)
    if not isinstance(solver, AMG):
        raise CException("FlowModelA: %s is not supported as a flow solver" % type(solver).__name__)
    return solver, None


class FlowModelA:
    """Mirror of `models_atyped_double.FlowModelA` (F/FlowModel.h:17-95, F/FlowModel.i): SIMPLE
    iterations -- momentum assembly + solve, Rhie-Chow pressure correction assembly + solve, the
    pressure / mass-flux / velocity corrections -- all on the device through the C ABI
    (fvmgpu_flow_*). Boundary types: "NoSlipWall", "SlipJump", "Symmetry", "VelocityBoundary",
    "PressureBoundary" (F/FlowModel_impl.h:636-677). On a partitioned mesh (fvm_b200.partition) every
    rank runs this same code on its part; halo exchanges and all-reduces happen inside the library."""

    def __init__(self, geom_fields, flow_fields, meshes, lib=None):
        self.geom, self.fields, self.meshes, self.lib = geom_fields, flow_fields, list(meshes), lib
        self._bcMap, self._vcMap = {}, {}
        self._options = FlowModelOptions()
        self._flows = {}
        self._niters = 0
        self._initialMomentumNorm = None
        self._initialContinuityNorm = None
        self.timings = []
        for mesh in self.meshes:  # FlowModel::Impl ctor, F/FlowModel_impl.h:93-141
            vc = FlowVC()
            vc.vcType = "flow"
            self._vcMap[mesh.getID()] = vc
            for fg in mesh.getBoundaryFaceGroups():
                bc = FlowBC()
                self._bcMap[fg.id] = bc
                if fg.groupType == "wall":
                    bc.bcType = "NoSlipWall"
                elif fg.groupType == "velocity-inlet":
                    bc.bcType = "VelocityBoundary"
                elif fg.groupType == "pressure-inlet" or fg.groupType == "pressure-outlet":
                    bc.bcType = "PressureBoundary"
                elif fg.groupType == "symmetry":
                    bc.bcType = "Symmetry"
                else:
                    raise CException("FlowModel: unknown face group type " + fg.groupType)

    def getBCMap(self):
        return self._bcMap

    def getVCMap(self):
        return self._vcMap

    def getOptions(self):
        return self._options

    def printBCs(self):  # F/FlowModel_impl.h:1569-1585
        for gid, bc in self._bcMap.items():
            print("Face Group %d:" % gid)
            print("    bc type " + bc.bcType)
            for k, v in bc.items():
                print("   %s  %s" % (k, v))

    def _flow_opts(self):
        o = self._options
        return capi.DeviceFlow.opts(float(o["momentumURF"]), float(o["pressureURF"]), int(o.transient),
                                    int(o.timeDiscretizationOrder), float(o["timeStep"]), int(o.correctVelocity),
                                    float(o["operatingPressure"]), float(o["operatingTemperature"]),
                                    float(o["molecularWeight"]), int(o.incompressible))

    def init(self):  # F/FlowModel_impl.h:148-340
        f, o = self.fields, self._options
        for mesh in self.meshes:
            if mesh.device is None:
                raise CException("FlowModel.init: mesh metrics not initialised (MeshMetricsCalculatorA.init)")
            cells, faces = mesh.getCells(), mesh.getFaces()
            n, nf = cells.getCount(), faces.getCount()
            vc = self._vcMap[mesh.getID()]
            v0 = np.array([o["initialXVelocity"], o["initialYVelocity"], o["initialZVelocity"]], float)
            f.velocity[cells] = np.tile(v0, (n, 1))
            if o.transient:
                f.velocityN1[cells] = f.velocity[cells].copy()
                if o.timeDiscretizationOrder > 1:
                    f.velocityN2[cells] = f.velocity[cells].copy()
            f.pressure[cells] = np.full(n, float(o["initialPressure"]))
            f.pressure[faces] = np.full(nf, float(o["initialPressure"]))
            f.density[cells] = np.full(n, float(vc["density"]))
            f.viscosity[cells] = np.full(n, float(vc["viscosity"]))
            f.pressureGradient[cells] = np.zeros((n, 3))
            f.velocityGradient[cells] = np.zeros((n, 9))
            f.continuityResidual[cells] = np.zeros(n)
            f.massFlux[faces] = np.zeros(nf)
            lib = self.lib or mesh.device.lib
            fl = capi.DeviceFlow(lib, mesh.device)
            self._flows[mesh.getID()] = fl
            if "cell_global" in mesh.raw:   # a mesh part: the reference cell is the globally lowest cell (:931-994)
                own = np.asarray(mesh.raw.cell_global)[:mesh.raw.n_cells]
                hit = np.nonzero(own == 0)[0]
                fl.set_reference_cell(int(hit[0]) if len(hit) else -1)
            self._upload(mesh, fl, with_flux=False)
            fl.init()
            f.massFlux[faces][:] = fl.get_field(capi.FLOW_MASS_FLUX)
            f.continuityResidual[cells][:] = fl.get_field(capi.FLOW_CONT_RESID)
        self._niters = 0
        self._initialMomentumNorm = None
        self._initialContinuityNorm = None

    def _upload(self, mesh, fl, with_flux=True):
        f, o = self.fields, self._options
        cells, faces = mesh.getCells(), mesh.getFaces()
        fl.set_field(capi.FLOW_VELOCITY, f.velocity[cells])
        fl.set_field(capi.FLOW_PRESSURE, f.pressure[cells])
        fl.set_field(capi.FLOW_FACE_PRESSURE, f.pressure[faces])
        fl.set_field(capi.FLOW_DENSITY, f.density[cells])
        fl.set_field(capi.FLOW_VISCOSITY, f.viscosity[cells])
        if with_flux:
            fl.set_field(capi.FLOW_MASS_FLUX, f.massFlux[faces])
            fl.set_field(capi.FLOW_CONT_RESID, f.continuityResidual[cells])
        if o.transient:
            fl.set_field(capi.FLOW_VELOCITY_N1, f.velocityN1[cells])
            if o.timeDiscretizationOrder > 1:
                fl.set_field(capi.FLOW_VELOCITY_N2, f.velocityN2[cells])
        for fg in mesh.getBoundaryFaceGroups():
            bc = self._bcMap[fg.id]
            if bc.bcType == "NoSlipWall":
                fl.set_bc(fg.id, capi.FLOWBC_NOSLIP_WALL, [float(bc["specifiedXVelocity"]),
                                                            float(bc["specifiedYVelocity"]),
                                                            float(bc["specifiedZVelocity"])])
            elif bc.bcType == "SlipJump":   # F/FlowModel_impl.h:666-672
                fl.set_bc(fg.id, capi.FLOWBC_SLIP_JUMP, [float(bc["specifiedXVelocity"]), float(bc["specifiedYVelocity"]),
                                                         float(bc["specifiedZVelocity"]), 0.0,
                                                         float(bc["accomodationCoefficient"])])
            elif bc.bcType == "Symmetry":
                fl.set_bc(fg.id, capi.FLOWBC_SYMMETRY, [0.0, 0.0, 0.0])
            elif bc.bcType in ("VelocityBoundary", "PressureBoundary"):
                kind = capi.FLOWBC_VELOCITY if bc.bcType == "VelocityBoundary" else capi.FLOWBC_PRESSURE
                fl.set_bc(fg.id, kind, [float(bc["specifiedXVelocity"]), float(bc["specifiedYVelocity"]),
                                        float(bc["specifiedZVelocity"]), float(bc["specifiedPressure"])])
            else:
                raise CException(bc.bcType + " not implemented for FlowModel")

    def _download(self, mesh, fl):
        f = self.fields
        cells, faces = mesh.getCells(), mesh.getFaces()
        f.velocity[cells][:] = fl.get_field(capi.FLOW_VELOCITY)
        f.pressure[cells][:] = fl.get_field(capi.FLOW_PRESSURE)
        f.pressure[faces][:] = fl.get_field(capi.FLOW_FACE_PRESSURE)
        f.massFlux[faces][:] = fl.get_field(capi.FLOW_MASS_FLUX)
        f.continuityResidual[cells][:] = fl.get_field(capi.FLOW_CONT_RESID)
        f.pressureGradient[cells][:] = fl.get_field(capi.FLOW_PRESSURE_GRADIENT)
        f.velocityGradient[cells][:] = fl.get_field(capi.FLOW_VELOCITY_GRADIENT)

    def advance(self, niter):
        """FlowModel::advance, F/FlowModel_impl.h:1433-1471 (one mesh per model). Returns True when
        both normalised residuals fall below the tolerances."""
        o = self._options
        if len(self.meshes) != 1:
            raise CException("FlowModelA: one mesh per model in this release")
        mesh = self.meshes[0]
        fl = self._flows[mesh.getID()]
        lib = fl.lib
        msolver, mbcg = _solver_args(o.getMomentumLinearSolver())
        psolver, pbcg = _solver_args(o.getPressureLinearSolver())
        mdev, pdev = msolver._device(lib), psolver._device(lib)
        if mdev is pdev:
            raise CException("FlowModelA: momentum and pressure solvers must be distinct objects")
        self._upload(mesh, fl)
        fo = self._flow_opts()
        converged = False
        vname, pname = self.fields.velocity.name, self.fields.pressure.name
        for _ in range(niter):
            t = {}
            lib.timer_start(1)
            fl.assemble_momentum(fo)
            t["momentum_assemble_ms"] = lib.timer_stop(1)
            lib.timer_start(1)
            mnorm, mits = fl.solve_momentum(mdev, mbcg)            # solveMomentum :730-770
            t["momentum_solve_ms"] = lib.timer_stop(1)
            mdev.cleanup()
            lib.timer_start(1)
            fl.assemble_continuity(fo)
            t["continuity_assemble_ms"] = lib.timer_stop(1)
            lib.timer_start(1)
            cnorm, cits = fl.solve_continuity(pdev, fo, pbcg)      # solveContinuity :1410-1430
            t["continuity_solve_ms"] = lib.timer_stop(1)
            pdev.cleanup()
            t["momentum_iterations"], t["pressure_iterations"] = [int(i) for i in mits], int(cits)
            if self._initialMomentumNorm is None:
                self._initialMomentumNorm = mnorm.copy()
            if self._initialContinuityNorm is None:
                self._initialContinuityNorm = cnorm
            if self._niters < 5:  # setMax of the first iterations' norms, :1441-1445
                self._initialMomentumNorm = np.maximum(self._initialMomentumNorm, mnorm)
                self._initialContinuityNorm = max(self._initialContinuityNorm, cnorm)
            with np.errstate(divide="ignore", invalid="ignore"):
                mratio = np.where(self._initialMomentumNorm > 0, mnorm / self._initialMomentumNorm, 0.0)
            cratio = cnorm / self._initialContinuityNorm if self._initialContinuityNorm > 0 else 0.0
            mshow, cshow = (mratio, cratio) if o.printNormalizedResiduals else (mnorm, cnorm)
            print("%d: [%s : [%g %g %g]];[%s : %g]" % (self._niters, vname, mshow[0], mshow[1], mshow[2], pname, cshow))
            t["momentum_norm"], t["continuity_norm"] = mnorm, cnorm
            self.timings.append(t)
            self._niters += 1
            if np.all(mratio < o.momentumTolerance) and cratio < o.continuityTolerance:
                converged = True
                break
        self._download(mesh, fl)
        return converged

    def dumpContinuityMatrix(self, file_base):
        """Impl::dumpContinuityMatrix, F/FlowModel_impl.h:1560-1622: one momentum solve, then the pressure-
        correction system as a MatrixMarket file + right-hand side (T/FLOW_CONTINUITY_MATRIX)."""
        o = self._options
        mesh = self.meshes[0]
        fl = self._flows[mesh.getID()]
        msolver, mbcg = _solver_args(o.getMomentumLinearSolver())
        mdev = msolver._device(fl.lib)
        self._upload(mesh, fl)
        fo = self._flow_opts()
        fl.assemble_momentum(fo)
        fl.solve_momentum(mdev, mbcg)
        mdev.cleanup()
        fl.assemble_continuity(fo)
        d = fl.download_continuity()
        n = mesh.getCells().getSelfCount()
        row, col = mesh.cc_row, mesh.cc_col
        lines = []
        for i in range(n):
            lines.append("%d %d %f" % (i + 1, i + 1, d["diag"][i]))
            for jp in range(row[i], row[i + 1]):
                if col[jp] < n:
                    lines.append("%d %d %f" % (i + 1, col[jp] + 1, d["offdiag"][jp]))
        with open(file_base + ".mat", "w") as fh:
            fh.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (n, n, len(lines)))
            fh.write("\n".join(lines) + "\n")
        with open(file_base + ".rhs", "w") as fh:
            fh.write("".join("%f\n" % (-v) for v in d["b"][:n]))
        self._download(mesh, fl)

    def updateTime(self):  # F/FlowModel_impl.h:351-370
        f, o = self.fields, self._options
        for mesh in self.meshes:
            cells = mesh.getCells()
            if o.timeDiscretizationOrder > 1:
                f.velocityN2[cells][:] = f.velocityN1[cells]
            f.velocityN1[cells][:] = f.velocity[cells]

    def getPressureIntegral(self, mesh, face_group_id):  # F/FlowModel_impl.h:1587-1611: sum of pFace * A
        faces = mesh.getFaces()
        for fg in mesh.getBoundaryFaceGroups():
            if fg.id == face_group_id:
                o, c = fg.site.getOffset(), fg.site.getCount()
                return (self.fields.pressure[faces][o:o + c, None] * self.geom.area[faces][o:o + c]).sum(axis=0)
        raise CException("getPressureIntegral: invalid faceGroupID")


# ----------------------------------------------------------------------------- ElectricModel
K_SI, QE, E0_SI = 1.3806503e-23, 1.60217646e-19, 8.854187826e-12   # F/PhysicsConstant.h


class ElectricBC(FloatVarDict):
    """F/ElectricBC.h:13-29"""

    def __init__(self):
        super().__init__()
        for name, v in (("specifiedPotential", 300.0), ("specifiedPotentialFlux", 0.0), ("specifiedXElecField", 0.0),
                        ("specifiedYElecField", 0.0), ("specifiedZElecField", 0.0), ("specifiedCharge", 0.0),
                        ("specifiedChargeFlux", 0.0), ("timeStep", 1.0)):
            self.defineVar(name, v)
        self.bcType = ""


class ElectricVC(FloatVarDict):
    """F/ElectricBC.h:31-39"""

    def __init__(self):
        super().__init__()
        self.defineVar("dielectric_constant", 7.9)
        self.vcType = ""


class ElectricModelConstants(FloatVarDict):
    """F/ElectricBC.h:41-75"""

    def __init__(self):
        super().__init__()
        for name, v in (("dielectric_ionization", 3.0), ("dielectric_bandgap", 5.0), ("optical_dielectric_constant", 4.0),
                        ("dielectric_thickness", 2.5e-7), ("dielectric_constant", 7.9), ("electron_capture_cross", 1e-17),
                        ("membrane_workfunction", 5.0), ("substrate_workfunction", 5.0), ("membrane_voltage", 0.0),
                        ("substrate_voltage", 0.0), ("OP_temperature", 300.0), ("electron_effmass", 0.5),
                        ("poole_frenkel_emission_frequency", 1.0e12), ("electron_mobility", 50e4),
                        ("electron_saturation_velocity", 1e9), ("voltage", 100.0), ("substrate_id", 5), ("membrane_id", 4),
                        ("nLevel", 0), ("normal_direction", 2), ("nTrap", 1)):
            self.defineVar(name, v)
        self.electron_trapdensity = []
        self.electron_trapdepth = []


class ElectricModelOptions(FloatVarDict):
    """F/ElectricBC.h:78-169"""

    def __init__(self):
        super().__init__()
        for name, v in (("initialCharge", 0.0), ("initialPotential", 0.0), ("initialTotalCharge", 0.0),
                        ("initialTunnelingCharge", 1.0), ("timeStep", 0.1), ("Interface_A_coeff", 1.0),
                        ("Interface_B_coeff", 0.0)):
            self.defineVar(name, v)
        self.electrostaticsTolerance = 1e-8
        self.chargetransportTolerance = 1e-8
        self.electrostaticsLinearSolver = None
        self.chargetransportLinearSolver = None
        self.timeDiscretizationOrder = 1
        self.transient_enable = True
        self.ibm_enable = False
        self.electrostatics_enable = True
        self.chargetransport_enable = True
        self.tunneling_enable = False
        self.emission_enable = False
        self.capture_enable = False
        self.injection_enable = False
        self.drift_enable = False
        self.diffusion_enable = False
        self.trapbandtunneling_enable = False
        self.ButlerVolmer = False
        self.printNormalizedResiduals = True

    @staticmethod
    def _default_solver():  # F/ElectricBC.h:140-165
        ls = AMG()
        ls.relativeTolerance = 1e-3
        ls.nMaxIterations = 20
        ls.verbosity = 0
        return ls

    def getElectroStaticsLinearSolver(self):
        if self.electrostaticsLinearSolver is None:
            self.electrostaticsLinearSolver = self._default_solver()
        return self.electrostaticsLinearSolver

    def getChargeTransportLinearSolver(self):
        if self.chargetransportLinearSolver is None:
            self.chargetransportLinearSolver = self._default_solver()
        return self.chargetransportLinearSolver


class ElectricFields:
    """F/ElectricFields.h:18-52 (fields the electrostatics + drift path reads or writes)"""

    def __init__(self, base_name):
        for n in ("potential", "potential_flux", "potential_gradient", "electric_field", "dielectric_constant",
                  "total_charge", "electron_velocity", "charge", "chargeFlux", "convectionFlux", "chargeN1",
                  "chargeN2", "one", "zero"):
            setattr(self, n, Field(base_name + "." + n))


class ElectricModelA:
    """Mirror of `models_atyped_double.ElectricModelA` (F/ElectricModel.h, F/ElectricModel_impl.h).

    * electrostatics (F/ElectricModel_impl.h:377-410, 552-767): Poisson equation for the potential --
      diffusion with dielectric_constant, total_charge source, BCs SpecifiedPotential,
      SpecifiedPotentialFlux, Symmetry, SpecialDielectricBoundary -- then updateElectricField.
    * charge transport (:412-436, 771-924) with the drift and transient terms: only component nTrap of
      the charge vector is convected (DriftDiscretization); without the tunnelling / injection /
      emission / capture source models (out of scope, SURVEY §2: 1-D column physics) the reference's
      3x3 blocks stay diagonal, so each component is solved as a scalar system on the device.
    Everything numerical runs on the GPU through the C ABI."""

    def __init__(self, geom_fields, electric_fields, meshes, lib=None):
        self.geom, self.fields, self.meshes, self.lib = geom_fields, electric_fields, list(meshes), lib
        self._bcMap, self._vcMap = {}, {}
        self._options, self._constants = ElectricModelOptions(), ElectricModelConstants()
        self._pot, self._chg = {}, {}
        self._niters = 0
        self._timing, self.timings = {}, []
        self._initialElectroStaticsNorm = None
        self._initialChargeTransportNorm = None
        for mesh in self.meshes:  # Impl ctor, F/ElectricModel_impl.h:71-110
            vc = ElectricVC()
            vc.vcType = "dielectric"
            self._vcMap[mesh.getID()] = vc
            for fg in mesh.getBoundaryFaceGroups():
                bc = ElectricBC()
                self._bcMap[fg.id] = bc
                if fg.groupType == "wall":
                    bc.bcType = "SpecifiedPotential"
                elif fg.groupType == "symmetry":
                    bc.bcType = "Symmetry"

    def getBCMap(self):
        return self._bcMap

    def getBC(self, gid):
        return self._bcMap[gid]

    def getVCMap(self):
        return self._vcMap

    def getVC(self, mid):
        return self._vcMap[mid]

    def getOptions(self):
        return self._options

    def getConstants(self):
        return self._constants

    def _unsupported(self):
        o = self._options
        for flag in ("tunneling_enable", "emission_enable", "capture_enable", "injection_enable",
                     "trapbandtunneling_enable", "diffusion_enable", "ibm_enable", "ButlerVolmer"):
            if getattr(o, flag):
                raise CException("ElectricModelA: %s is not supported on the GPU path" % flag)

    def init(self):  # F/ElectricModel_impl.h:112-330
        f, o, c = self.fields, self._options, self._constants
        self._unsupported()
        for mesh in self.meshes:
            if mesh.device is None:
                raise CException("ElectricModel.init: mesh metrics not initialised (MeshMetricsCalculatorA.init)")
            cells, faces = mesh.getCells(), mesh.getFaces()
            n, nf = cells.getCount(), faces.getCount()
            vc = self._vcMap[mesh.getID()]
            lib = self.lib or mesh.device.lib
            if o.electrostatics_enable:
                f.potential[cells] = np.full(n, float(o["initialPotential"]))
                f.dielectric_constant[cells] = np.full(n, float(vc["dielectric_constant"]) * E0_SI)
                f.total_charge[cells] = (np.full(n, float(o["initialTotalCharge"]) * -QE)
                                         if vc.vcType == "dielectric" else np.zeros(n))
                f.potential_gradient[cells] = np.zeros((n, 3))
                f.electric_field[cells] = np.zeros((n, 3))
                for fg in mesh.getBoundaryFaceGroups():
                    f.potential_flux[fg.site] = np.zeros(fg.site.getCount())
                self._pot[mesh.getID()] = LinearSystem(lib, mesh=mesh.device, field_name=f.potential.name)
            if o.chargetransport_enable and vc.vcType == "dielectric":
                f.charge[cells] = np.zeros((n, 3))
                if o.transient_enable:
                    f.chargeN1[cells] = np.zeros((n, 3))
                    if o.timeDiscretizationOrder > 1:
                        f.chargeN2[cells] = np.zeros((n, 3))
                f.electron_velocity[cells] = np.zeros((n, 3))
                f.convectionFlux[faces] = np.zeros(nf)
                f.one[cells] = np.ones(n)
                f.zero[cells] = np.zeros(n)
                self._chg[mesh.getID()] = LinearSystem(lib, mesh=mesh.device, field_name=f.charge.name)
        self._niters = 0
        self._initialElectroStaticsNorm = None
        self._initialChargeTransportNorm = None

    def updateTime(self):  # F/ElectricModel_impl.h:338-375
        f, o = self.fields, self._options
        for mesh in self.meshes:
            cells = mesh.getCells()
            if o.timeDiscretizationOrder > 1:
                f.chargeN2[cells][:] = f.chargeN1[cells]
            f.chargeN1[cells][:] = f.charge[cells]

    # ---- electrostatics
    def _assemble_electrostatics(self, mesh):
        """initElectroStaticsLinearization + linearizeElectroStatics + initSolve (F/ElectricModel_impl.h:552-767);
        returns the system and the ids of the symmetry groups"""
        f, o, c = self.fields, self._options, self._constants
        ls = self._pot[mesh.getID()]
        cells = mesh.getCells()
        ls.set_field(capi.FIELD_X, f.potential[cells])
        ls.set_field(capi.FIELD_DIFFUSIVITY, f.dielectric_constant[cells])
        ls.set_field(capi.FIELD_SOURCE, f.total_charge[cells])
        sym = []
        for fg in mesh.getBoundaryFaceGroups():
            bc = self._bcMap[fg.id]
            if bc.bcType == "SpecifiedPotential":
                v = bc["specifiedPotential"]
                if isinstance(v, np.ndarray):
                    ls.set_bc(fg.id, capi.BC_DIRICHLET, [0.0], per_face=v)
                else:
                    ls.set_bc(fg.id, capi.BC_DIRICHLET, [float(v)])
            elif bc.bcType == "SpecifiedPotentialFlux":
                ls.set_bc(fg.id, capi.BC_NEUMANN, [float(bc["specifiedPotentialFlux"])])
            elif bc.bcType == "Symmetry":
                ls.set_bc(fg.id, capi.BC_NEUMANN, [0.0])
                sym.append(fg.id)
            elif bc.bcType == "SpecialDielectricBoundary":  # applyDielectricInterfaceBC (src = 0 in the model), :734-745
                coeff = float(c["dielectric_constant"]) * E0_SI / float(c["dielectric_thickness"])
                v = bc["specifiedPotential"]
                if isinstance(v, np.ndarray):
                    ls.set_bc(fg.id, capi.BC_DIELECTRIC_INTERFACE, [0.0, coeff, 0.0], per_face=v)
                else:
                    ls.set_bc(fg.id, capi.BC_DIELECTRIC_INTERFACE, [float(v), coeff, 0.0])
            else:
                raise CException(bc.bcType + " not implemented for ElectricModel")
        ls.lib.timer_start(1)
        ls.assemble(diffusion=1, convection=0, source=1, time_order=0, dt=0.0, underrelax=0.0, apply_bcs=1,
                    eliminate_boundary=1, interface_thickness=float(c["dielectric_thickness"]))
        return ls, sym

    def _solve_electrostatics(self, mesh):
        f, o, c = self.fields, self._options, self._constants
        cells = mesh.getCells()
        ls, sym = self._assemble_electrostatics(mesh)
        solver = o.getElectroStaticsLinearSolver()
        rnorm = solver.solve(ls)
        solver.cleanup()
        ls.post_solve_update()
        self._timing["electrostatics_ms"] = self._timing.get("electrostatics_ms", 0.0) + ls.lib.timer_stop(1)
        self._timing["electrostatics_iterations"] = int(getattr(solver, "lastIterations", 0))
        f.potential[cells][:] = ls.get_field(capi.FIELD_X)
        bflux = ls.get_field(capi.FIELD_BFLUX)
        for fg in mesh.getBoundaryFaceGroups():
            off = fg.site.getOffset()
            f.potential_flux[fg.site][:] = bflux[off:off + fg.site.getCount()]
        # updateElectricField (+ updateElectronVelocity, updateConvectionFlux when charge transport is on)
        f.electric_field[cells][:] = ls.electric_field()
        f.potential_gradient[cells][:] = -f.electric_field[cells]
        if o.chargetransport_enable and mesh.getID() in self._chg:
            vel = ls.drift_flux_into(self._chg[mesh.getID()], float(c["electron_mobility"]),
                                     float(c["electron_saturation_velocity"]), sym)
            f.electron_velocity[cells][:] = vel
            f.convectionFlux[mesh.getFaces()][:] = self._chg[mesh.getID()].get_field(capi.FIELD_FACE_FLUX)
        return rnorm

    # ---- charge transport: one scalar system per component of the charge vector
    def _solve_charge_transport(self, mesh):
        f, o, c = self.fields, self._options, self._constants
        ls = self._chg[mesh.getID()]
        cells, faces = mesh.getCells(), mesh.getFaces()
        n_trap = int(c["nTrap"])
        if not 0 <= n_trap <= 2:
            raise CException("ElectricModelA: nTrap must be 0..2 (the charge vector has 3 components)")
        ls.set_field(capi.FIELD_FACE_FLUX, f.convectionFlux[faces])
        ls.set_field(capi.FIELD_DENSITY, f.one[cells])
        bcs_on = o.drift_enable or o.diffusion_enable
        for fg in mesh.getBoundaryFaceGroups():   # "dielectric charging uses fixed zero dirichlet bc", :893-905
            ls.set_bc(fg.id, capi.BC_DIRICHLET, [0.0])
        norms = np.zeros(3)
        solver = o.getChargeTransportLinearSolver()
        for k in range(3):
            ls.set_field(capi.FIELD_X, np.ascontiguousarray(f.charge[cells][:, k]))
            order = 0
            if o.transient_enable:
                order = o.timeDiscretizationOrder
                ls.set_field(capi.FIELD_X_N1, np.ascontiguousarray(f.chargeN1[cells][:, k]))
                if order > 1:
                    ls.set_field(capi.FIELD_X_N2, np.ascontiguousarray(f.chargeN2[cells][:, k]))
            ls.lib.timer_start(1)
            ls.assemble(diffusion=0, convection=1 if (o.drift_enable and k == n_trap) else 0, source=0,
                        time_order=order, dt=float(o["timeStep"]), underrelax=0.0, apply_bcs=1 if bcs_on else 0,
                        eliminate_boundary=1)
            norms[k] = solver.solve(ls)
            solver.cleanup()
            ls.post_solve_update()
            self._timing["charge_ms"] = self._timing.get("charge_ms", 0.0) + ls.lib.timer_stop(1)
            f.charge[cells][:, k] = ls.get_field(capi.FIELD_X)
        return norms

    def advance(self, niter):
        """ElectricModel::advance, F/ElectricModel_impl.h:929-998. Returns the electrostatics flag."""
        o = self._options
        self._unsupported()
        if len(self.meshes) != 1:
            raise CException("ElectricModelA: one mesh per model in this release")
        mesh = self.meshes[0]
        flag1 = False
        self._timing = {}          # device time (CUDA events) of assembly + solve + update per equation, this call
        self.timings.append(self._timing)
        if o.electrostatics_enable:
            for _ in range(niter):
                e = self._solve_electrostatics(mesh)
                if self._initialElectroStaticsNorm is None:
                    self._initialElectroStaticsNorm = e
                if self._niters < 5:
                    self._initialElectroStaticsNorm = max(self._initialElectroStaticsNorm, e)
                ratio = e / self._initialElectroStaticsNorm if self._initialElectroStaticsNorm > 0 else 0.0
                print("%d: [%s : %g];" % (self._niters, self.fields.potential.name,
                                          ratio if o.printNormalizedResiduals else e))
                if ratio < o.electrostaticsTolerance:
                    flag1 = True
                    break
        if o.chargetransport_enable and mesh.getID() in self._chg:
            for _ in range(niter):
                cn = self._solve_charge_transport(mesh)
                if self._initialChargeTransportNorm is None:
                    self._initialChargeTransportNorm = cn.copy()
                if self._niters < 5:
                    self._initialChargeTransportNorm = np.maximum(self._initialChargeTransportNorm, cn)
                with np.errstate(divide="ignore", invalid="ignore"):
                    cr = np.where(self._initialChargeTransportNorm > 0, cn / self._initialChargeTransportNorm, 0.0)
                show = cr if o.printNormalizedResiduals else cn
                print("%d: [%s : [%g %g %g]]" % (self._niters, self.fields.charge.name, show[0], show[1], show[2]))
        self._niters += 1
        return flag1
