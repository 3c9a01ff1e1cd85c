// fvm_gpu_adaptor.h -- the reference-side binding of libfvmgpu.so.
//
// This header is what a maintainer of btanasoi/fvm adds next to src/fvm/src/modules/fvmbase/: it
// is compiled AGAINST THE REFERENCE'S OWN HEADERS and plugs the CUDA path in behind the two
// interfaces the hot path sits behind, without touching the SWIG/Python surface:
//
//   class GpuAMG     : public LinearSolver   (F/LinearSolver.h:11-31)  -- drop-in for AMG
//   class GpuBCGStab : public LinearSolver                              -- drop-in for BCGStab
//   GpuCG / GpuJacobiSolver / GpuILU0Solver / GpuBCGStabILU0            -- drop-ins for CG, JacobiSolver, ILU0Solver
//   class GpuScalarLinearizer                                           -- replaces the body of
//        ThermalModel<T>::Impl::linearize (F/ThermalModel_impl.h:236-398: GradientModel::compute,
//        Linearizer::linearize over the discretization list and the GenericBCS loop) plus
//        LinearSystem::initSolve's boundary elimination, writing into the reference's own
//        CRMatrix<T,T,T> / MultiField arrays; `iterateOnDevice` is the whole outer iteration
//        (linearize -> solve -> postSolve -> updateSolution, F/ThermalModel_impl.h:424-456) with the matrix
//        never leaving the device: only the fields go up and the new field comes down.
//   class GpuFlowModel                                                  -- FlowModel<T>::advance
//        (F/FlowModel_impl.h:730-770, 1410-1471: solveMomentum, solveContinuity, residual bookkeeping) on the
//        device, driven by the reference's own FlowModel object for boundary conditions, material and options,
//        reading and writing the reference's FlowFields arrays.
//
// Scripts select it exactly like any other solver:  tmodel.getOptions().linearSolver = GpuAMG()
// Errors from the C ABI are rethrown as CException (F/CException.h:16-21).
#ifndef FVM_GPU_ADAPTOR_H_
#define FVM_GPU_ADAPTOR_H_

#include <algorithm>
#include <cmath>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "AMG.h"
#include "CException.h"
#include "CRMatrix.h"
#include "FlowFields.h"
#include "FlowModel.h"
#include "GeomFields.h"
#include "LinearSolver.h"
#include "LinearSystem.h"
#include "Mesh.h"
#include "MultiFieldReduction.h"
#include "fvmgpu.h"

namespace fvmgpu_adaptor {

inline void check(int rc) {
  if (rc != 0) throw CException(std::string("libfvmgpu: ") + fvmgpu_last_error());
}

inline void ensureInit(int device = 0) { check(fvmgpu_init(device)); }

typedef CRMatrix<double, double, double> ScalarMatrix;
typedef Array<double> DArray;

// The single scalar cell system of a LinearSystem: (ArrayIndex, matrix)
struct ScalarSystemView {
  MultiField::ArrayIndex index;
  ScalarMatrix* matrix;
  ScalarSystemView() : index(0, 0), matrix(0) {}
};

inline ScalarSystemView findScalarSystem(LinearSystem& ls) {
  ScalarSystemView v;
  const MultiField::ArrayIndexList& idx = ls.getB().getArrayIndices();
  for (size_t i = 0; i < idx.size(); i++) {
    if (!ls.getMatrix().hasMatrix(idx[i], idx[i])) continue;
    ScalarMatrix* m = dynamic_cast<ScalarMatrix*>(&ls.getMatrix().getMatrix(idx[i], idx[i]));
    if (m) {
      if (v.matrix) throw CException("GpuAMG: more than one CRMatrix<double,double,double> block");
      v.matrix = m;
      v.index = idx[i];
    }
  }
  if (!v.matrix) throw CException("GpuAMG: no CRMatrix<double,double,double> block in this LinearSystem");
  return v;
}

// ------------------------------------------------------------------------------------ solvers
class GpuAMG : public LinearSolver {
 public:
  GpuAMG()
      : maxCoarseLevels(30), nPreSweeps(0), nPostSweeps(1), coarseGroupSize(2), weightRatioThreshold(0.65),
        cycleType(AMG::V_CYCLE), smootherType(AMG::GAUSS_SEIDEL), scaleCorrections(true), _solver(0), _system(0),
        _for(0), _totalIterations(0) {}
  virtual ~GpuAMG() { cleanup(); if (_solver) fvmgpu_amg_destroy(_solver); }

  // same public tunables as AMG (F/AMG.h:74-81)
  int maxCoarseLevels, nPreSweeps, nPostSweeps, coarseGroupSize;
  double weightRatioThreshold;
  AMG::CycleType cycleType;
  AMG::SmootherType smootherType;
  bool scaleCorrections;

  virtual MFRPtr solve(LinearSystem& ls) {
    ScalarSystemView v = upload(ls);
    double r0 = 0, r = 0;
    int it = 0;
    check(fvmgpu_amg_solve(_solver, _system, &r0, &r, &it));
    _totalIterations += it;
    download(ls, v);
    if (verbosity > 0) {
      std::cout << "0: [" << v.index.first->getName() << " : " << r0 << "]" << std::endl;
      std::cout << it << ": [" << v.index.first->getName() << " : " << r << "]" << std::endl;
    }
    return norm(v, r0);
  }
  virtual void smooth(LinearSystem& ls) {  // one cycle on (b, delta), used by BCGStab (F/BCGStab.cpp:85-89)
    ScalarSystemView v = upload(ls);
    check(fvmgpu_amg_smooth(_solver, _system));
    download(ls, v);
  }
  virtual void cleanup() {
    if (_solver) fvmgpu_amg_cleanup(_solver);
    if (_system) { fvmgpu_system_destroy(_system); _system = 0; }
    _for = 0;
  }
  int getTotalIterations() const { return _totalIterations; }
  fvmgpu_solver_t handle() { return _solver; }
  fvmgpu_system_t system() { return _system; }

  // AMG::solve on a system that already lives on the device (assembled there by GpuScalarLinearizer): nothing is
  // copied; delta stays in the system's FVMGPU_FIELD_DELTA. Returns the initial residual norm.
  double solveOnDevice(fvmgpu_system_t sys, int* iterations = 0) {
    syncOptions();
    double r0 = 0, r = 0;
    int it = 0;
    check(fvmgpu_amg_solve(_solver, sys, &r0, &r, &it));
    _totalIterations += it;
    if (iterations) *iterations = it;
    return r0;
  }
  void syncOptions() {
    ensureInit();
    fvmgpu_amg_opts o;
    o.nMaxIterations = nMaxIterations; o.verbosity = verbosity;
    o.relativeTolerance = relativeTolerance; o.absoluteTolerance = absoluteTolerance;
    o.maxCoarseLevels = maxCoarseLevels; o.nPreSweeps = nPreSweeps; o.nPostSweeps = nPostSweeps;
    o.coarseGroupSize = coarseGroupSize; o.weightRatioThreshold = weightRatioThreshold;
    o.cycleType = (int)cycleType; o.smootherType = (int)smootherType;
    if (!_solver) check(fvmgpu_amg_create(&_solver, &o));
    else check(fvmgpu_amg_set_opts(_solver, &o));
  }

  ScalarSystemView upload(LinearSystem& ls) {
    ensureInit();
    fvmgpu_amg_opts o;
    o.nMaxIterations = nMaxIterations; o.verbosity = verbosity;
    o.relativeTolerance = relativeTolerance; o.absoluteTolerance = absoluteTolerance;
    o.maxCoarseLevels = maxCoarseLevels; o.nPreSweeps = nPreSweeps; o.nPostSweeps = nPostSweeps;
    o.coarseGroupSize = coarseGroupSize; o.weightRatioThreshold = weightRatioThreshold;
    o.cycleType = (int)cycleType; o.smootherType = (int)smootherType;
    if (!_solver) check(fvmgpu_amg_create(&_solver, &o));
    else check(fvmgpu_amg_set_opts(_solver, &o));
    ScalarSystemView v = findScalarSystem(ls);
    const CRConnectivity& conn = v.matrix->getConnectivity();
    const StorageSite& site = conn.getRowSite();
    const DArray& b = dynamic_cast<const DArray&>(ls.getB()[v.index]);
    const DArray& delta = dynamic_cast<const DArray&>(ls.getDelta()[v.index]);
    if (_for != &ls) {  // AMG keys its hierarchy on the LinearSystem (F/AMG.cpp:222-226)
      if (_system) fvmgpu_system_destroy(_system);
      _system = 0;
      check(fvmgpu_system_create_raw(&_system, site.getSelfCount(), site.getCount() - site.getSelfCount(),
                                     (const int*)conn.getRow().getData(), (const int*)conn.getCol().getData(),
                                     (const double*)v.matrix->getDiag().getData(),
                                     (const double*)v.matrix->getOffDiag().getData(), (const double*)b.getData()));
      _for = &ls;
    } else {
      check(fvmgpu_system_set_field(_system, FVMGPU_FIELD_B, (const double*)b.getData(), b.getLength()));
    }
    check(fvmgpu_system_set_field(_system, FVMGPU_FIELD_DELTA, (const double*)delta.getData(), delta.getLength()));
    return v;
  }
  void download(LinearSystem& ls, const ScalarSystemView& v) {
    DArray& delta = dynamic_cast<DArray&>(ls.getDelta()[v.index]);
    check(fvmgpu_system_get_field(_system, FVMGPU_FIELD_DELTA, (double*)delta.getData(), delta.getLength()));
  }
  static MFRPtr norm(const ScalarSystemView& v, double r0) {
    MFRPtr r(new MultiFieldReduction());
    shared_ptr<DArray> a(new DArray(1));
    (*a)[0] = r0;
    r->addArray(*v.index.first, a);
    return r;
  }

 private:
  fvmgpu_solver_t _solver;
  fvmgpu_system_t _system;
  LinearSystem* _for;
  int _totalIterations;
};

class GpuBCGStab : public LinearSolver {
 public:
  GpuBCGStab() : preconditioner(0), _totalIterations(0) {}
  GpuAMG* preconditioner;
  virtual MFRPtr solve(LinearSystem& ls) {
    if (!preconditioner) throw CException("GpuBCGStab: no preconditioner");
    ScalarSystemView v = preconditioner->upload(ls);
    double r0 = 0, r = 0;
    int it = 0;
    check(fvmgpu_bcgstab_solve(preconditioner->handle(), preconditioner->system(), nMaxIterations,
                               relativeTolerance, absoluteTolerance, &r0, &r, &it));
    _totalIterations += it;
    preconditioner->download(ls, v);
    if (verbosity > 0) {
      std::cout << "0: [" << v.index.first->getName() << " : " << r0 << "]" << std::endl;
      std::cout << it << ": [" << v.index.first->getName() << " : " << r << "]" << std::endl;
    }
    return GpuAMG::norm(v, r0);
  }
  virtual void smooth(LinearSystem&) { throw CException("cannot use BCGStab as preconditioner"); }
  virtual void cleanup() { if (preconditioner) preconditioner->cleanup(); }
  int getTotalIterations() const { return _totalIterations; }

 private:
  int _totalIterations;
};

// CG (F/CG.cpp:24-140), JacobiSolver (F/JacobiSolver.cpp:46-95), ILU0Solver (F/ILU0Solver.cpp:46-93) and BCGStab with
// the ILU0 preconditioner (T/PARALLEL_CAVITY_ILU0): the same pattern, other entry points.
class GpuKrylovLike : public LinearSolver {
 public:
  enum Kind { CG_AMG, JACOBI, ILU0, BCGSTAB_ILU0 };
  explicit GpuKrylovLike(Kind k) : preconditioner(0), _kind(k), _totalIterations(0) {}
  GpuAMG* preconditioner;   // CG_AMG: the AMG whose cycle preconditions; the others: any GpuAMG (it owns the device system)
  virtual MFRPtr solve(LinearSystem& ls) {
    if (!preconditioner) throw CException("GpuKrylovLike: set `preconditioner` to a GpuAMG");
    ScalarSystemView v = preconditioner->upload(ls);
    double r0 = 0, r = 0;
    int it = 0;
    fvmgpu_solver_t h = preconditioner->handle();
    fvmgpu_system_t sys = preconditioner->system();
    switch (_kind) {
      case CG_AMG: check(fvmgpu_cg_solve(h, sys, nMaxIterations, relativeTolerance, absoluteTolerance, &r0, &r, &it)); break;
      case JACOBI: check(fvmgpu_jacobi_solve(h, sys, nMaxIterations, relativeTolerance, absoluteTolerance, &r0, &r, &it)); break;
      case ILU0: check(fvmgpu_ilu0_solve(h, sys, nMaxIterations, relativeTolerance, absoluteTolerance, &r0, &r, &it, 0)); break;
      case BCGSTAB_ILU0:
        check(fvmgpu_bcgstab_ilu0_solve(h, sys, nMaxIterations, relativeTolerance, absoluteTolerance, &r0, &r, &it));
        break;
    }
    _totalIterations += it;
    preconditioner->download(ls, v);
    if (verbosity > 0) {
      std::cout << "0: [" << v.index.first->getName() << " : " << r0 << "]" << std::endl;
      std::cout << it << ": [" << v.index.first->getName() << " : " << r << "]" << std::endl;
    }
    return GpuAMG::norm(v, r0);
  }
  virtual void smooth(LinearSystem&) { throw CException("GpuKrylovLike: not usable as a preconditioner"); }
  virtual void cleanup() { if (preconditioner) preconditioner->cleanup(); }
  int getTotalIterations() const { return _totalIterations; }

 private:
  Kind _kind;
  int _totalIterations;
};
struct GpuCG : GpuKrylovLike { GpuCG() : GpuKrylovLike(CG_AMG) {} };
struct GpuJacobiSolver : GpuKrylovLike { GpuJacobiSolver() : GpuKrylovLike(JACOBI) {} };
struct GpuILU0Solver : GpuKrylovLike { GpuILU0Solver() : GpuKrylovLike(ILU0) {} };
struct GpuBCGStabILU0 : GpuKrylovLike { GpuBCGStabILU0() : GpuKrylovLike(BCGSTAB_ILU0) {} };

// ------------------------------------------------------------------------------------ assembly
// Device mirror of one reference Mesh + GeomFields, created once (geometry outlives the models).
class GpuMesh {
 public:
  GpuMesh(const Mesh& mesh, const GeomFields& geom) : _h(0), _mesh(mesh) {
    ensureInit();
    const StorageSite& cells = mesh.getCells();
    const StorageSite& faces = mesh.getFaces();
    const CRConnectivity& fc = mesh.getAllFaceCells();
    const CRConnectivity& cc = mesh.getCellCells();
    std::vector<int> gOff, gCnt, gId, gKind;
    foreach (const FaceGroupPtr fg, mesh.getAllFaceGroups()) {
      gOff.push_back(fg->site.getOffset());
      gCnt.push_back(fg->site.getCount());
      gId.push_back(fg->id);
      gKind.push_back(fg->groupType == "interior" ? FVMGPU_GROUP_INTERIOR
                      : fg->groupType == "interface" ? FVMGPU_GROUP_INTERFACE
                      : fg->groupType == "symmetry" ? FVMGPU_GROUP_SYMMETRY : FVMGPU_GROUP_BOUNDARY);
    }
    check(fvmgpu_mesh_create(&_h, mesh.getDimension(), cells.getSelfCount(), cells.getCount(), faces.getCount(),
                             (const int*)fc.getCol().getData(), (const int*)cc.getRow().getData(),
                             (const int*)cc.getCol().getData(), (int)gOff.size(), &gOff[0], &gCnt[0], &gId[0],
                             &gKind[0]));
    check(fvmgpu_mesh_set_geometry(_h, (const double*)geom.area[faces].getData(),
                                   (const double*)geom.areaMag[faces].getData(),
                                   (const double*)geom.coordinate[faces].getData(),
                                   (const double*)geom.coordinate[cells].getData(),
                                   (const double*)geom.volume[cells].getData(),
                                   (const int*)geom.ibType[cells].getData()));
  }
  ~GpuMesh() { if (_h) fvmgpu_mesh_destroy(_h); }
  fvmgpu_mesh_t handle() const { return _h; }
  const Mesh& mesh() const { return _mesh; }

 private:
  GpuMesh(const GpuMesh&);
  fvmgpu_mesh_t _h;
  const Mesh& _mesh;
};

struct GpuBC {  // one GenericBCS call per boundary group
  int groupId, kind;
  double p[4];
};

// Scalar transport linearization on the GPU (thermal / electrostatic potential / any
// CRMatrix<T,T,T> transport equation): gradient + diffusion + [convection] + source +
// [time derivative] + BCs + boundary elimination, results copied into the reference's arrays.
class GpuScalarLinearizer {
 public:
  explicit GpuScalarLinearizer(const GpuMesh& gm) : _gm(gm), _sys(0) { check(fvmgpu_system_create(&_sys, gm.handle())); }
  ~GpuScalarLinearizer() { if (_sys) fvmgpu_system_destroy(_sys); }

  void linearize(LinearSystem& ls, Field& varField, const Field& diffusivity, const Field& source,
                 const std::vector<GpuBC>& bcs, const fvmgpu_assemble_opts& opts, Field* gradientField = 0) {
    const StorageSite& cells = _gm.mesh().getCells();
    MultiField::ArrayIndex xi(&varField, &cells);
    ScalarMatrix& m = dynamic_cast<ScalarMatrix&>(ls.getMatrix().getMatrix(xi, xi));
    DArray& x = dynamic_cast<DArray&>(ls.getX()[xi]);
    DArray& b = dynamic_cast<DArray&>(ls.getB()[xi]);
    const long long nt = cells.getCount();
    check(fvmgpu_system_set_field(_sys, FVMGPU_FIELD_X, (const double*)x.getData(), nt));
    check(fvmgpu_system_set_field(_sys, FVMGPU_FIELD_DIFFUSIVITY, (const double*)diffusivity[cells].getData(), nt));
    check(fvmgpu_system_set_field(_sys, FVMGPU_FIELD_SOURCE, (const double*)source[cells].getData(), nt));
    for (size_t i = 0; i < bcs.size(); i++) check(fvmgpu_system_set_bc(_sys, bcs[i].groupId, bcs[i].kind, bcs[i].p, 4, 0));
    check(fvmgpu_assemble(_sys, &opts));
    check(fvmgpu_download_system(_sys, (double*)m.getDiag().getData(), (double*)m.getOffDiag().getData(),
                                 (double*)b.getData(), 0));
    check(fvmgpu_system_get_field(_sys, FVMGPU_FIELD_X, (double*)x.getData(), nt));
    if (gradientField)
      check(fvmgpu_system_get_field(_sys, FVMGPU_FIELD_GRADIENT, (double*)(*gradientField)[cells].getData(), 3 * nt));
  }
  fvmgpu_system_t handle() { return _sys; }

  // One outer iteration with the linear system resident on the device from assembly to update
  // (XModel::Impl::advance's loop body, F/ThermalModel_impl.h:428-449): fields up, linearize + initSolve,
  // solver.solve, postSolve + updateSolution, the new field (and the boundary fluxes, over all faces) down.
  // Returns the initial residual 1-norm (what LinearSolver::solve returns).
  double iterateOnDevice(GpuAMG& solver, Field& varField, const Field& diffusivity, const Field& source,
                         const std::vector<GpuBC>& bcs, const fvmgpu_assemble_opts& opts, double* boundaryFlux = 0,
                         int* iterations = 0) {
    const StorageSite& cells = _gm.mesh().getCells();
    DArray& x = dynamic_cast<DArray&>(varField[cells]);
    const long long nt = cells.getCount();
    check(fvmgpu_system_set_field(_sys, FVMGPU_FIELD_X, (const double*)x.getData(), nt));
    check(fvmgpu_system_set_field(_sys, FVMGPU_FIELD_DIFFUSIVITY, (const double*)diffusivity[cells].getData(), nt));
    check(fvmgpu_system_set_field(_sys, FVMGPU_FIELD_SOURCE, (const double*)source[cells].getData(), nt));
    for (size_t i = 0; i < bcs.size(); i++) check(fvmgpu_system_set_bc(_sys, bcs[i].groupId, bcs[i].kind, bcs[i].p, 4, 0));
    check(fvmgpu_assemble(_sys, &opts));
    const double r0 = solver.solveOnDevice(_sys, iterations);
    check(fvmgpu_amg_cleanup(solver.handle()));
    check(fvmgpu_post_solve_update(_sys));
    check(fvmgpu_system_get_field(_sys, FVMGPU_FIELD_X, (double*)x.getData(), nt));
    if (boundaryFlux)
      check(fvmgpu_system_get_field(_sys, FVMGPU_FIELD_BFLUX, boundaryFlux, _gm.mesh().getFaces().getCount()));
    return r0;
  }

 private:
  const GpuMesh& _gm;
  fvmgpu_system_t _sys;
};

// ------------------------------------------------------------------------------------ FlowModel
// FlowModel<double>::advance on the device. The reference FlowModel object stays the owner of the boundary-condition
// map, the material map and the options (scripts keep writing fmodel.getBCMap()[id].bcType = ... as before) and of
// the FlowFields arrays; its init() has been called. Solvers: GpuAMG objects (AMG cycles, the reference default) --
// set `momentumSolver` / `pressureSolver` or leave the defaults of FlowModelOptions (rel 1e-1, 20 iterations).
class GpuFlowModel {
 public:
  GpuFlowModel(const GpuMesh& gm, FlowModel<double>& model, FlowFields& fields)
      : _gm(gm), _model(model), _f(fields), _flow(0), _niters(0), _haveInitial(false) {
    check(fvmgpu_flow_create(&_flow, gm.handle()));
    momentumSolver.relativeTolerance = 1e-1; momentumSolver.nMaxIterations = 20; momentumSolver.verbosity = 0;
    pressureSolver.relativeTolerance = 1e-1; pressureSolver.nMaxIterations = 20; pressureSolver.verbosity = 0;
    for (int k = 0; k < 3; k++) _mNorm0[k] = 0;
    _cNorm0 = 0;
  }
  ~GpuFlowModel() { if (_flow) fvmgpu_flow_destroy(_flow); }
  GpuAMG momentumSolver, pressureSolver;

  // after FlowModel::init(): fields and boundary conditions to the device, default face mass fluxes and continuity
  // residual computed there (F/FlowModel_impl.h:222-340) and mirrored back
  void init() {
    upload(false);
    check(fvmgpu_flow_init(_flow));
    const StorageSite& cells = _gm.mesh().getCells();
    const StorageSite& faces = _gm.mesh().getFaces();
    get(FVMGPU_FLOW_MASS_FLUX, _f.massFlux[faces], faces.getCount());
    get(FVMGPU_FLOW_CONT_RESID, _f.continuityResidual[cells], cells.getCount());
    _niters = 0;
    _haveInitial = false;
  }

  // FlowModel::advance (F/FlowModel_impl.h:1433-1471): true when both normalised residuals are below the tolerances
  bool advance(int niter) {
    FlowModelOptions<double>& o = _model.getOptions();
    fvmgpu_flow_opts fo;
    fo.momentumURF = o["momentumURF"]; fo.pressureURF = o["pressureURF"];
    fo.transient = o.transient ? 1 : 0; fo.time_order = o.timeDiscretizationOrder; fo.dt = o["timeStep"];
    fo.correctVelocity = o.correctVelocity ? 1 : 0;
    fo.operatingPressure = o["operatingPressure"]; fo.operatingTemperature = o["operatingTemperature"];
    fo.molecularWeight = o["molecularWeight"]; fo.incompressible = o.incompressible ? 1 : 0;
    momentumSolver.syncOptions();
    pressureSolver.syncOptions();
    upload(true);
    bool converged = false;
    for (int n = 0; n < niter; n++) {
      double mNorm[3] = {0, 0, 0}, cNorm = 0;
      int mIts[3] = {0, 0, 0}, cIts = 0;
      check(fvmgpu_flow_assemble_momentum(_flow, &fo));                                        // solveMomentum :730-770
      check(fvmgpu_flow_solve_momentum(_flow, momentumSolver.handle(), 0, 0, 0.0, 0.0, mNorm, mIts));
      check(fvmgpu_amg_cleanup(momentumSolver.handle()));
      check(fvmgpu_flow_assemble_continuity(_flow, &fo));                                      // solveContinuity :1410-1430
      check(fvmgpu_flow_solve_continuity(_flow, pressureSolver.handle(), 0, 0, 0.0, 0.0, &fo, &cNorm, &cIts));
      check(fvmgpu_amg_cleanup(pressureSolver.handle()));
      if (!_haveInitial) {
        for (int k = 0; k < 3; k++) _mNorm0[k] = mNorm[k];
        _cNorm0 = cNorm;
        _haveInitial = true;
      }
      if (_niters < 5) {   // setMax, :1441-1445
        for (int k = 0; k < 3; k++) _mNorm0[k] = std::max(_mNorm0[k], mNorm[k]);
        _cNorm0 = std::max(_cNorm0, cNorm);
      }
      double mr[3], cr = _cNorm0 > 0 ? cNorm / _cNorm0 : 0.0, mag = 0;
      for (int k = 0; k < 3; k++) { mr[k] = _mNorm0[k] > 0 ? mNorm[k] / _mNorm0[k] : 0.0; mag += mr[k] * mr[k]; }
      lastMomentumNorm[0] = mNorm[0]; lastMomentumNorm[1] = mNorm[1]; lastMomentumNorm[2] = mNorm[2];
      lastContinuityNorm = cNorm;
      if (o.printNormalizedResiduals)
        std::cout << _niters << ": [" << _f.velocity.getName() << " : [" << mr[0] << " " << mr[1] << " " << mr[2] << "]];["
                  << _f.pressure.getName() << " : " << cr << "]" << std::endl;
      else
        std::cout << _niters << ": [" << _f.velocity.getName() << " : [" << mNorm[0] << " " << mNorm[1] << " " << mNorm[2]
                  << "]];[" << _f.pressure.getName() << " : " << cNorm << "]" << std::endl;
      _niters++;
      // Vector::operator< compares magnitudes (F/Vector.h:169-172)
      if (std::sqrt(mag) < o.momentumTolerance && cr < o.continuityTolerance) { converged = true; break; }
    }
    download();
    return converged;
  }
  double lastMomentumNorm[3], lastContinuityNorm;
  fvmgpu_flow_t handle() { return _flow; }

 private:
  void set(int field, const ArrayBase& a, long long n) { check(fvmgpu_flow_set_field(_flow, field, (const double*)a.getData(), n)); }
  void get(int field, ArrayBase& a, long long n) { check(fvmgpu_flow_get_field(_flow, field, (double*)a.getData(), n)); }
  void upload(bool withFlux) {
    const Mesh& mesh = _gm.mesh();
    const StorageSite& cells = mesh.getCells();
    const StorageSite& faces = mesh.getFaces();
    const long long nt = cells.getCount(), nf = faces.getCount();
    set(FVMGPU_FLOW_VELOCITY, _f.velocity[cells], 3 * nt);
    set(FVMGPU_FLOW_PRESSURE, _f.pressure[cells], nt);
    set(FVMGPU_FLOW_FACE_PRESSURE, _f.pressure[faces], nf);
    set(FVMGPU_FLOW_DENSITY, _f.density[cells], nt);
    set(FVMGPU_FLOW_VISCOSITY, _f.viscosity[cells], nt);
    if (withFlux) {
      set(FVMGPU_FLOW_MASS_FLUX, _f.massFlux[faces], nf);
      set(FVMGPU_FLOW_CONT_RESID, _f.continuityResidual[cells], nt);
    }
    FlowModelOptions<double>& o = _model.getOptions();
    if (o.transient) {
      set(FVMGPU_FLOW_VELOCITY_N1, _f.velocityN1[cells], 3 * nt);
      if (o.timeDiscretizationOrder > 1) set(FVMGPU_FLOW_VELOCITY_N2, _f.velocityN2[cells], 3 * nt);
    }
    foreach (const FaceGroupPtr fgPtr, mesh.getBoundaryFaceGroups()) {
      const FaceGroup& fg = *fgPtr;
      FlowBC<double>& bc = *_model.getBCMap()[fg.id];
      double p[5] = {bc["specifiedXVelocity"], bc["specifiedYVelocity"], bc["specifiedZVelocity"], bc["specifiedPressure"],
                     bc["accomodationCoefficient"]};
      int kind;
      if (bc.bcType == "NoSlipWall") kind = FVMGPU_FLOWBC_NOSLIP_WALL;
      else if (bc.bcType == "SlipJump") { kind = FVMGPU_FLOWBC_SLIP_JUMP; p[3] = 0.0; }
      else if (bc.bcType == "Symmetry") kind = FVMGPU_FLOWBC_SYMMETRY;
      else if (bc.bcType == "VelocityBoundary") kind = FVMGPU_FLOWBC_VELOCITY;
      else if (bc.bcType == "PressureBoundary") kind = FVMGPU_FLOWBC_PRESSURE;
      else throw CException(bc.bcType + " not implemented for FlowModel");
      check(fvmgpu_flow_set_bc(_flow, fg.id, kind, p, 5));
    }
  }
  void download() {
    const StorageSite& cells = _gm.mesh().getCells();
    const StorageSite& faces = _gm.mesh().getFaces();
    const long long nt = cells.getCount(), nf = faces.getCount();
    get(FVMGPU_FLOW_VELOCITY, _f.velocity[cells], 3 * nt);
    get(FVMGPU_FLOW_PRESSURE, _f.pressure[cells], nt);
    get(FVMGPU_FLOW_FACE_PRESSURE, _f.pressure[faces], nf);
    get(FVMGPU_FLOW_MASS_FLUX, _f.massFlux[faces], nf);
    get(FVMGPU_FLOW_CONT_RESID, _f.continuityResidual[cells], nt);
    // (the gradient fields get their arrays from the reference's GradientModel on first use: mirrored when present)
    if (_f.pressureGradient.hasArray(cells)) get(FVMGPU_FLOW_PRESSURE_GRADIENT, _f.pressureGradient[cells], 3 * nt);
    if (_f.velocityGradient.hasArray(cells)) get(FVMGPU_FLOW_VELOCITY_GRADIENT, _f.velocityGradient[cells], 9 * nt);
  }
  GpuFlowModel(const GpuFlowModel&);
  const GpuMesh& _gm;
  FlowModel<double>& _model;
  FlowFields& _f;
  fvmgpu_flow_t _flow;
  int _niters;
  bool _haveInitial;
  double _mNorm0[3], _cNorm0;
};

}  // namespace fvmgpu_adaptor
#endif
