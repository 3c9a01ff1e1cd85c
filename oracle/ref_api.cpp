// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the
// product path (fvm_b200/, libfvmgpu.so). Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference leg may load the library built from this file.
//
// This translation unit is OUR glue: a C ABI (for ctypes) around the UNMODIFIED reference
// classes compiled in place from /root/reference (no reference source is copied here).
// It lets the tests feed identical raw mesh arrays to the reference CPU implementation and
// to the CUDA path, and read back the reference's assembled CRMatrix / residual / solution.
//
// Reference entry points driven (all public API unless noted):
//   Mesh raw ctor                 F/Mesh.h:93-99, F/Mesh.cpp:132-247
//   FluentReader                  I/FluentReader.h:82-89
//   MeshMetricsCalculator::init   F/MeshMetricsCalculator_impl.h:1944-2041
//   ThermalModel<double>          F/ThermalModel.h:28-52, Impl F/ThermalModel_impl.h
//   FlowModel<double>             F/FlowModel.h:17-95, Impl F/FlowModel_impl.h (SIMPLE: solveMomentum :730-770,
//                                 discretizeContinuity :1394-1407, solveContinuity :1410-1430, advance :1433-1471)
//   AMG / BCGStab                 F/AMG.cpp:219-298, F/BCGStab.cpp:26-170
// `Impl` of the models is a private nested class; to read the assembled system at the same
// points as Impl::dumpMatrix (F/ThermalModel_impl.h:499-539) this TU (and only this TU)
// compiles the reference headers with private/protected opened up.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#define private public
#define protected public
#include <atype.h>
#include "AMG.h"
#include "BCGStab.h"
#include "ILU0Solver.h"
#include "CG.h"
#include "JacobiSolver.h"
#include "CRMatrix.h"
#include "FluentReader.h"
#include "GeomFields.h"
#include "Mesh.h"
#include "MeshMetricsCalculator.h"
#include "MeshMetricsCalculator_impl.h"
#include "ThermalFields.h"
#include "ThermalModel.h"
#include "ThermalModel_impl.h"
#include "FlowFields.h"
#include "FlowModel.h"
#include "FlowModel_impl.h"
#include "ElectricFields.h"
#include "ElectricModel.h"
#include "ElectricModel_impl.h"
#include "VacancyFields.h"
#include "VacancyModel.h"
#include "VacancyModel_impl.h"
#include "SpeciesFields.h"
#include "SpeciesModel.h"
#include "SpeciesModel_impl.h"
#undef private
#undef protected

template class MeshMetricsCalculator<double>;
template class ThermalModel<double>;
template class FlowModel<double>;
template class ElectricModel<double>;
template class SpeciesModel<double>;
template class VacancyModel<double>;

typedef Vector<double, 3> Vec3;
typedef Array<Vec3> Vec3Array;
typedef Array<double> DArray;
typedef Array<int> IArray;

static thread_local std::string g_err;
static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct RefMesh {
  std::shared_ptr<FluentReader> reader;
  std::vector<std::shared_ptr<Mesh>> owned;
  MeshList meshes;
  std::shared_ptr<GeomFields> geom;
  std::shared_ptr<MeshMetricsCalculator<double>> metrics;
  Mesh& mesh() { return *meshes[0]; }
};

struct SolverCfg {  // mirrors the public tunables of F/AMG.h:74-81 and F/LinearSolver.h:15-20
  int kind;         // 0 = AMG, 1 = BCGStab preconditioned by one AMG cycle, 2 = ILU0Solver,
                    // 3 = BCGStab preconditioned by ILU0Solver, 4 = CG preconditioned by AMG, 5 = JacobiSolver
  int nMaxIterations;
  int verbosity;
  double relativeTolerance;
  double absoluteTolerance;
  int maxCoarseLevels;
  int nPreSweeps;
  int nPostSweeps;
  int coarseGroupSize;
  double weightRatioThreshold;
  int cycleType;     // 0 V, 1 W, 2 F
  int smootherType;  // 0 GS, 1 Jacobi
};

struct RefSolver {
  std::shared_ptr<AMG> amg;
  std::shared_ptr<BCGStab> bcg;
  std::shared_ptr<ILU0Solver> ilu;
  std::shared_ptr<CG> cg;
  std::shared_ptr<JacobiSolver> jac;
  LinearSolver* top = nullptr;
};

static RefSolver make_solver(const SolverCfg& c) {
  RefSolver s;
  s.amg.reset(new AMG());
  s.amg->maxCoarseLevels = c.maxCoarseLevels;
  s.amg->nPreSweeps = c.nPreSweeps;
  s.amg->nPostSweeps = c.nPostSweeps;
  s.amg->coarseGroupSize = c.coarseGroupSize;
  s.amg->weightRatioThreshold = c.weightRatioThreshold;
  s.amg->cycleType = (AMG::CycleType)c.cycleType;
  s.amg->smootherType = (AMG::SmootherType)c.smootherType;
  if (c.kind == 0) {
    s.top = s.amg.get();
  } else if (c.kind == 2) {
    s.ilu.reset(new ILU0Solver());
    s.top = s.ilu.get();
  } else if (c.kind == 3) {
    s.ilu.reset(new ILU0Solver());
    s.ilu->verbosity = 0;
    s.bcg.reset(new BCGStab());
    s.bcg->preconditioner = s.ilu.get();
    s.top = s.bcg.get();
  } else if (c.kind == 4) {
    s.cg.reset(new CG());
    s.cg->preconditioner = s.amg.get();
    s.amg->verbosity = 0;
    s.top = s.cg.get();
  } else if (c.kind == 5) {
    s.jac.reset(new JacobiSolver());
    s.top = s.jac.get();
  } else {
    s.bcg.reset(new BCGStab());
    s.bcg->preconditioner = s.amg.get();
    s.amg->verbosity = 0;
    s.top = s.bcg.get();
  }
  s.top->nMaxIterations = c.nMaxIterations;
  s.top->verbosity = c.verbosity;
  s.top->relativeTolerance = c.relativeTolerance;
  s.top->absoluteTolerance = c.absoluteTolerance;
  return s;
}

// RAII capture of std::cout (the reference prints its convergence history there,
// F/AMG.cpp:239-271, F/ThermalModel_impl.h:443)
struct CoutCapture {
  std::ostringstream os;
  std::streambuf* old;
  // the reference's Vector<T,N> printer leaves std::cout in scientific notation for the rest of the process
  // (F/Vector.h:66): every capture starts from the default formatting, as a fresh reference process would
  std::ios_base::fmtflags flags;
  std::streamsize prec;
  CoutCapture() : old(std::cout.rdbuf(os.rdbuf())), flags(std::cout.flags()), prec(std::cout.precision()) {
    std::cout.unsetf(std::ios_base::floatfield);
    std::cout.precision(6);
  }
  ~CoutCapture() { std::cout.rdbuf(old); std::cout.flags(flags); std::cout.precision(prec); }
};

static void copy_text(const std::string& s, char* out, int cap) {
  if (!out || cap <= 0) return;
  int n = std::min<int>((int)s.size(), cap - 1);
  std::memcpy(out, s.data(), n);
  out[n] = 0;
}

struct RefVacancy {
  RefMesh* m;
  std::shared_ptr<VacancyFields> fields;
  std::shared_ptr<VacancyModel<double>> model;
  RefSolver solver;
};

struct RefSpecies {
  RefMesh* m;
  std::shared_ptr<SpeciesModel<double>> model;
  RefSolver solver;
};

struct RefThermal {
  RefMesh* m;
  std::shared_ptr<ThermalFields> fields;
  std::shared_ptr<ThermalModel<double>> model;
  RefSolver solver;
};

#define TRY try {
#define CATCH(rv)                      \
  }                                    \
  catch (const std::exception& e) {    \
    g_err = e.what();                  \
    return rv;                         \
  }                                    \
  catch (...) {                        \
    g_err = "unknown exception";       \
    return rv;                         \
  }

extern "C" {

const char* fvmref_last_error() { return g_err.c_str(); }

static void finish_mesh(RefMesh* rm) {
  rm->geom.reset(new GeomFields("geom"));
  rm->metrics.reset(new MeshMetricsCalculator<double>(*rm->geom, rm->meshes));
  rm->metrics->init();
  // single-rank identity local<->global cell maps: the -DFVM_PARALLEL build of FlowModel reads
  // them (setDirichlet, F/FlowModel_impl.h:943-968); MeshPartitioner fills them even for one part
  for (Mesh* mesh : rm->meshes) {
    if (!mesh->_localToGlobal) {
      const int n = mesh->getCells().getCount();
      std::shared_ptr<Array<int>> l2g(new Array<int>(n));
      for (int i = 0; i < n; i++) { (*l2g)[i] = i; mesh->_globalToLocal[i] = i; }
      mesh->_localToGlobal = l2g;
    }
  }
}

void* fvmref_mesh_from_cas(const char* path) {
  TRY RefMesh* rm = new RefMesh;
  rm->reader.reset(new FluentReader(path));
  rm->reader->readMesh();
  rm->meshes = rm->reader->getMeshList();
  finish_mesh(rm);
  return rm;
  CATCH(nullptr)
}

// Raw arrays exactly as the reference's raw Mesh ctor takes them (F/Mesh.h:93-99).
void* fvmref_mesh_from_raw(int dim, int nCells, int nNodes, const double* nodes, int nFaces,
                           const int* faceCells, const int* faceNodes, const int* faceNodeCount,
                           int nGroups, const int* faceGroupSize) {
  TRY RefMesh* rm = new RefMesh;
  Vec3Array coords(nNodes);
  for (int i = 0; i < nNodes; i++)
    for (int k = 0; k < 3; k++) coords[i][k] = nodes[3 * i + k];
  IArray fc(2 * nFaces), fnc(nFaces), fgs(nGroups);
  long nfn = 0;
  for (int f = 0; f < nFaces; f++) {
    fc[2 * f] = faceCells[2 * f];
    fc[2 * f + 1] = faceCells[2 * f + 1];
    fnc[f] = faceNodeCount[f];
    nfn += faceNodeCount[f];
  }
  IArray fn((int)nfn);
  for (long i = 0; i < nfn; i++) fn[(int)i] = faceNodes[i];
  for (int g = 0; g < nGroups; g++) fgs[g] = faceGroupSize[g];
  std::shared_ptr<Mesh> mesh(new Mesh(dim, nCells, coords, fc, fn, fnc, fgs));
  rm->owned.push_back(mesh);
  rm->meshes.push_back(mesh.get());
  finish_mesh(rm);
  return rm;
  CATCH(nullptr)
}

// same, with some boundary groups re-typed "symmetry" BEFORE the metrics are computed (the raw Mesh
// ctor makes every boundary group a "wall", F/Mesh.cpp:188-200): ghost centroids are reflected
// (F/MeshMetricsCalculator_impl.h:205-220) and gradients reflected (F/GradientModel.h:536-549)
void* fvmref_mesh_from_raw_sym(int dim, int nCells, int nNodes, const double* nodes, int nFaces,
                               const int* faceCells, const int* faceNodes, const int* faceNodeCount,
                               int nGroups, const int* faceGroupSize, int nSym, const int* symGroupIds) {
  TRY RefMesh* rm = new RefMesh;
  Vec3Array coords(nNodes);
  for (int i = 0; i < nNodes; i++)
    for (int k = 0; k < 3; k++) coords[i][k] = nodes[3 * i + k];
  IArray fc(2 * nFaces), fnc(nFaces), fgs(nGroups);
  long nfn = 0;
  for (int f = 0; f < nFaces; f++) {
    fc[2 * f] = faceCells[2 * f];
    fc[2 * f + 1] = faceCells[2 * f + 1];
    fnc[f] = faceNodeCount[f];
    nfn += faceNodeCount[f];
  }
  IArray fn((int)nfn);
  for (long i = 0; i < nfn; i++) fn[(int)i] = faceNodes[i];
  for (int g = 0; g < nGroups; g++) fgs[g] = faceGroupSize[g];
  std::shared_ptr<Mesh> mesh(new Mesh(dim, nCells, coords, fc, fn, fnc, fgs));
  for (const FaceGroupPtr fg : mesh->getBoundaryFaceGroups())
    for (int k = 0; k < nSym; k++)
      if (fg->id == symGroupIds[k]) fg->groupType = "symmetry";
  rm->owned.push_back(mesh);
  rm->meshes.push_back(mesh.get());
  finish_mesh(rm);
  return rm;
  CATCH(nullptr)
}

// same with group types given by code: 3 = "symmetry", 4 = "dielectric interface" (a group whose ghost cells stand for
// cells behind a thin dielectric layer: MeshMetricsCalculator leaves their centroid / volume alone,
// F/MeshMetricsCalculator_impl.h:208-209,445-446; DiffusionDiscretization takes its thin-layer branch on its faces,
// F/DiffusionDiscretization.h:97-151). Nothing in the reference tree creates such a group; a script would.
void* fvmref_mesh_from_raw_typed(int dim, int nCells, int nNodes, const double* nodes, int nFaces,
                                 const int* faceCells, const int* faceNodes, const int* faceNodeCount,
                                 int nGroups, const int* faceGroupSize, int nTyped, const int* groupIds, const int* codes) {
  TRY RefMesh* rm = new RefMesh;
  Vec3Array coords(nNodes);
  for (int i = 0; i < nNodes; i++)
    for (int k = 0; k < 3; k++) coords[i][k] = nodes[3 * i + k];
  IArray fc(2 * nFaces), fnc(nFaces), fgs(nGroups);
  long nfn = 0;
  for (int f = 0; f < nFaces; f++) {
    fc[2 * f] = faceCells[2 * f];
    fc[2 * f + 1] = faceCells[2 * f + 1];
    fnc[f] = faceNodeCount[f];
    nfn += faceNodeCount[f];
  }
  IArray fn((int)nfn);
  for (long i = 0; i < nfn; i++) fn[(int)i] = faceNodes[i];
  for (int g = 0; g < nGroups; g++) fgs[g] = faceGroupSize[g];
  std::shared_ptr<Mesh> mesh(new Mesh(dim, nCells, coords, fc, fn, fnc, fgs));
  for (const FaceGroupPtr fg : mesh->getBoundaryFaceGroups())
    for (int k = 0; k < nTyped; k++)
      if (fg->id == groupIds[k]) fg->groupType = codes[k] == 3 ? "symmetry" : (codes[k] == 4 ? "dielectric interface" : "wall");
  rm->owned.push_back(mesh);
  rm->meshes.push_back(mesh.get());
  finish_mesh(rm);
  return rm;
  CATCH(nullptr)
}
// overwrite centroid / volume of cells [first, first + count) in the reference's GeomFields (ghost cells of a
// "dielectric interface" group have no metrics of their own)
int fvmref_mesh_set_cell_geometry(void* h, int first, int count, const double* centroid, const double* volume) {
  TRY RefMesh* rm = (RefMesh*)h;
  const StorageSite& cells = rm->mesh().getCells();
  Vec3Array& cx = dynamic_cast<Vec3Array&>(rm->geom->coordinate[cells]);
  DArray& cv = dynamic_cast<DArray&>(rm->geom->volume[cells]);
  for (int c = 0; c < count; c++) {
    for (int k = 0; k < 3; k++) cx[first + c][k] = centroid[3 * c + k];
    cv[first + c] = volume[c];
  }
  return 0;
  CATCH(-1)
}

void fvmref_mesh_free(void* h) { delete (RefMesh*)h; }

// out[0..7] = dim, nCellsSelf, nCellsTotal, nFaces, nnz(cellCells), nFaceGroups(all), nNodes, meshId
int fvmref_mesh_sizes(void* h, int* out) {
  TRY Mesh& mesh = ((RefMesh*)h)->mesh();
  out[0] = mesh.getDimension();
  out[1] = mesh.getCells().getSelfCount();
  out[2] = mesh.getCells().getCount();
  out[3] = mesh.getFaces().getCount();
  out[4] = mesh.getCellCells().getCol().getLength();
  out[5] = (int)mesh.getAllFaceGroups().size();
  out[6] = mesh.getNodes().getCount();
  out[7] = mesh.getID();
  return 0;
  CATCH(-1)
}

// Mesh::getCellNodes (F/Mesh.cpp:425-451: cellFaces x faceNodes, then Cell<T>::orderCellFacesAndNodes, F/Cell.cpp:96-200):
// row[nCells+1] / col of the SELF cells. Pass col == NULL to get the row offsets only (col length = row[nCells]).
int fvmref_mesh_cell_nodes(void* h, int* row, int* col) {
  TRY Mesh& mesh = ((RefMesh*)h)->mesh();
  const CRConnectivity& cn = mesh.getCellNodes();
  const int n = mesh.getCells().getSelfCount();
  row[0] = 0;
  for (int c = 0; c < n; c++) row[c + 1] = row[c] + cn.getCount(c);
  if (col)
    for (int c = 0; c < n; c++)
      for (int k = 0; k < cn.getCount(c); k++) col[row[c] + k] = cn(c, k);
  return 0;
  CATCH(-1)
}

// Mesh::getNodeCoordinates(): xyz[3 * nNodes] in the mesh's own node numbering
int fvmref_mesh_node_coordinates(void* h, double* xyz) {
  TRY Mesh& mesh = ((RefMesh*)h)->mesh();
  const Array<Vector<double, 3>>& x = mesh.getNodeCoordinates();
  for (int i = 0; i < x.getLength(); i++)
    for (int d = 0; d < 3; d++) xyz[3 * i + d] = x[i][d];
  return 0;
  CATCH(-1)
}

// groupKind: 0 interior, 1 boundary, 2 interface.  pairToCol: F/CRConnectivity.cpp:729-792.
int fvmref_mesh_connectivity(void* h, int* faceCells, int* ccRow, int* ccCol, int* pairToCol,
                             int* groupOffset, int* groupCount, int* groupId, int* groupKind) {
  TRY Mesh& mesh = ((RefMesh*)h)->mesh();
  const CRConnectivity& fc = mesh.getAllFaceCells();
  const int nFaces = mesh.getFaces().getCount();
  for (int f = 0; f < nFaces; f++) {
    faceCells[2 * f] = fc(f, 0);
    faceCells[2 * f + 1] = fc(f, 1);
  }
  const CRConnectivity& cc = mesh.getCellCells();
  const IArray& row = cc.getRow();
  const IArray& col = cc.getCol();
  for (int i = 0; i < row.getLength(); i++) ccRow[i] = row[i];
  for (int i = 0; i < col.getLength(); i++) ccCol[i] = col[i];
  if (pairToCol) {
    const Array<Vector<int, 2>>& p2c = cc.getPairToColMapping(fc);
    for (int f = 0; f < nFaces; f++) {
      pairToCol[2 * f] = p2c[f][0];
      pairToCol[2 * f + 1] = p2c[f][1];
    }
  }
  int g = 0;
  for (const FaceGroupPtr fg : mesh.getAllFaceGroups()) {
    groupOffset[g] = fg->site.getOffset();
    groupCount[g] = fg->site.getCount();
    groupId[g] = fg->id;
    groupKind[g] = fg->groupType == "interior" ? 0 : (fg->groupType == "interface" ? 2 : (fg->groupType == "symmetry" ? 3 :
                   (fg->groupType == "dielectric interface" ? 4 : 1)));
    g++;
  }
  return 0;
  CATCH(-1)
}

int fvmref_mesh_geometry(void* h, double* faceArea, double* faceAreaMag, double* faceCentroid,
                         double* cellCentroid, double* cellVolume, int* ibType) {
  TRY RefMesh* rm = (RefMesh*)h;
  Mesh& mesh = rm->mesh();
  const StorageSite& cells = mesh.getCells();
  const StorageSite& faces = mesh.getFaces();
  const Vec3Array& fa = dynamic_cast<const Vec3Array&>(rm->geom->area[faces]);
  const DArray& fam = dynamic_cast<const DArray&>(rm->geom->areaMag[faces]);
  const Vec3Array& fx = dynamic_cast<const Vec3Array&>(rm->geom->coordinate[faces]);
  const Vec3Array& cx = dynamic_cast<const Vec3Array&>(rm->geom->coordinate[cells]);
  const DArray& cv = dynamic_cast<const DArray&>(rm->geom->volume[cells]);
  const IArray& ib = dynamic_cast<const IArray&>(rm->geom->ibType[cells]);
  const int nF = faces.getCount(), nC = cells.getCount();
  for (int f = 0; f < nF; f++) {
    for (int k = 0; k < 3; k++) {
      faceArea[3 * f + k] = fa[f][k];
      faceCentroid[3 * f + k] = fx[f][k];
    }
    faceAreaMag[f] = fam[f];
  }
  for (int c = 0; c < nC; c++) {
    for (int k = 0; k < 3; k++) cellCentroid[3 * c + k] = cx[c][k];
    cellVolume[c] = cv[c];
    if (ibType) ibType[c] = ib[c];
  }
  return 0;
  CATCH(-1)
}

// ---------------------------------------------------------------- VacancyModel (F/VacancyModel.h:18-55)

void* fvmref_vacancy_create(void* h) {
  TRY RefMesh* rm = (RefMesh*)h;
  RefVacancy* t = new RefVacancy;
  t->m = rm;
  t->fields.reset(new VacancyFields("vacancy"));
  t->model.reset(new VacancyModel<double>(*rm->geom, *t->fields, rm->meshes));
  return t;
  CATCH(nullptr)
}
void fvmref_vacancy_free(void* h) { delete (RefVacancy*)h; }
int fvmref_vacancy_set_bc(void* h, int id, const char* bcType, const char* var, double value) {
  TRY RefVacancy* t = (RefVacancy*)h;
  auto& bcMap = t->model->getBCMap();
  if (bcMap.find(id) == bcMap.end()) throw CException("no such boundary id");
  VacancyBC<double>& bc = *bcMap[id];
  if (bcType && bcType[0]) bc.bcType = bcType;
  if (var && var[0]) {
    auto pos = bc.find(var);
    if (pos == bc.end()) throw CException(std::string("unknown bc var ") + var);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
int fvmref_vacancy_set_vc(void* h, const char* var, double value) {
  TRY RefVacancy* t = (RefVacancy*)h;
  for (auto& kv : t->model->getVCMap()) {
    auto pos = kv.second->find(var);
    if (pos == kv.second->end()) throw CException(std::string("unknown vc var ") + var);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
int fvmref_vacancy_set_option(void* h, const char* name, double value) {
  TRY RefVacancy* t = (RefVacancy*)h;
  VacancyModelOptions<double>& o = t->model->getOptions();
  std::string n(name);
  if (n == "relativeTolerance") o.relativeTolerance = value;
  else if (n == "absoluteTolerance") o.absoluteTolerance = value;
  else if (n == "transient") o.transient = value != 0;
  else if (n == "timeDiscretizationOrder") o.timeDiscretizationOrder = (int)value;
  else if (n == "useCentralDifference") o.useCentralDifference = value != 0;
  else {
    auto pos = o.find(n);
    if (pos == o.end()) throw CException("unknown option " + n);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
int fvmref_vacancy_set_solver(void* h, const SolverCfg* cfg) {
  TRY RefVacancy* t = (RefVacancy*)h;
  t->solver = make_solver(*cfg);
  t->model->getOptions().linearSolver = t->solver.top;
  return 0;
  CATCH(-1)
}
int fvmref_vacancy_init(void* h) {
  TRY((RefVacancy*)h)->model->init();
  return 0;
  CATCH(-1)
}
// name: concentration, diffusioncoefficient, source, specificVaca, concentrationN1, concentrationN2 (cells),
//       convectionFlux (all faces)
double* fvmref_vacancy_field(void* h, const char* name, int* len) {
  TRY RefVacancy* t = (RefVacancy*)h;
  Mesh& mesh = t->m->mesh();
  const StorageSite& cells = mesh.getCells();
  std::string n(name);
  VacancyFields& vf = *t->fields;
  ArrayBase* a = nullptr;
  if (n == "concentration") a = &vf.concentration[cells];
  else if (n == "diffusioncoefficient") a = &vf.diffusioncoefficient[cells];
  else if (n == "source") a = &vf.source[cells];
  else if (n == "specificVaca") a = &vf.specificVaca[cells];
  else if (n == "concentrationN1") a = &vf.concentrationN1[cells];
  else if (n == "concentrationN2") a = &vf.concentrationN2[cells];
  else if (n == "convectionFlux") a = &vf.convectionFlux[mesh.getFaces()];
  else throw CException("unknown field " + n);
  if (len) *len = a->getDataSize() / (int)sizeof(double);
  return (double*)a->getData();
  CATCH(nullptr)
}
int fvmref_vacancy_advance(void* h, int niter, char* text, int textCap) {
  TRY RefVacancy* t = (RefVacancy*)h;
  CoutCapture cap;
  t->model->advance(niter);
  copy_text(cap.os.str(), text, textCap);
  return 0;
  CATCH(-1)
}
int fvmref_vacancy_update_time(void* h) {
  TRY((RefVacancy*)h)->model->updateTime();
  return 0;
  CATCH(-1)
}
int fvmref_vacancy_flux_integral(void* h, int groupId, double* out) {
  TRY RefVacancy* t = (RefVacancy*)h;
  *out = t->model->getVacaFluxIntegral(t->m->mesh(), groupId);
  return 0;
  CATCH(-1)
}

// ---------------------------------------------------------------- SpeciesModel (F/SpeciesModel.h:17-55)

void* fvmref_species_create(void* h, int nSpecies) {
  TRY RefMesh* rm = (RefMesh*)h;
  RefSpecies* t = new RefSpecies;
  t->m = rm;
  t->model.reset(new SpeciesModel<double>(*rm->geom, rm->meshes, nSpecies));
  return t;
  CATCH(nullptr)
}
void fvmref_species_free(void* h) { delete (RefSpecies*)h; }
int fvmref_species_set_bc(void* h, int species, int id, const char* bcType, const char* var, double value) {
  TRY RefSpecies* t = (RefSpecies*)h;
  auto& bcMap = t->model->getBCMap(species);
  if (bcMap.find(id) == bcMap.end()) throw CException("no such boundary id");
  SpeciesBC<double>& bc = *bcMap[id];
  if (bcType && bcType[0]) bc.bcType = bcType;
  if (var && var[0]) {
    auto pos = bc.find(var);
    if (pos == bc.end()) throw CException(std::string("unknown bc var ") + var);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
int fvmref_species_set_vc(void* h, int species, const char* var, double value) {
  TRY RefSpecies* t = (RefSpecies*)h;
  for (auto& kv : t->model->getVCMap(species)) {
    auto pos = kv.second->find(var);
    if (pos == kv.second->end()) throw CException(std::string("unknown vc var ") + var);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
int fvmref_species_set_option(void* h, const char* name, double value) {
  TRY RefSpecies* t = (RefSpecies*)h;
  SpeciesModelOptions<double>& o = t->model->getOptions();
  std::string n(name);
  if (n == "relativeTolerance") o.relativeTolerance = value;
  else if (n == "absoluteTolerance") o.absoluteTolerance = value;
  else if (n == "transient") o.transient = value != 0;
  else if (n == "timeDiscretizationOrder") o.timeDiscretizationOrder = (int)value;
  else if (n == "useCentralDifference") o.useCentralDifference = value != 0;
  else {
    auto pos = o.find(n);
    if (pos == o.end()) throw CException("unknown option " + n);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
int fvmref_species_set_solver(void* h, const SolverCfg* cfg) {
  TRY RefSpecies* t = (RefSpecies*)h;
  t->solver = make_solver(*cfg);
  t->model->getOptions().linearSolver = t->solver.top;
  return 0;
  CATCH(-1)
}
int fvmref_species_init(void* h) {
  TRY((RefSpecies*)h)->model->init();
  return 0;
  CATCH(-1)
}
// name: massFraction, diffusivity, source, massFractionN1, massFractionN2 (cells), convectionFlux (all faces)
double* fvmref_species_field(void* h, int species, const char* name, int* len) {
  TRY RefSpecies* t = (RefSpecies*)h;
  Mesh& mesh = t->m->mesh();
  const StorageSite& cells = mesh.getCells();
  std::string n(name);
  SpeciesFields& sf = t->model->getSpeciesFields(species);
  ArrayBase* a = nullptr;
  if (n == "massFraction") a = &sf.massFraction[cells];
  else if (n == "diffusivity") a = &sf.diffusivity[cells];
  else if (n == "source") a = &sf.source[cells];
  else if (n == "massFractionN1") a = &sf.massFractionN1[cells];
  else if (n == "massFractionN2") a = &sf.massFractionN2[cells];
  else if (n == "convectionFlux") a = &sf.convectionFlux[mesh.getFaces()];
  else throw CException("unknown field " + n);
  if (len) *len = a->getDataSize() / (int)sizeof(double);
  return (double*)a->getData();
  CATCH(nullptr)
}
int fvmref_species_advance(void* h, int niter, char* text, int textCap) {
  TRY RefSpecies* t = (RefSpecies*)h;
  CoutCapture cap;
  t->model->advance(niter);
  copy_text(cap.os.str(), text, textCap);
  return 0;
  CATCH(-1)
}
int fvmref_species_update_time(void* h) {
  TRY((RefSpecies*)h)->model->updateTime();
  return 0;
  CATCH(-1)
}
// getMassFluxIntegral / getAverageMassFraction / getMassFractionResidual: what = 0 / 1 / 2 (arg = face group id for 0)
int fvmref_species_query(void* h, int species, int what, int arg, double* out) {
  TRY RefSpecies* t = (RefSpecies*)h;
  if (what == 0) *out = t->model->getMassFluxIntegral(t->m->mesh(), arg, species);
  else if (what == 1) *out = t->model->getAverageMassFraction(t->m->mesh(), species);
  else *out = t->model->getMassFractionResidual(species);
  return 0;
  CATCH(-1)
}

// ---------------------------------------------------------------- ThermalModel

void* fvmref_thermal_create(void* h) {
  TRY RefMesh* rm = (RefMesh*)h;
  RefThermal* t = new RefThermal;
  t->m = rm;
  t->fields.reset(new ThermalFields("therm"));
  t->model.reset(new ThermalModel<double>(*rm->geom, *t->fields, rm->meshes));
  return t;
  CATCH(nullptr)
}
void fvmref_thermal_free(void* h) { delete (RefThermal*)h; }

// bcType "" leaves the type unchanged; var "" sets no value.
int fvmref_thermal_set_bc(void* h, int id, const char* bcType, const char* var, double value) {
  TRY RefThermal* t = (RefThermal*)h;
  auto& bcMap = t->model->getBCMap();
  if (bcMap.find(id) == bcMap.end()) throw CException("no such boundary id");
  ThermalBC<double>& bc = *bcMap[id];
  if (bcType && bcType[0]) bc.bcType = bcType;
  if (var && var[0]) {
    auto pos = bc.find(var);
    if (pos == bc.end()) throw CException(std::string("unknown bc var ") + var);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
int fvmref_thermal_set_vc(void* h, const char* var, double value) {
  TRY RefThermal* t = (RefThermal*)h;
  for (auto& kv : t->model->getVCMap()) {
    auto pos = kv.second->find(var);
    if (pos == kv.second->end()) throw CException(std::string("unknown vc var ") + var);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
// name in {initialTemperature,timeStep} (float vars) or
// {relativeTolerance,absoluteTolerance,transient,timeDiscretizationOrder,useCentralDifference}
int fvmref_thermal_set_option(void* h, const char* name, double value) {
  TRY RefThermal* t = (RefThermal*)h;
  ThermalModelOptions<double>& o = t->model->getOptions();
  std::string n(name);
  if (n == "relativeTolerance") o.relativeTolerance = value;
  else if (n == "absoluteTolerance") o.absoluteTolerance = value;
  else if (n == "transient") o.transient = value != 0;
  else if (n == "timeDiscretizationOrder") o.timeDiscretizationOrder = (int)value;
  else if (n == "useCentralDifference") o.useCentralDifference = value != 0;
  else {
    auto pos = o.find(n);
    if (pos == o.end()) throw CException("unknown option " + n);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
int fvmref_thermal_set_solver(void* h, const SolverCfg* cfg) {
  TRY RefThermal* t = (RefThermal*)h;
  t->solver = make_solver(*cfg);
  t->model->getOptions().linearSolver = t->solver.top;
  return 0;
  CATCH(-1)
}
int fvmref_thermal_init(void* h) {
  TRY((RefThermal*)h)->model->init();
  return 0;
  CATCH(-1)
}

// Raw pointer into the reference's host Array for a cell/face field (valid after init()).
// name: temperature, conductivity, source, specificHeat, temperatureN1, temperatureN2 (cells, double),
//       temperatureGradient (cells, 3 doubles), convectionFlux (all faces, double)
double* fvmref_thermal_field(void* h, const char* name, int* len) {
  TRY RefThermal* t = (RefThermal*)h;
  Mesh& mesh = t->m->mesh();
  const StorageSite& cells = mesh.getCells();
  std::string n(name);
  ThermalFields& tf = *t->fields;
  ArrayBase* a = nullptr;
  if (n == "temperature") a = &tf.temperature[cells];
  else if (n == "conductivity") a = &tf.conductivity[cells];
  else if (n == "source") a = &tf.source[cells];
  else if (n == "specificHeat") a = &tf.specificHeat[cells];
  else if (n == "temperatureN1") a = &tf.temperatureN1[cells];
  else if (n == "temperatureN2") a = &tf.temperatureN2[cells];
  else if (n == "temperatureGradient") a = &tf.temperatureGradient[cells];
  else if (n == "convectionFlux") a = &tf.convectionFlux[mesh.getFaces()];
  else throw CException("unknown field " + n);
  if (len) *len = a->getDataSize() / (int)sizeof(double);
  return (double*)a->getData();
  CATCH(nullptr)
}

// heatFlux array of one boundary group (length = faces in the group)
int fvmref_thermal_heat_flux(void* h, int groupId, double* out) {
  TRY RefThermal* t = (RefThermal*)h;
  for (const FaceGroupPtr fg : t->m->mesh().getBoundaryFaceGroups())
    if (fg->id == groupId) {
      const DArray& a = dynamic_cast<const DArray&>(t->fields->heatFlux[fg->site]);
      for (int i = 0; i < a.getLength(); i++) out[i] = a[i];
      return 0;
    }
  throw CException("no such boundary id");
  CATCH(-1)
}

// Assemble exactly as Impl::advance / Impl::dumpMatrix do (F/ThermalModel_impl.h:424-436,499-520)
// and copy out the cell-cell CRMatrix + b (+ x, which Dirichlet BCs modify, F/GenericBCS.h:104).
// stage 0: after linearize(), before initSolve();  stage 1: after initSolve() (boundary rows
// eliminated, F/LinearSystem.cpp:56).  isBoundary may be null.
int fvmref_thermal_assemble(void* h, int stage, double* diag, double* offdiag, double* b, double* x,
                            int* isBoundary, double* seconds) {
  TRY RefThermal* t = (RefThermal*)h;
  ThermalModel<double>::Impl& impl = *t->model->_impl;
  Mesh& mesh = t->m->mesh();
  const StorageSite& cells = mesh.getCells();
  double t0 = now_s();
  LinearSystem ls;
  impl.initLinearization(ls);
  ls.initAssembly();
  impl.linearize(ls);
  if (stage >= 1) ls.initSolve();
  double t1 = now_s();
  if (seconds) *seconds = t1 - t0;
  MultiField::ArrayIndex tIndex(&t->fields->temperature, &cells);
  typedef CRMatrix<double, double, double> M;
  M& m = dynamic_cast<M&>(ls.getMatrix().getMatrix(tIndex, tIndex));
  const DArray& d = m.getDiag();
  const DArray& od = m.getOffDiag();
  const DArray& bb = dynamic_cast<const DArray&>(ls.getB()[tIndex]);
  const DArray& xx = dynamic_cast<const DArray&>(ls.getX()[tIndex]);
  for (int i = 0; i < d.getLength(); i++) {
    if (diag) diag[i] = d[i];
    if (b) b[i] = bb[i];
    if (x) x[i] = xx[i];
    if (isBoundary) isBoundary[i] = m._isBoundary[i] ? 1 : 0;
  }
  if (offdiag)
    for (int i = 0; i < od.getLength(); i++) offdiag[i] = od[i];
  return 0;
  CATCH(-1)
}

// The reference's own advance() (F/ThermalModel_impl.h:424-456), stdout captured.
int fvmref_thermal_advance(void* h, int niter, char* text, int textCap, double* seconds) {
  TRY RefThermal* t = (RefThermal*)h;
  CoutCapture cap;
  double t0 = now_s();
  t->model->advance(niter);
  if (seconds) *seconds = now_s() - t0;
  copy_text(cap.os.str(), text, textCap);
  return 0;
  CATCH(-1)
}

// Same statement sequence as Impl::advance for ONE outer iteration, with a timer around each
// phase. times[0..4] = assemble (initLinearization+initAssembly+linearize+initSolve),
// solve (incl. hierarchy setup), postSolve+updateSolution, AMG total iterations (as double),
// initial residual 1-norm.
int fvmref_thermal_advance_timed(void* h, double* times, char* text, int textCap) {
  TRY RefThermal* t = (RefThermal*)h;
  ThermalModel<double>::Impl& impl = *t->model->_impl;
  Mesh& mesh = t->m->mesh();
  CoutCapture cap;
  double t0 = now_s();
  LinearSystem ls;
  impl.initLinearization(ls);
  ls.initAssembly();
  impl.linearize(ls);
  ls.initSolve();
  double t1 = now_s();
  int it0 = t->solver.amg ? t->solver.amg->getTotalIterations() : 0;
  MFRPtr rNorm(impl._options.getLinearSolver().solve(ls));
  double t2 = now_s();
  impl._options.getLinearSolver().cleanup();
  ls.postSolve();
  ls.updateSolution();
  double t3 = now_s();
  impl._niters++;
  times[0] = t1 - t0;
  times[1] = t2 - t1;
  times[2] = t3 - t2;
  times[3] = t->solver.amg ? t->solver.amg->getTotalIterations() - it0 : -1;
  MultiField::ArrayIndex tIndex(&t->fields->temperature, &mesh.getCells());
  const DArray& rn = dynamic_cast<const DArray&>((*rNorm)[t->fields->temperature]);
  times[4] = rn[0];
  copy_text(cap.os.str(), text, textCap);
  return 0;
  CATCH(-1)
}

// ---------------------------------------------------------------- stand-alone linear solve
// Builds a LinearSystem the way MMReader::getLS does (I/MMReader.cpp:79-184) from a CSR
// pattern with separate diagonal (row/col exclude the diagonal), sign convention
// r = b + A x (F/CRMatrix.h:407-426), then runs the reference solver. nGhost extra rows
// follow the nSelf interior rows (F/CRMatrix.h:308,414).
// ---------------------------------------------------------------- FlowModel (SIMPLE)
struct RefFlow {
  RefMesh* m;
  std::shared_ptr<FlowFields> fields;
  std::shared_ptr<FlowModel<double>> model;
  RefSolver momSolver, prSolver;
  std::shared_ptr<LinearSystem> keep;  // keeps the last assembled system alive
};

void* fvmref_flow_create(void* h) {
  TRY RefMesh* rm = (RefMesh*)h;
  RefFlow* t = new RefFlow;
  t->m = rm;
  t->fields.reset(new FlowFields("flow"));
  t->model.reset(new FlowModel<double>(*rm->geom, *t->fields, rm->meshes));
  return t;
  CATCH(nullptr)
}
void fvmref_flow_free(void* h) { delete (RefFlow*)h; }

int fvmref_flow_set_bc(void* h, int id, const char* bcType, const char* var, double value) {
  TRY RefFlow* t = (RefFlow*)h;
  auto& bcMap = t->model->getBCMap();
  if (bcMap.find(id) == bcMap.end()) throw CException("no such boundary id");
  FlowBC<double>& bc = *bcMap[id];
  if (bcType && bcType[0]) bc.bcType = bcType;
  if (var && var[0]) {
    auto pos = bc.find(var);
    if (pos == bc.end()) throw CException(std::string("unknown bc var ") + var);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
int fvmref_flow_set_vc(void* h, const char* var, double value) {
  TRY RefFlow* t = (RefFlow*)h;
  for (auto& kv : t->model->getVCMap()) {
    auto pos = kv.second->find(var);
    if (pos == kv.second->end()) throw CException(std::string("unknown vc var ") + var);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
int fvmref_flow_set_option(void* h, const char* name, double value) {
  TRY RefFlow* t = (RefFlow*)h;
  FlowModelOptions<double>& o = t->model->getOptions();
  std::string n(name);
  if (n == "momentumTolerance") o.momentumTolerance = value;
  else if (n == "continuityTolerance") o.continuityTolerance = value;
  else if (n == "transient") o.transient = value != 0;
  else if (n == "correctVelocity") o.correctVelocity = value != 0;
  else if (n == "timeDiscretizationOrder") o.timeDiscretizationOrder = (int)value;
  else if (n == "printNormalizedResiduals") o.printNormalizedResiduals = value != 0;
  else {
    auto pos = o.find(n);
    if (pos == o.end()) throw CException("unknown option " + n);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
int fvmref_flow_update_time(void* h) {   // FlowModel::updateTime, F/FlowModel_impl.h:351-370
  TRY RefFlow* t = (RefFlow*)h;
  t->model->updateTime();
  return 0;
  CATCH(-1)
}
int fvmref_flow_set_solver(void* h, int which, const SolverCfg* cfg) {
  TRY RefFlow* t = (RefFlow*)h;
  if (which == 0) { t->momSolver = make_solver(*cfg); t->model->getOptions().momentumLinearSolver = t->momSolver.top; }
  else { t->prSolver = make_solver(*cfg); t->model->getOptions().pressureLinearSolver = t->prSolver.top; }
  return 0;
  CATCH(-1)
}
int fvmref_flow_init(void* h) {
  TRY RefFlow* t = (RefFlow*)h;
  CoutCapture cap;
  t->model->init();
  return 0;
  CATCH(-1)
}
// VIEW of a reference host array (doubles); cell vectors are AoS (3 / 9 doubles per cell)
double* fvmref_flow_field(void* h, const char* name, int* len) {
  TRY RefFlow* t = (RefFlow*)h;
  Mesh& mesh = t->m->mesh();
  const StorageSite& cells = mesh.getCells();
  const StorageSite& faces = mesh.getFaces();
  FlowFields& f = *t->fields;
  FlowModel<double>::Impl& impl = *t->model->_impl;
  std::string n(name);
  ArrayBase* a = nullptr;
  int width = 1;
  if (n == "velocity") { a = &f.velocity[cells]; width = 3; }
  else if (n == "pressure") a = &f.pressure[cells];
  else if (n == "facePressure") a = &f.pressure[faces];
  else if (n == "massFlux") a = &f.massFlux[faces];
  else if (n == "density") a = &f.density[cells];
  else if (n == "viscosity") a = &f.viscosity[cells];
  else if (n == "continuityResidual") a = &f.continuityResidual[cells];
  else if (n == "pressureGradient") { a = &f.pressureGradient[cells]; width = 3; }
  else if (n == "velocityGradient") { a = &f.velocityGradient[cells]; width = 9; }
  else if (n == "momAp") { if (!impl._momApField) throw CException("momAp not available"); a = &(*impl._momApField)[cells]; width = 3; }
  else if (n == "previousVelocity") { if (!impl._previousVelocity) throw CException("previousVelocity not available"); a = &(*impl._previousVelocity)[cells]; width = 3; }
  else throw CException("unknown field " + n);
  if (len) *len = a->getLength() * width;
  return (double*)a->getData();
  CATCH(nullptr)
}
// initMomentumLinearization + initAssembly + linearizeMomentum + initSolve (F/FlowModel_impl.h:730-737)
int fvmref_flow_momentum_system(void* h, double* diag3, double* offdiag, double* b3) {
  TRY RefFlow* t = (RefFlow*)h;
  FlowModel<double>::Impl& impl = *t->model->_impl;
  Mesh& mesh = t->m->mesh();
  const StorageSite& cells = mesh.getCells();
  std::shared_ptr<LinearSystem> ls(new LinearSystem());
  impl.initMomentumLinearization(*ls);
  ls->initAssembly();
  impl.linearizeMomentum(*ls);
  ls->initSolve();
  MultiField::ArrayIndex vIndex(&t->fields->velocity, &cells);
  typedef CRMatrix<DiagonalTensor<double, 3>, double, Vec3> M;
  M& m = dynamic_cast<M&>(ls->getMatrix().getMatrix(vIndex, vIndex));
  const Array<DiagonalTensor<double, 3>>& d = m.getDiag();
  const DArray& od = m.getOffDiag();
  const Vec3Array& bb = dynamic_cast<const Vec3Array&>(ls->getB()[vIndex]);
  for (int i = 0; i < d.getLength(); i++)
    for (int k = 0; k < 3; k++) { diag3[3 * i + k] = d[i][k]; b3[3 * i + k] = bb[i][k]; }
  for (int i = 0; i < od.getLength(); i++) offdiag[i] = od[i];
  t->keep = ls;
  return 0;
  CATCH(-1)
}
// Impl::solveMomentum (:730-770): returns the initial residual 1-norms of the 3 components
int fvmref_flow_solve_momentum(void* h, double* rnorm3) {
  TRY RefFlow* t = (RefFlow*)h;
  FlowModel<double>::Impl& impl = *t->model->_impl;
  CoutCapture cap;
  MFRPtr r = impl.solveMomentum();
  const Vec3Array& a = dynamic_cast<const Vec3Array&>((*r)[t->fields->velocity]);
  for (int k = 0; k < 3; k++) rnorm3[k] = a[0][k];
  return 0;
  CATCH(-1)
}
// Impl::discretizeContinuity (:1394-1407): needs a preceding solve_momentum (momAp, previousVelocity);
// NB it also overwrites massFlux with the Rhie-Chow face fluxes
int fvmref_flow_continuity_system(void* h, double* diag, double* offdiag, double* b, int* isBoundary) {
  TRY RefFlow* t = (RefFlow*)h;
  FlowModel<double>::Impl& impl = *t->model->_impl;
  Mesh& mesh = t->m->mesh();
  const StorageSite& cells = mesh.getCells();
  std::shared_ptr<LinearSystem> ls = impl.discretizeContinuity();
  MultiField::ArrayIndex pIndex(&t->fields->pressure, &cells);
  typedef CRMatrix<double, double, double> M;
  M& m = dynamic_cast<M&>(ls->getMatrix().getMatrix(pIndex, pIndex));
  const DArray& d = m.getDiag();
  const DArray& od = m.getOffDiag();
  const DArray& bb = dynamic_cast<const DArray&>(ls->getB()[pIndex]);
  for (int i = 0; i < d.getLength(); i++) {
    diag[i] = d[i]; b[i] = bb[i];
    if (isBoundary) isBoundary[i] = m._isBoundary[i] ? 1 : 0;
  }
  for (int i = 0; i < od.getLength(); i++) offdiag[i] = od[i];
  t->keep = ls;
  return 0;
  CATCH(-1)
}
int fvmref_flow_solve_continuity(void* h, double* rnorm) {
  TRY RefFlow* t = (RefFlow*)h;
  FlowModel<double>::Impl& impl = *t->model->_impl;
  CoutCapture cap;
  MFRPtr r = impl.solveContinuity();
  const DArray& a = dynamic_cast<const DArray&>((*r)[t->fields->pressure]);
  rnorm[0] = a[0];
  return 0;
  CATCH(-1)
}
// FlowModel::advance (:1433-1471); returns 1 if converged, stdout (residual history) captured
int fvmref_flow_advance(void* h, int niter, char* text, int textCap, double* seconds) {
  TRY RefFlow* t = (RefFlow*)h;
  CoutCapture cap;
  double t0 = now_s();
  bool conv = t->model->advance(niter);
  if (seconds) *seconds = now_s() - t0;
  copy_text(cap.os.str(), text, textCap);
  return conv ? 1 : 0;
  CATCH(-1)
}

// ---------------------------------------------------------------- ElectricModel
struct RefElectric {
  RefMesh* m;
  std::shared_ptr<ElectricFields> fields;
  std::shared_ptr<ElectricModel<double>> model;
  RefSolver esSolver, ctSolver;
};
void* fvmref_electric_create(void* h) {
  TRY RefMesh* rm = (RefMesh*)h;
  RefElectric* t = new RefElectric;
  t->m = rm;
  t->fields.reset(new ElectricFields("elec"));
  t->model.reset(new ElectricModel<double>(*rm->geom, *t->fields, rm->meshes));
  return t;
  CATCH(nullptr)
}
void fvmref_electric_free(void* h) { delete (RefElectric*)h; }
int fvmref_electric_set_bc(void* h, int id, const char* bcType, const char* var, double value) {
  TRY RefElectric* t = (RefElectric*)h;
  auto& bcMap = t->model->getBCMap();
  if (bcMap.find(id) == bcMap.end()) throw CException("no such boundary id");
  ElectricBC<double>& bc = *bcMap[id];
  if (bcType && bcType[0]) bc.bcType = bcType;
  if (var && var[0]) {
    auto pos = bc.find(var);
    if (pos == bc.end()) throw CException(std::string("unknown bc var ") + var);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
// kind 0: VC var, 1: option (float var or named flag), 2: constant
int fvmref_electric_set(void* h, int kind, const char* name, double value) {
  TRY RefElectric* t = (RefElectric*)h;
  std::string n(name);
  if (kind == 0) {
    for (auto& kv : t->model->getVCMap()) {
      auto pos = kv.second->find(n);
      if (pos == kv.second->end()) throw CException("unknown vc var " + n);
      pos->second.constant = value;
    }
  } else if (kind == 1) {
    ElectricModelOptions<double>& o = t->model->getOptions();
    if (n == "electrostaticsTolerance") o.electrostaticsTolerance = value;
    else if (n == "chargetransportTolerance") o.chargetransportTolerance = value;
    else if (n == "electrostatics_enable") o.electrostatics_enable = value != 0;
    else if (n == "chargetransport_enable") o.chargetransport_enable = value != 0;
    else if (n == "transient_enable") o.transient_enable = value != 0;
    else if (n == "drift_enable") o.drift_enable = value != 0;
    else if (n == "timeDiscretizationOrder") o.timeDiscretizationOrder = (int)value;
    else if (n == "printNormalizedResiduals") o.printNormalizedResiduals = value != 0;
    else {
      auto pos = o.find(n);
      if (pos == o.end()) throw CException("unknown option " + n);
      pos->second.constant = value;
    }
  } else {
    ElectricModelConstants<double>& c = t->model->getConstants();
    auto pos = c.find(n);
    if (pos == c.end()) throw CException("unknown constant " + n);
    pos->second.constant = value;
  }
  return 0;
  CATCH(-1)
}
int fvmref_electric_set_solver(void* h, int which, const SolverCfg* cfg) {
  TRY RefElectric* t = (RefElectric*)h;
  if (which == 0) { t->esSolver = make_solver(*cfg); t->model->getOptions().electrostaticsLinearSolver = t->esSolver.top; }
  else { t->ctSolver = make_solver(*cfg); t->model->getOptions().chargetransportLinearSolver = t->ctSolver.top; }
  return 0;
  CATCH(-1)
}
int fvmref_electric_init(void* h) {
  TRY RefElectric* t = (RefElectric*)h;
  CoutCapture cap;
  t->model->init();
  return 0;
  CATCH(-1)
}
double* fvmref_electric_field(void* h, const char* name, int* len) {
  TRY RefElectric* t = (RefElectric*)h;
  Mesh& mesh = t->m->mesh();
  const StorageSite& cells = mesh.getCells();
  const StorageSite& faces = mesh.getFaces();
  ElectricFields& f = *t->fields;
  std::string n(name);
  ArrayBase* a = nullptr;
  int width = 1;
  if (n == "potential") a = &f.potential[cells];
  else if (n == "dielectric_constant") a = &f.dielectric_constant[cells];
  else if (n == "total_charge") a = &f.total_charge[cells];
  else if (n == "potential_gradient") { a = &f.potential_gradient[cells]; width = 3; }
  else if (n == "electric_field") { a = &f.electric_field[cells]; width = 3; }
  else if (n == "electron_velocity") { a = &f.electron_velocity[cells]; width = 3; }
  else if (n == "convectionFlux") a = &f.convectionFlux[faces];
  else if (n == "charge") { a = &f.charge[cells]; width = 3; }
  else if (n == "chargeN1") { a = &f.chargeN1[cells]; width = 3; }
  else throw CException("unknown field " + n);
  if (len) *len = a->getLength() * width;
  return (double*)a->getData();
  CATCH(nullptr)
}
// initElectroStaticsLinearization + initAssembly + linearizeElectroStatics + initSolve (:377-385)
int fvmref_electric_potential_system(void* h, double* diag, double* offdiag, double* b) {
  TRY RefElectric* t = (RefElectric*)h;
  ElectricModel<double>::Impl& impl = *t->model->_impl;
  const StorageSite& cells = t->m->mesh().getCells();
  LinearSystem ls;
  impl.initElectroStaticsLinearization(ls);
  ls.initAssembly();
  impl.linearizeElectroStatics(ls);
  ls.initSolve();
  MultiField::ArrayIndex pIndex(&t->fields->potential, &cells);
  typedef CRMatrix<double, double, double> M;
  M& m = dynamic_cast<M&>(ls.getMatrix().getMatrix(pIndex, pIndex));
  const DArray& bb = dynamic_cast<const DArray&>(ls.getB()[pIndex]);
  for (int i = 0; i < m.getDiag().getLength(); i++) { diag[i] = m.getDiag()[i]; b[i] = bb[i]; }
  for (int i = 0; i < m.getOffDiag().getLength(); i++) offdiag[i] = m.getOffDiag()[i];
  return 0;
  CATCH(-1)
}
// ElectricModel::advance (:929-998); stdout captured; returns 1 if electrostatics converged
int fvmref_electric_advance(void* h, int niter, char* text, int textCap) {
  TRY RefElectric* t = (RefElectric*)h;
  CoutCapture cap;
  bool conv = t->model->advance(niter);
  copy_text(cap.os.str(), text, textCap);
  return conv ? 1 : 0;
  CATCH(-1)
}
int fvmref_electric_update_time(void* h) {
  TRY RefElectric* t = (RefElectric*)h;
  CoutCapture cap;
  t->model->updateTime();
  return 0;
  CATCH(-1)
}

// levelSizes (cap 64 ints, -1 terminated) receives the coarse level sizes.
int fvmref_linsolve(int nSelf, int nGhost, const int* row, const int* col, const double* diag,
                    const double* offdiag, const double* b, const SolverCfg* cfg, double* x,
                    double* rnorm0, int* iters, int* levelSizes, char* text, int textCap,
                    double* seconds) {
  TRY const int n = nSelf + nGhost;
  StorageSite site(nSelf, nGhost);
  CRConnectivity cm(site, site);
  cm.initCount();
  for (int i = 0; i < n; i++) cm.addCount(i, row[i + 1] - row[i]);
  cm.finishCount();
  typedef CRMatrix<double, double, double> M;
  std::shared_ptr<M> m(new M(cm));
  Field field("test");
  MultiField::ArrayIndex rowI(&field, &site);
  std::shared_ptr<DArray> xPtr(new DArray(n));
  xPtr->zero();
  LinearSystem ls;
  ls.getX().addArray(rowI, xPtr);
  ls.getMatrix().addMatrix(rowI, rowI, m);
  ls.initAssembly();
  DArray& d = m->getDiag();
  DArray& od = m->getOffDiag();
  for (int i = 0; i < n; i++) {
    d[i] = diag[i];
    for (int k = row[i]; k < row[i + 1]; k++) {
      int pos = cm.add(i, col[k]);
      od[pos] = offdiag[k];
    }
  }
  cm.finishAdd();
  DArray& bb = dynamic_cast<DArray&>(ls.getB()[rowI]);
  for (int i = 0; i < n; i++) bb[i] = b[i];
  ls.initSolve();
  RefSolver s = make_solver(*cfg);
  CoutCapture cap;
  double t0 = now_s();
  MFRPtr rn = s.top->solve(ls);
  if (seconds) *seconds = now_s() - t0;
  if (iters) *iters = s.amg->getTotalIterations();
  if (levelSizes) {
    int k = 0;
    for (auto& cl : s.amg->_coarseLinearSystems)
      if (k < 63) levelSizes[k++] = cl->getMatrix().getLocalSize();
    levelSizes[k] = -1;
  }
  s.top->cleanup();
  ls.postSolve();
  const DArray& delta = dynamic_cast<const DArray&>(ls.getDelta()[rowI]);
  for (int i = 0; i < n; i++) x[i] = delta[i];
  if (rnorm0) *rnorm0 = dynamic_cast<const DArray&>((*rn)[field])[0];
  copy_text(cap.os.str(), text, textCap);
  return 0;
  CATCH(-1)
}

}  // extern "C"
