// fvm_b200 / libfvmgpu -- device-resident mirrors of the reference's data model.
//   Mesh    <- Mesh + StorageSite + CRConnectivity + GeomFields   (F/Mesh.h, F/StorageSite.h:18-112,
//              F/CRConnectivity.h:48-222, F/GeomFields.h)
//   System  <- LinearSystem + CRMatrix<T,T,T> + boundary flux rows (F/LinearSystem.h:11-64,
//              F/CRMatrix.h:86-1751, F/FluxJacobianMatrix.h, F/DiagonalMatrix.h)
// Layout in HBM (all FP64 / int32):
//   cellGeom[c]  = {cx,cy,cz,volume}      one 32 B sector per gathered neighbour
//   faceGeom[f]  = {Ax,Ay,Az,|A|}         one 32 B sector per face
//   cellState[c] = {gx,gy,gz,x}           written by the gradient kernel, gathered by assembly
//   CSR row/col as the reference's cellCells (diag implicit), entryFace[k] = 2*face+side for the
//   k-th off-diagonal entry (side 0: the row is the face's c0; 1: it is c1) -- the inverse of the
//   reference's pairToCol map (F/CRConnectivity.cpp:729-792), which makes the face loop a
//   row-parallel GATHER in the same summation order as the reference's face-order scatter.
#pragma once
#include "comm.cuh"
#include "../../include/fvmgpu.h"

namespace fvmgpu {

struct FaceGroup {
  int offset, count, id, kind;
};

struct Mesh {
  int dim = 3;
  int nSelf = 0, nTotal = 0, nFaces = 0;
  long long nnz = 0;
  int nInteriorFaces = 0;
  std::vector<FaceGroup> groups;
  DBuf<int> faceCells;   // 2F
  DBuf<int> row;         // Nt+1
  DBuf<int> col;         // nnz
  DBuf<int> entryFace;   // nnz : 2*f+side
  DBuf<int> pairToCol;   // 2F  : (pos01,pos10)
  DBuf<int> faceGroupOf; // per boundary-ish face (f >= nInteriorFaces): group index
  DBuf<int> groupKindDev; // FVMGPU_GROUP_* per group
  bool hasGeometry = false;
  DBuf<double4> cellGeom;  // Nt
  DBuf<double4> faceGeom;  // F
  DBuf<double> bFaceCen;   // centroids of the boundary faces only, 3 * (F - nInteriorFaces) (SlipJump walls)
  DBuf<double> gradW;      // 3*nnz, SoA: wx[nnz], wy[nnz], wz[nnz]
  // halo (multi-GPU): StorageSite scatter/gather maps per neighbouring rank (F/StorageSite.h:58-84)
  Halo halo;
  std::vector<int> haloScatterHost, haloGatherHost;  // host copies (the AMG setup derives coarse maps from them)
};

// per boundary group GenericBCS entry, staged to shared memory by the assembly kernel
struct BcEntry {
  int offset, count;  // face range
  int kind;           // FVMGPU_BC_* or -1 (none set)
  int groupKind;      // FVMGPU_GROUP_*
  double p[4];
  const double* perFace;  // device pointer or null
};

struct System {
  Mesh* mesh = nullptr;  // null for raw systems
  // matrix pattern (aliases mesh buffers when mesh != null)
  int nSelf = 0, nTotal = 0;
  long long nnz = 0;
  const int* row = nullptr;
  const int* col = nullptr;
  DBuf<int> rawRow, rawCol;
  DBuf<double> diag, off, b, delta;
  DBuf<int> isBoundary;
  // model fields
  DBuf<double> x, diffusivity, source, faceFlux, xN1, xN2, density, contResid;
  bool hasFaceFlux = false, hasXN1 = false, hasXN2 = false;
  DBuf<double4> cellState;  // {gx,gy,gz,x}
  DBuf<double> aux3a, aux3b;  // ElectricModel: electric field / electron velocity (3*Nt AoS)
  bool gradientValid = false;
  DBuf<double> xGhostNew;   // staged Dirichlet values for ghost cells (Nt - nSelf)
  // boundary flux side system, indexed by (face - nInteriorFaces)
  DBuf<double> bflux, rflux, coeffL, coeffR;
  std::vector<BcEntry> bcs;           // one per face group (host copy)
  std::vector<DBuf<double>> bcPerFace;  // owning storage of per-face values
  DBuf<BcEntry> bcsDev;
  bool bcsDirty = true;
  unsigned long long version = 0;  // nextVersion() stamp of the matrix values: renewed by every assemble / matrix change
  unsigned long long patternVersion = 0;  // stamp of the CSR pattern (set once at creation)
  bool noHalo = false;             // replicated (merged coarse) system: solved without communication
};

}  // namespace fvmgpu
