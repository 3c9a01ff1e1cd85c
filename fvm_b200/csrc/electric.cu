// fvm_b200 / libfvmgpu -- ElectricModel-specific kernels (FP64, sm_100a).
//
// The electrostatics step of ElectricModel (F/ElectricModel_impl.h:377-410, 552-767) is the scalar
// transport path of assemble.cu (DiffusionDiscretization<T,T,T> with dielectric_constant as the
// diffusivity, SourceDiscretization with total_charge, GenericBCS Dirichlet / Neumann / dielectric
// boundary). What is specific to the model and lives here:
//   updateElectricField     :1001-1020   E = -grad(potential)
//   updateElectronVelocity  :1023-1048   v = -mobility E, limited to the saturation velocity
//   updateConvectionFlux    :1050-1092   face flux 0.5 (v0.A + v1.A); boundary faces v0.A, 0 on symmetry
// The drift step (DriftDiscretization, F/DriftDiscretization.h:83-112) convects only component nTrap
// of the charge vector with that face flux; with the tunnelling / capture / emission source models
// out of scope the 3x3 blocks of the charge system stay diagonal, so each charge component is a
// scalar transport equation assembled by the same fused kernel (convection + time derivative +
// zero-Dirichlet BCs) -- see fvm_b200/models.py ElectricModelA.
// Compiled with -fmad=false (same IEEE operation order as the reference).
#include "solver.cuh"

namespace fvmgpu {

struct ElectricFieldRows {  // E = -potential_gradient for every cell (incl. ghosts)
  const double4* state; double* E;
  FVM_DEV void operator()(long long i) const {
    const double4 s = state[i];
    E[3 * i] = -s.x; E[3 * i + 1] = -s.y; E[3 * i + 2] = -s.z;
  }
};
struct ElectronVelocityRows {  // F/ElectricModel_impl.h:1037-1045
  const double* E; double mobility, vsat; double* vel;
  FVM_DEV void operator()(long long i) const {
    const double e0 = E[3 * i], e1 = E[3 * i + 1], e2 = E[3 * i + 2];
    const double v0 = mobility * e0, v1 = mobility * e1, v2 = mobility * e2;
    const double magV = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
    if (magV < vsat) {
      vel[3 * i] = -mobility * e0; vel[3 * i + 1] = -mobility * e1; vel[3 * i + 2] = -mobility * e2;
    } else {
      const double magE = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
      vel[3 * i] = -vsat * (e0 / magE); vel[3 * i + 1] = -vsat * (e1 / magE); vel[3 * i + 2] = -vsat * (e2 / magE);
    }
  }
};
// NB (reference quirk, reproduced for parity): the boundary loop (:1070-1088) indexes
// mesh.getAllFaceCells() with the GROUP-LOCAL face index, i.e. boundary face k of a group takes the
// velocity of cell c0 of face k of the whole mesh (an interior face), not of its own neighbour cell.
struct DriftFluxFaces {  // F/ElectricModel_impl.h:1064-1088
  int nInteriorFaces; const int* faceCells; const int* faceGroupOf; const int* groupIsSymmetry; const int* groupOffset;
  const int* groupKind;
  const double4* faceGeom; const double* vel; double* flux;
  FVM_DEV double vdotA(int c, const double4 fg) const {
    double s = 0.0;
    s += vel[3 * (size_t)c] * fg.x; s += vel[3 * (size_t)c + 1] * fg.y; s += vel[3 * (size_t)c + 2] * fg.z;
    return s;
  }
  FVM_DEV void operator()(long long ff) const {
    const int f = (int)ff;
    const int c0 = faceCells[2 * f], c1 = faceCells[2 * f + 1];
    const double4 fg = faceGeom[f];
    // the reference's first loop runs over ALL faces, its second one over the BOUNDARY groups only: a partition
    // interface face keeps the two-sided average (c1 = the ghost cell holding the owner's velocity)
    if (f >= nInteriorFaces && groupKind[faceGroupOf[f - nInteriorFaces]] != FVMGPU_GROUP_INTERFACE) {
      const int g = faceGroupOf[f - nInteriorFaces];
      const int cq = faceCells[2 * (f - groupOffset[g])];  // see NB above
      flux[f] = groupIsSymmetry[g] ? 0.0 : vdotA(cq, fg);
      return;
    }
    flux[f] = 0.5 * (vdotA(c0, fg) + vdotA(c1, fg));
  }
};

// potential system -> E (host copy optional)
void electricField(System* s, double* E_host) {
  requireReady();
  if (!s->mesh) fail("electric_field: needs a mesh system");
  computeGradient(s);
  const size_t nt = (size_t)s->nTotal;
  if (s->aux3a.n < 3 * nt) s->aux3a.alloc(3 * nt);
  parallelFor((long long)nt, ElectricFieldRows{s->cellState.p, s->aux3a.p});
  if (E_host) s->aux3a.download(E_host, 3 * nt);
}

// electron velocity from the potential system's E, drift face flux into the charge system's FACE_FLUX
void electricDriftFlux(System* potential, System* charge, double mobility, double vsat, int nSym, const int* symGroupIds,
                       double* vel_host) {
  requireReady();
  Mesh* m = potential->mesh;
  if (!m || charge->mesh != m) fail("electric_drift_flux: both systems must live on the same mesh");
  const size_t nt = (size_t)m->nTotal;
  if (potential->aux3a.n < 3 * nt) fail("electric_drift_flux: call fvmgpu_electric_field first");
  if (potential->aux3b.n < 3 * nt) potential->aux3b.alloc(3 * nt);
  parallelFor((long long)nt, ElectronVelocityRows{potential->aux3a.p, mobility, vsat, potential->aux3b.p});
  std::vector<int> isSym(m->groups.size(), 0);
  for (size_t g = 0; g < m->groups.size(); g++)
    for (int k = 0; k < nSym; k++)
      if (m->groups[g].id == symGroupIds[k] && m->groups[g].kind != FVMGPU_GROUP_INTERIOR) isSym[g] = 1;
  std::vector<int> gOff(m->groups.size(), 0);
  for (size_t g = 0; g < m->groups.size(); g++) gOff[g] = m->groups[g].offset;
  DBuf<int> isSymDev, gOffDev;
  isSymDev.upload(isSym.data(), isSym.size());
  gOffDev.upload(gOff.data(), gOff.size());
  if (charge->faceFlux.n < (size_t)m->nFaces) charge->faceFlux.alloc((size_t)m->nFaces);
  parallelFor(m->nFaces, DriftFluxFaces{m->nInteriorFaces, m->faceCells.p, m->faceGroupOf.p, isSymDev.p, gOffDev.p, m->groupKindDev.p, m->faceGeom.p,
                                        potential->aux3b.p, charge->faceFlux.p});
  charge->hasFaceFlux = true;
  if (vel_host) potential->aux3b.download(vel_host, 3 * nt);
  streamSync();
}

}  // namespace fvmgpu
