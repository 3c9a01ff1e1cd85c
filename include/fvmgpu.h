/* fvmgpu.h -- C ABI of libfvmgpu.so: the B200 (sm_100a, FP64) implementation of MEMOSA-FVM's
 * per-iteration hot path (cell-centred face-loop assembly + AMG / BCGStab linear solve).
 *
 * Plain C: opaque handles, raw pointers and sizes only. Unless a name ends in `_d` every
 * pointer is HOST memory owned by the caller; the library copies what it needs during the
 * call and never retains or frees a host pointer. All floating point is IEEE double, all
 * indices int32. Every function returns 0 on success, non-zero on failure (the message is
 * in fvmgpu_last_error()); no C++ exception crosses this boundary. There is NO CPU
 * fallback: without a CUDA device every compute entry point fails.
 *
 * Each entry point cites the reference interface it stands in for
 * (paths relative to the reference tree; F/ = src/fvm/src/modules/fvmbase/).
 * One caller thread per handle; one process per GPU (F/ has no threading either).
 */
#ifndef FVMGPU_H_
#define FVMGPU_H_

#ifdef __cplusplus
extern "C" {
#endif

#define FVMGPU_VERSION 100

typedef struct fvmgpu_mesh_s* fvmgpu_mesh_t;     /* Mesh + StorageSite + CRConnectivity + GeomFields */
typedef struct fvmgpu_system_s* fvmgpu_system_t; /* LinearSystem + CRMatrix<T,T,T> + its fields      */
typedef struct fvmgpu_solver_s* fvmgpu_solver_t; /* AMG (optionally wrapped by BCGStab)              */

/* ---- face-group kinds (FaceGroup::groupType, F/Mesh.h:28-43) ---- */
enum {
  FVMGPU_GROUP_INTERIOR = 0,
  FVMGPU_GROUP_BOUNDARY = 1,  /* "wall", "velocity-inlet", "pressure-outlet", ...           */
  FVMGPU_GROUP_INTERFACE = 2, /* partition interface (F/Mesh.cpp:267-274)                    */
  FVMGPU_GROUP_SYMMETRY = 3,  /* boundary whose ghost gradient is reflected (F/GradientModel.h:536-549) */
  /* "dielectric interface": the ghost cells of the group stand for cells behind a thin dielectric layer. Their
   * centroid / volume are the CALLER's (F/MeshMetricsCalculator_impl.h:208-209,445-446 skips the group; so
   * fvmgpu_mesh_compute_geometry rejects it: use fvmgpu_mesh_set_geometry), no gradient is copied into them
   * (F/GradientModel.h:538-539), and the diffusion coefficient of their faces is the thin-layer one
   * (F/DiffusionDiscretization.h:97-151, fvmgpu_assemble_opts::interface_thickness) */
  FVMGPU_GROUP_DIELECTRIC_INTERFACE = 4
};

/* ---- boundary-condition kinds = GenericBCS::apply*BC (F/GenericBCS.h) ---- */
enum {
  FVMGPU_BC_DIRICHLET = 0,     /* applyDirichletBC     :77-115   p[0] = value                     */
  FVMGPU_BC_NEUMANN = 1,       /* applyNeumannBC       :129-157  p[0] = specified flux (per area) */
  FVMGPU_BC_EXTRAPOLATION = 2, /* applyExtrapolationBC :180-212                                   */
  FVMGPU_BC_CONVECTIVE = 3,    /* applyConvectionBC    :214-245  p[0] = h, p[1] = Xinf            */
  FVMGPU_BC_RADIATIVE = 4,     /* applyRadiationBC     :253-288  p[0] = emissivity, p[1] = Xinf   */
  FVMGPU_BC_MIXED = 5,         /* applyMixedBC         :290-323  p[0] = h, p[1] = emissivity, p[2] = Xinf */
  FVMGPU_BC_INTERFACE = 6,     /* applyInterfaceBC     :325-356                                   */
  /* ThermalModel "SpecifiedTemperature" with a convecting flux present: per face
   * extrapolation where flux>0 else Dirichlet (F/ThermalModel_impl.h:313-331); p[0] = value */
  FVMGPU_BC_DIRICHLET_OR_OUTFLOW = 7,
  /* applyDielectricInterfaceBC :367-407: flux = -h (x - Xinf) |A| + source |A| / 2;
   * p[0] = Xinf (per-face values override it: FloatValEvaluator bT[f], F/ElectricModel_impl.h:734-745),
   * p[1] = h = dielectric_constant / dielectric_thickness, p[2] = source */
  FVMGPU_BC_DIELECTRIC_INTERFACE = 8
};

/* ---- named per-system device fields ---- */
enum {
  FVMGPU_FIELD_X = 0,          /* unknown, nCellsTotal (ls.getX(): aliases the model field)            */
  FVMGPU_FIELD_DIFFUSIVITY = 1,/* nCellsTotal  (DiffusionDiscretization::_diffusivityField)           */
  FVMGPU_FIELD_SOURCE = 2,     /* nCellsTotal  (SourceDiscretization::_sourceField)                   */
  FVMGPU_FIELD_FACE_FLUX = 3,  /* nFaces       (ConvectionDiscretization::_convectingFluxField)       */
  FVMGPU_FIELD_X_N1 = 4,       /* nCellsTotal  (TimeDerivativeDiscretization::_varN1Field)            */
  FVMGPU_FIELD_X_N2 = 5,       /* nCellsTotal  (..::_varN2Field)                                      */
  FVMGPU_FIELD_DENSITY = 6,    /* nCellsTotal  (..::_densityField; rho*cp for ThermalModel)           */
  FVMGPU_FIELD_CONT_RESID = 7, /* nCellsTotal  (ConvectionDiscretization::_continuityResidualField)   */
  FVMGPU_FIELD_GRADIENT = 8,   /* 3*nCellsTotal AoS (Gradient<T>, F/Gradient.h:199) -- output         */
  FVMGPU_FIELD_BFLUX = 9,      /* nFaces: boundary flux unknowns (heatFlux, ls.getX()[fIndex]) -- output*/
  FVMGPU_FIELD_DELTA = 10,     /* nCellsTotal  (ls.getDelta())                                        */
  FVMGPU_FIELD_B = 11,         /* nCellsTotal  (ls.getB())                                            */
  FVMGPU_FIELD_BFLUX_BOUNDARY = 12, /* nFaces - nInteriorFaces: FIELD_BFLUX without the interior faces' zeros -- output */
  FVMGPU_FIELD_COUNT = 13
};

/* ---- assembly options: which Discretization objects of the list are present
 *      (Linearizer::linearize, F/Linearizer.cpp:16-33; list built in
 *      F/ThermalModel_impl.h:236-296) ---- */
typedef struct {
  int diffusion;        /* DiffusionDiscretization<T,T,T>       F/DiffusionDiscretization.h:65-232 */
  int convection;       /* ConvectionDiscretization (upwind)    F/ConvectionDiscretization.h:166-199;
                           2 = useCentralDifference branch :119-164 (incl. its x[c0]+x[c0] quirk)     */
  int source;           /* SourceDiscretization                 F/SourceDiscretization.h:54-57      */
  int time_order;       /* 0 steady, 1 / 2 = TimeDerivativeDiscretization order  :149-155 / :102-108 */
  double dt;            /* time step for time_order>0                                               */
  double underrelax;    /* >0: Underrelaxer diag /= urf          F/Underrelaxer.h:49-52; 0 = none    */
  int apply_bcs;        /* run the per-group GenericBCS table set by fvmgpu_system_set_bc           */
  int eliminate_boundary; /* LinearSystem::initSolve -> CRMatrix::eliminateBoundaryEquations
                             F/LinearSystem.cpp:39-64, F/CRMatrix.h:899-944,1064-1085               */
  double interface_thickness; /* DiffusionDiscretization::_thickness (0 in ThermalModel, dielectric_thickness in
                             ElectricModel, F/ElectricModel_impl.h:563): only read on FVMGPU_GROUP_DIELECTRIC_INTERFACE faces */
} fvmgpu_assemble_opts;

/* ---- solver options = public tunables of AMG (F/AMG.h:74-81, defaults F/AMG.cpp:14-22)
 *      and LinearSolver (F/LinearSolver.h:15-20) ---- */
enum { FVMGPU_CYCLE_V = 0, FVMGPU_CYCLE_W = 1, FVMGPU_CYCLE_F = 2 };
enum { FVMGPU_SMOOTHER_GAUSS_SEIDEL = 0, /* multicolour GS: forward = colours ascending, reverse = descending */
       FVMGPU_SMOOTHER_JACOBI = 1 };
typedef struct {
  int nMaxIterations;
  int verbosity;
  double relativeTolerance;
  double absoluteTolerance;
  int maxCoarseLevels;
  int nPreSweeps;
  int nPostSweeps;
  int coarseGroupSize;
  double weightRatioThreshold;
  int cycleType;
  int smootherType;
} fvmgpu_amg_opts;
void fvmgpu_amg_default_opts(fvmgpu_amg_opts* o); /* the reference defaults */

/* ---- library ---- */
int fvmgpu_init(int device);                /* select device, create streams; idempotent */
int fvmgpu_shutdown(void);
const char* fvmgpu_last_error(void);        /* CException message equivalent (F/CException.h:16-21) */
int fvmgpu_version(void);
int fvmgpu_device_info(char* name, int cap, int* sm_count, double* mem_gb);
int fvmgpu_synchronize(void);

/* device-timeline stopwatch: CUDA events recorded on the library's compute stream */
int fvmgpu_timer_start(int slot);           /* slot 0..15 */
int fvmgpu_timer_stop(int slot, double* ms);/* records, synchronizes the stop event, returns elapsed */
/* counts of kernels the library launched / bytes copied since init (gpu_launches claim) */
int fvmgpu_counters(long long* kernel_launches, long long* h2d_bytes, long long* d2h_bytes);
int fvmgpu_flush_l2(void);                  /* writes a 256 MiB scratch buffer (> 126 MB L2) */
/* page-locked host memory for the caller's field arrays (the reference's Array<T> storage, F/Array.h:22-54, would be
 * allocated with this instead of new[]): copies to and from it run at PCIe speed and need no staging */
int fvmgpu_host_alloc(void** out, unsigned long long bytes);
int fvmgpu_host_free(void* p);
/* per-launch profiler: between begin and end every kernel launch is bracketed by CUDA events on
 * the compute stream; end returns one record per (kernel class, rows): names[i*nameStride..] is
 * the (mangled) functor name, rows[i] the logical threads, launches[i], ms[i] the summed duration */
int fvmgpu_profile_begin(void);
int fvmgpu_profile_end(int cap, char* names, int nameStride, long long* rows, long long* launches,
                       double* ms, int* count);

/* ---- mesh: what Mesh/StorageSite/CRConnectivity/GeomFields hold for one mesh ----
 * faceCells[2*nFaces]        Mesh::getAllFaceCells()                  F/Mesh.h, F/CRConnectivity.h:48-222
 * cellCellsRow/Col           Mesh::getCellCells() (diag implicit)     F/Mesh.cpp:479-492
 * groups                     Mesh::getAllFaceGroups(): contiguous face ranges, interior first
 * nCellsSelf/nCellsTotal     StorageSite::getSelfCount()/getCount()   F/StorageSite.h:18-112
 */
int fvmgpu_mesh_create(fvmgpu_mesh_t* out, int dim, int nCellsSelf, int nCellsTotal, int nFaces,
                       const int* faceCells, const int* cellCellsRow, const int* cellCellsCol,
                       int nGroups, const int* groupOffset, const int* groupCount,
                       const int* groupId, const int* groupKind);
/* GeomFields arrays (AoS Vector<double,3> as the reference stores them, F/Vector.h:229):
 * area[3F], areaMag[F], coordinate[faces] (may be NULL), coordinate[cells][3Nt], volume[Nt],
 * ibType[Nt] (may be NULL; anything but IBTYPE_FLUID=-1 is rejected: IB is out of scope).
 * Also builds the least-squares gradient weights (GradientModel::getLeastSquaresGradientMatrix2D/3D,
 * F/GradientModel.h:126-436), cached per mesh like the reference's static map (:453-468). */
int fvmgpu_mesh_set_geometry(fvmgpu_mesh_t mesh, const double* faceArea, const double* faceAreaMag,
                             const double* faceCentroid, const double* cellCentroid,
                             const double* cellVolume, const int* ibType);
/* MeshMetricsCalculator<T>::init on the device (F/MeshMetricsCalculator_impl.h:58-120, 128-236, 238-304,
 * 373-389, 392-460): face areas / magnitudes / centroids, cell centroids and volumes from the node
 * coordinates and the face-node connectivity (CSR: faceNodeOffsets[nFaces+1], faceNodes), installed in
 * the mesh exactly like fvmgpu_mesh_set_geometry (incl. the gradient weights) and copied to the host
 * arrays that are not NULL (the GeomFields arrays of the reference). Unpartitioned meshes only:
 * interface ghosts take the remote cell's geometry from the partitioner. */
int fvmgpu_mesh_compute_geometry(fvmgpu_mesh_t mesh, int nNodes, const double* nodeCoords, const int* faceNodeOffsets,
                                 const int* faceNodes, double* faceArea, double* faceAreaMag, double* faceCentroid,
                                 double* cellCentroid, double* cellVolume);
/* StorageSite scatter/gather maps per neighbour rank (F/StorageSite.h:58-84); CSR-style offsets */
int fvmgpu_mesh_set_halo(fvmgpu_mesh_t mesh, int nNeigh, const int* peerRank, const int* scatterOff,
                         const int* scatterIdx, const int* gatherOff, const int* gatherIdx);
int fvmgpu_mesh_destroy(fvmgpu_mesh_t mesh);
/* parity hooks */
int fvmgpu_mesh_download_pair_to_col(fvmgpu_mesh_t mesh, int* pairToCol /*2F*/); /* F/CRConnectivity.cpp:729-792 */
int fvmgpu_mesh_download_gradient_weights(fvmgpu_mesh_t mesh, double* coeffs /*3*nnz AoS*/); /* GradientMatrix::_coeffs */

/* ---- linear system on a mesh: CRMatrix<T,T,T>(cellCells) + x,b,delta,residual + per boundary
 *      group FluxJacobianMatrix/DiagonalMatrix rows (F/ThermalModel_impl.h:181-234) ---- */
int fvmgpu_system_create(fvmgpu_system_t* out, fvmgpu_mesh_t mesh);
/* stand-alone system from a CSR pattern with separate diagonal, r = b + A x convention
 * (MMReader::getLS, I/MMReader.cpp:79-184). Rows nSelf..nSelf+nGhost-1 are ghost rows. */
int fvmgpu_system_create_raw(fvmgpu_system_t* out, int nSelf, int nGhost, const int* row,
                             const int* col, const double* diag, const double* offdiag,
                             const double* b);
int fvmgpu_system_destroy(fvmgpu_system_t sys);
int fvmgpu_system_set_field(fvmgpu_system_t sys, int field, const double* host, long long n);
int fvmgpu_system_fill_field(fvmgpu_system_t sys, int field, double value);
int fvmgpu_system_get_field(fvmgpu_system_t sys, int field, double* host, long long n);
/* GenericBCS for one boundary group id; perFace (may be NULL) overrides p[0] per face
 * (FloatValEvaluator with a Field, F/FloatVarDict.h:100-140) */
int fvmgpu_system_set_bc(fvmgpu_system_t sys, int groupId, int bcKind, const double* p, int np,
                         const double* perFace);

/* GradientModel<T>::compute (F/GradientModel.h:472-602): gradient of FIELD_X into FIELD_GRADIENT,
 * boundary ghost cells copy (or reflect, symmetry groups) the neighbour's gradient */
int fvmgpu_compute_gradient(fvmgpu_system_t sys);
/* initAssembly + Linearizer::linearize + BC loop (+ initSolve elimination): ONE fused
 * gather kernel over rows, deterministic (no atomics), every output written once. */
int fvmgpu_assemble(fvmgpu_system_t sys, const fvmgpu_assemble_opts* opts);
/* parity hook: CRMatrix::getDiag()/getOffDiag() and ls.getB() (F/CRMatrix.h:856-865) */
int fvmgpu_download_system(fvmgpu_system_t sys, double* diag /*Nt*/, double* offdiag /*nnz*/,
                           double* b /*Nt*/, int* isBoundary /*Nt, may be NULL*/);

/* ---- solvers: LinearSolver::{solve,smooth,cleanup} (F/LinearSolver.h:11-31) ---- */
int fvmgpu_amg_create(fvmgpu_solver_t* out, const fvmgpu_amg_opts* opts);
int fvmgpu_amg_set_opts(fvmgpu_solver_t s, const fvmgpu_amg_opts* opts);
/* AMG::solve (F/AMG.cpp:219-282): builds the hierarchy when the system changed (createCoarseLevels
 * :149-210), cycles until ||r||_1 < abs or ratio < rel. rnorm0 = initial residual 1-norm (the
 * value the reference returns), iters = cycles run. delta is left in FIELD_DELTA. */
int fvmgpu_amg_solve(fvmgpu_solver_t s, fvmgpu_system_t sys, double* rnorm0, double* rnorm,
                     int* iters);
int fvmgpu_amg_smooth(fvmgpu_solver_t s, fvmgpu_system_t sys); /* AMG::smooth :285-298 */
int fvmgpu_amg_cleanup(fvmgpu_solver_t s);                      /* AMG::cleanup :212-217 */
int fvmgpu_amg_destroy(fvmgpu_solver_t s);
/* hierarchy report: sizes[l] rows and nnzs[l] off-diagonal entries per level (level 0 = finest) */
int fvmgpu_amg_levels(fvmgpu_solver_t s, int cap, int* nLevels, long long* sizes, long long* nnzs,
                      int* colours);
/* storage report: colBytes[l] = bytes of column index the level's row kernels read per stored entry -- 4 for plain
 * int32 columns, 2.125 where the 16-bit offsets (one int32 base per 32 entries) are in use, in between when only
 * some 32-row slices of the level qualify. bench.py's byte model of a row pass uses it. */
int fvmgpu_amg_level_col_bytes(fvmgpu_solver_t s, int cap, double* colBytes);
/* smoother ordering of one level (inspection / tests): nat[r] = row of level-row r in the numbering the
 * level was built from (level 0: the system's rows), colourStart[0..nColours] = first level-row of each
 * colour class. Rows of one class are relaxed concurrently, so no stored a_ij may join two of them. */
int fvmgpu_amg_level_order(fvmgpu_solver_t s, int level, long long cap, int* nat, int* nColours,
                           long long* colourStart);
/* Verification hook (tests; not a fast path). While a callback is registered, every AMG hierarchy of this process
 * takes its aggregates from the CALLER instead of the library's parallel pairing: per level the callback receives
 * the level matrix as a CSR in natural row numbering (diag separate, entries in stored order, isBoundary[i] != 0:
 * row not to be coarsened), fills coarseIndex[0..nRows-1] (-1 = not coarsened) and returns the number of aggregates
 * (< 0: error). Levels built this way are smoothed in natural row order (dependency wavefronts), i.e. exactly a
 * sequential forward / reverse Gauss-Seidel. tests/ register the oracle's restatement of the reference's sequential
 * agglomeration (CRMatrix::createCoarsening, F/CRMatrix.h:468-586) to reproduce the reference's AMG goldens digit
 * for digit with this library's kernels. fn = NULL restores the default. Single rank only. */
typedef int (*fvmgpu_aggregate_fn)(void* user, int nRows, const int* row, const int* col, const double* diag,
                                   const double* offdiag, const int* isBoundary, int groupSize,
                                   double weightRatioThreshold, int* coarseIndex);
int fvmgpu_debug_set_aggregator(fvmgpu_aggregate_fn fn, void* user);
/* Measurement aid: with FVMGPU_TAIL_TRACE=1 in the environment the fused coarse-level V-cycle kernel stamps the GPU's
 * global timer after each of its barrier phases; this returns the stamps of its LAST launch (ns) with their tags
 * (level within the fused stretch << 8 | phase kind: 1 restriction, 2 residual, 3 prolongation, 0x10 | colour a
 * Gauss-Seidel pass, 0x20 a Jacobi pass). *n = entries written. */
int fvmgpu_debug_tail_trace(int cap, unsigned long long* times_ns, int* tags, int* n);
/* host wall clock of the last fvmgpu_amg_solve, split into the hierarchy build (0 when the hierarchy
 * was reused) and the cycle loop, in milliseconds */
int fvmgpu_amg_last_timing(fvmgpu_solver_t s, double* setup_ms, double* cycles_ms);
/* residual history of the last solve: out[0..n-1], n returned */
int fvmgpu_solver_history(fvmgpu_solver_t s, int cap, double* out, int* n);
/* BCGStab::solve (F/BCGStab.cpp:26-170) right-preconditioned by one AMG cycle of `precond`;
 * its own nMaxIterations / tolerances are passed here (LinearSolver fields of the BCGStab object) */
int fvmgpu_bcgstab_solve(fvmgpu_solver_t precond, fvmgpu_system_t sys, int nMaxIterations,
                         double relativeTolerance, double absoluteTolerance, double* rnorm0,
                         double* rnorm, int* iters);
/* the same with ILU0Solver::smooth as the preconditioner (T/PARALLEL_CAVITY_ILU0): x = U^-1 L^-1 (-r) with the
 * reference's own ILU(0) (CRMatrix::compute_ILU0 / lowerSolve / upperSolve, F/CRMatrix.h:1546-1715); `s` is any
 * solver handle, it keeps the factors */
int fvmgpu_bcgstab_ilu0_solve(fvmgpu_solver_t s, fvmgpu_system_t sys, int nMaxIterations, double relativeTolerance,
                              double absoluteTolerance, double* rnorm0, double* rnorm, int* iters);
/* ILU0Solver::solve (F/ILU0Solver.cpp:46-93): delta = U^-1 L^-1 (-b) per sweep, residual 1-norm test.
 * levels (optional): dependency levels of the factorisation (one kernel launch each) */
int fvmgpu_ilu0_solve(fvmgpu_solver_t s, fvmgpu_system_t sys, int nMaxIterations, double relativeTolerance,
                      double absoluteTolerance, double* rnorm0, double* rnorm, int* iters, int* levels);
/* CG::solve (F/CG.cpp:24-140): conjugate gradients preconditioned by one AMG cycle of `precond` */
int fvmgpu_cg_solve(fvmgpu_solver_t precond, fvmgpu_system_t sys, int nMaxIterations,
                    double relativeTolerance, double absoluteTolerance, double* rnorm0, double* rnorm,
                    int* iters);
/* JacobiSolver::solve (F/JacobiSolver.cpp:46-95): Jacobi passes on the finest level only; `s` is any
 * solver handle (its hierarchy is not used) */
int fvmgpu_jacobi_solve(fvmgpu_solver_t s, fvmgpu_system_t sys, int nMaxIterations,
                        double relativeTolerance, double absoluteTolerance, double* rnorm0, double* rnorm,
                        int* iters);
/* LinearSystem::postSolve + updateSolution (F/LinearSystem.cpp:250-269): back-substitute the
 * eliminated boundary rows, solve the boundary-flux rows, x += delta, flux += dflux */
int fvmgpu_post_solve_update(fvmgpu_system_t sys);

/* ---- ElectricModel: the electrostatics step is the scalar path above (potential = FIELD_X,
 *      dielectric_constant = FIELD_DIFFUSIVITY, total_charge = FIELD_SOURCE; SpecifiedPotential =
 *      DIRICHLET, SpecifiedPotentialFlux / Symmetry = NEUMANN, SpecialDielectricBoundary =
 *      CONVECTIVE with h = eps/thickness: applyDielectricInterfaceBC with zero source,
 *      F/GenericBCS.h:367-407). Model-specific updates (F/ElectricModel_impl.h:1001-1092): ---- */
/* updateElectricField: potential gradient (GradientModel::compute) and E = -grad; E_host[3*nCellsTotal] or NULL */
int fvmgpu_electric_field(fvmgpu_system_t potential, double* E_host);
/* updateElectronVelocity + updateConvectionFlux: v = -mobility E limited to vsat (velocity_host
 * [3*nCellsTotal] or NULL); the drift face flux lands in `charge`'s FIELD_FACE_FLUX (device to
 * device), zero on the listed symmetry boundary groups. Call fvmgpu_electric_field first. */
int fvmgpu_electric_drift_flux(fvmgpu_system_t potential, fvmgpu_system_t charge, double mobility, double vsat,
                               int nSymmetryGroups, const int* symmetryGroupIds, double* velocity_host);

/* ---- FlowModel (SIMPLE): momentum + pressure-correction hot path (F/FlowModel_impl.h:522-1471,
 *      F/FlowModelInterior.h, F/FlowModelVelocityBC.h, F/MomentumPressureGradientDiscretization.h).
 *      Boundary types: NoSlipWall, Symmetry, VelocityBoundary, PressureBoundary, SlipJump. On several
 *      ranks (fvmgpu_comm_init) every rank holds the model of its mesh part and makes the same sequence
 *      of calls. Vector cell fields are AoS (Vector<T,3>,
 *      F/Vector.h:229; Gradient<Vector<T,3>> = 9 doubles [direction][component], F/Gradient.h:199). ---- */
typedef struct fvmgpu_flow_s* fvmgpu_flow_t;  /* FlowFields of one mesh + the two linear systems */
enum {
  FVMGPU_FLOW_VELOCITY = 0,          /* 3*nCellsTotal  FlowFields::velocity                        */
  FVMGPU_FLOW_PRESSURE = 1,          /* nCellsTotal    FlowFields::pressure[cells]                 */
  FVMGPU_FLOW_DENSITY = 2,           /* nCellsTotal                                                */
  FVMGPU_FLOW_VISCOSITY = 3,         /* nCellsTotal                                                */
  FVMGPU_FLOW_MASS_FLUX = 4,         /* nFaces         FlowFields::massFlux                        */
  FVMGPU_FLOW_FACE_PRESSURE = 5,     /* nFaces         FlowFields::pressure[faces]                 */
  FVMGPU_FLOW_PRESSURE_GRADIENT = 6, /* 3*nCellsTotal                                              */
  FVMGPU_FLOW_VELOCITY_GRADIENT = 7, /* 9*nCellsTotal                                              */
  FVMGPU_FLOW_CONT_RESID = 8,        /* nCellsTotal    FlowFields::continuityResidual              */
  FVMGPU_FLOW_MOM_AP = 9,            /* 3*nCellsTotal  Impl::_momApField (momentum diagonal)       */
  FVMGPU_FLOW_PREV_VELOCITY = 10,    /* 3*nCellsTotal  Impl::_previousVelocity                     */
  FVMGPU_FLOW_VELOCITY_N1 = 11,      /* 3*nCellsTotal  (transient)                                 */
  FVMGPU_FLOW_VELOCITY_N2 = 12
};
enum {
  FVMGPU_FLOWBC_NOSLIP_WALL = 0, /* applyDirichletBC(bVelocity); p[0..2] = specifiedX/Y/ZVelocity (F/FlowBC.h:10-21) */
  FVMGPU_FLOWBC_SYMMETRY = 1,    /* GenericBCS<Vector,DiagTensor,T>::applySymmetryBC (F/GenericBCS.h:569-615);
                                    zero mass flux; on face groups of kind SYMMETRY the velocity and pressure
                                    gradients of the ghost cells are reflected (F/GradientModel.h:21-86) */
  FVMGPU_FLOWBC_VELOCITY = 2,    /* "VelocityBoundary": per face extrapolation where massFlux > 0, else Dirichlet
                                    with p[0..2] (F/FlowModel_impl.h:650-668); fixed mass flux rho v.A in the
                                    continuity equation (F/FlowModelVelocityBC.h:11-103) */
  FVMGPU_FLOWBC_PRESSURE = 3,    /* "PressureBoundary": the same momentum treatment + fixedPressureMomentumBC,
                                    fixedPressureContinuityBC, pressureBoundaryPostContinuitySolve
                                    (F/FlowModelPressureBC.h:11-216); p[3] = specifiedPressure. With a pressure
                                    boundary present no reference cell / net-flux redistribution is used */
  FVMGPU_FLOWBC_SLIP_JUMP = 4    /* "SlipJump": slipJumpMomentumBC (F/FlowModelSlipJump.h:11-88): Dirichlet value
                                    per face = wall-parallel cell velocity * a*lambda / (dn + a*lambda) + the
                                    specified velocity's part; p[0..2] = specified velocity, p[4] =
                                    accomodationCoefficient; fixed-flux continuity like a wall. Needs the face
                                    centroids (fvmgpu_mesh_set_geometry / _compute_geometry) */
};
typedef struct {
  double momentumURF;   /* FlowModelOptions "momentumURF" (0.7)  F/FlowBC.h:42-50 */
  double pressureURF;   /* "pressureURF" (0.3)                                  */
  int transient;        /* TimeDerivativeDiscretization on the momentum equations */
  int time_order;
  double dt;
  int correctVelocity;  /* FlowModelOptions::correctVelocity (true)             */
  /* gas state for the slip-wall mean free path (F/FlowBC.h:50-52,64; F/FlowModelSlipJump.h:38-43,67-71) */
  double operatingPressure;     /* 101325 */
  double operatingTemperature;  /* 300    */
  double molecularWeight;       /* 28.966 */
  int incompressible;           /* true: pAbs = operatingPressure, else p[c0] + operatingPressure */
} fvmgpu_flow_opts;
int fvmgpu_flow_create(fvmgpu_flow_t* out, fvmgpu_mesh_t mesh);
int fvmgpu_flow_destroy(fvmgpu_flow_t flow);
int fvmgpu_flow_set_field(fvmgpu_flow_t flow, int field, const double* host, long long n);
int fvmgpu_flow_fill_field(fvmgpu_flow_t flow, int field, double value);
int fvmgpu_flow_get_field(fvmgpu_flow_t flow, int field, double* host, long long n);
int fvmgpu_flow_set_bc(fvmgpu_flow_t flow, int groupId, int bcKind, const double* p, int np);
/* multi-GPU only: the reference pressure-correction cell is the globally lowest fluid cell
 * (F/FlowModel_impl.h:931-994); its owner passes the local index, every other rank -1 (default 0) */
int fvmgpu_flow_set_reference_cell(fvmgpu_flow_t flow, int localCell);
/* FlowModel::init: default face mass fluxes + continuity residual (F/FlowModel_impl.h:222-340) */
int fvmgpu_flow_init(fvmgpu_flow_t flow);
/* initMomentumLinearization + initAssembly + linearizeMomentum + initSolve (:522-737) */
int fvmgpu_flow_assemble_momentum(fvmgpu_flow_t flow, const fvmgpu_flow_opts* opts);
/* parity hook: VVMatrix diag (3 per cell) / scalar offdiag / b (3 per cell) */
int fvmgpu_flow_download_momentum(fvmgpu_flow_t flow, double* diag3, double* offdiag, double* b3);
/* momentumLinearSolver.solve + postSolve + updateSolution + momAp (:744-768). bcgstab != 0:
 * BCGStab (its nMaxIterations / tolerances given here) preconditioned by `solver` (1: one AMG cycle,
 * 2: the reference's ILU(0), see fvmgpu_bcgstab_ilu0_solve); 3: JacobiSolver with those limits; 4: CG
 * preconditioned by one AMG cycle.
 * rnorm0[3] = initial residual 1-norm per velocity component, iters[3] */
int fvmgpu_flow_solve_momentum(fvmgpu_flow_t flow, fvmgpu_solver_t solver, int bcgstab, int bcgMaxIterations,
                               double bcgRelTol, double bcgAbsTol, double* rnorm0, int* iters);
/* discretizeContinuity (:1394-1407): Rhie-Chow mass fluxes + pressure-correction matrix */
int fvmgpu_flow_assemble_continuity(fvmgpu_flow_t flow, const fvmgpu_flow_opts* opts);
int fvmgpu_flow_download_continuity(fvmgpu_flow_t flow, double* diag, double* offdiag, double* b, int* isBoundary);
/* pressureLinearSolver.solve + postSolve + postContinuitySolve (:1410-1430, 1263-1339) */
int fvmgpu_flow_solve_continuity(fvmgpu_flow_t flow, fvmgpu_solver_t solver, int bcgstab, int bcgMaxIterations,
                                 double bcgRelTol, double bcgAbsTol, const fvmgpu_flow_opts* opts, double* rnorm0,
                                 int* iters);

/* ---- multi-GPU (one process per GPU, one mesh part per GPU; NCCL resolved at run time with dlopen)
 * The reference's MPI layer maps as follows (all stream ordered, no host staging):
 *   MultiField::sync / Field::syncLocal  (Isend/Irecv of packed ghosts, F/MultiField.cpp:488-551,
 *       F/Field.cpp:333-394)    -> ONE kernel that gathers the interface rows, stores them into the neighbour's
 *                                  memory over NVLink (CUDA IPC peer mapping), raises a flag there, waits for the
 *                                  neighbour's flag and scatters what arrived into the ghost cells (csrc/peer.cuh)
 *   MultiFieldReduction::reduceSum (Allreduce SUM, F/MultiFieldReduction.cpp:213-225)
 *                               -> every rank stores its partial sums into all peers and adds them in rank order
 *   LinearSystemMerger (coarse levels gathered below a size threshold, F/LinearSystemMerger.cpp)
 *                               -> coarse level all-gathered (peer stores) and solved replicated
 * The peer-memory transport needs all ranks on one node with peer access (NVLink / NVSwitch); otherwise, or with
 * FVMGPU_PEER=0, the same steps run as pack kernel + grouped ncclSend/ncclRecv + unpack kernel / ncclAllReduce /
 * ncclAllGather. NCCL is also what carries the IPC handles at set-up.
 * After fvmgpu_comm_init with nranks > 1, every rank must make the same sequence of solver /
 * assembly calls on meshes that carry halo maps (fvmgpu_mesh_set_halo). */
int fvmgpu_comm_unique_id(void* out128);                               /* ncclGetUniqueId      */
int fvmgpu_comm_init(int nranks, int rank, const void* uniqueId128);   /* ncclCommInitRank     */
int fvmgpu_comm_destroy(void);
int fvmgpu_comm_counters(long long* collectives);  /* halo exchanges + reductions issued since init */
/* Field::syncLocal for one cell field of the system (F/Field.cpp:333-394): ghost cells of the
 * interface groups receive the owning rank's values */
int fvmgpu_system_halo_exchange(fvmgpu_system_t sys, int field);

#ifdef __cplusplus
}
#endif
#endif /* FVMGPU_H_ */
