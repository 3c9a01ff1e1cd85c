// fvm_b200 / libfvmgpu -- AMG hierarchy and Krylov driver declarations (see solver.cu).
#pragma once
#include "structs.cuh"

#include <functional>

namespace fvmgpu {

// Column indices of a SELL-32 level as the hot row kernels read them. On top of the plain 32-bit array the level
// keeps a COMPRESSED copy: one base per (slice, entry position) and a 16-bit offset per element -- in a slice of 32
// neighbouring rows the k-th columns lie close together (rows store their entries in ascending column order; the
// colour ordering keeps mesh neighbours of neighbouring rows next to each other), so 2 bytes instead of 4 per stored
// entry cross the memory bus (12 -> 10 B per entry, ~11 % of a row pass on hexes). A slice whose k-th columns
// spread over more than 65535 rows (unstructured numbering, interface rows with halo columns) is marked in `mode`
// and read from the plain array.
struct SellCols {
  const int* scol; const unsigned short* c16; const int* base; const unsigned char* mode;   // mode == nullptr: plain only
  FVM_DEV bool compressed(int slice) const { return mode != nullptr && mode[slice] != 0; }
  FVM_DEV int at(int p, bool cmp) const { return cmp ? base[p >> 5] + (int)c16[p] : scol[p]; }
};

struct Level {
  int n = 0;                 // solved rows of this level
  long long nnzStored = 0;   // SELL elements incl. padding
  long long nnzTrue = 0;     // off-diagonal entries (-1: still on the device in nnzDev)
  DBuf<double> nnzDev;
  int nSlices = 0;
  DBuf<int> sliceOff;        // nSlices+1 element offsets into scol/sval (multiples of 32)
  DBuf<int> scol;
  DBuf<unsigned short> scol16;   // compressed copy of scol (see SellCols); empty when not built
  DBuf<int> colBase;             // nnzStored / 32 bases
  DBuf<unsigned char> sliceMode; // per slice: 1 = read the compressed copy
  DBuf<double> modeCount;        // number of such slices (device; fetched when somebody asks)
  long long compressedSlices = -1;
  SellCols cols() const { return SellCols{scol.p, scol16.p, colBase.p, scol16.p ? sliceMode.p : nullptr}; }
  DBuf<double> sval;
  DBuf<double> diag, b, x, r;
  DBuf<double> mb, mx, mr;   // the same three vectors with NC values per row (Amg::solveMulti)
  DBuf<int> nat;             // level row -> index in the level's natural (pre-colouring) numbering
  // colouring: rows [colourStart[c], colourStart[c+1]) have colour c
  int nColours = 0;
  std::vector<int> colourStart;
  std::vector<int> ifaceCount;  // multi-GPU: the first ifaceCount[c] rows of colour c have a halo column
  // link to the next coarser level (in ITS numbering)
  DBuf<int> ci;              // n: coarse row of each fine row, -1 = not coarsened
  DBuf<int> memOff, mem;     // aggregate (natural id) -> its fine rows (ascending)
  DBuf<int> cpos;            // aggregate (natural id) -> coarse row
  bool xZero = false;        // x is known to be identically zero
  bool rValid = false;       // r holds b + A x for the current x
  int rZeroFrom = 0, rZeroTo = 0;  // ... except on this row range, where r is an exact zero that is not stored
  // multi-GPU: x and r carry nGhost extra slots (columns >= n of the matrix) filled by the halo
  // exchange with the ranks that own those rows (the reference: MultiField::sync after every
  // sweep, F/MultiFieldMatrix.cpp:164,216,397)
  int nGhost = 0;
  Halo halo;
  double globalRows = 0;         // rows of this level summed over the ranks; anyTiny: some rank has <= 3
  bool anyTiny = false;
  DBuf<int> ghostCoarse;         // per ghost slot: x index in the NEXT level (>= its n), -1 = none
};

// pattern-only part of level 0 (see buildLevelFromCsr)
struct PatternCache {
  unsigned long long stamp = 0;
  int n = 0; bool dropGhost = false, splitIface = false;
  int nColours = 0, nSlices = 0;
  long long nnzStored = 0, nnzTrue = 0;
  std::vector<int> colourStart, ifaceCount;
  DBuf<int> perm, nat, sliceOff;
  DBuf<double> nnzDev;
};

struct Ilu0;
struct Ilu0Deleter { void operator()(Ilu0* p) const; };

struct Amg {
  fvmgpu_amg_opts opts;
  // ILU(0) factors of the last system (csrc/ilu.cu): ILU0Solver and BCGStab's ILU0 preconditioner
  std::unique_ptr<Ilu0, Ilu0Deleter> ilu;
  Ilu0& iluFor(System* sys);
  void iluSmooth(System* sys, const double* rhs, double* out);
  int iluLevels(System* sys);
  void iluSolve(System* sys, int nMaxIterations, double relTol, double absTol, double* rnorm0, double* rnorm, int* iters);
  int precondKind = 0;   // what bcgstab() applies: 0 one AMG cycle, 1 ILU(0)
  DBuf<double> natIn, natOut;
  std::vector<std::unique_ptr<Level>> levels;
  DBuf<int> perm0;           // system (natural) row -> level-0 row
  int cycleBudget = 1 << 30;   // upper bound of the cycles the coming solve may run (set by the solve entry points)
  int lastSolveCycles = -1;    // cycles the previous solve took; survives cleanup() like cache0
  PatternCache cache0;       // survives cleanup(): the next outer iteration assembles on the same pattern
  DBuf<double> scalars;      // device scalars for dots / norms
  System* builtFor = nullptr;          // the system of the last setup (used by the ILU preconditioner path)
  unsigned long long builtVersion = 0; // its System::version stamp: THE identity of the hierarchy (stamps are unique)
  // structural options the hierarchy was built with (a change rebuilds it) and the cycle options the captured
  // graphs bake in (a change re-captures them)
  int builtMaxCoarseLevels = -1, builtGroupSize = -1;
  double builtThreshold = -1;
  int graphOpts[5] = {-1, -1, -1, -1, -1};  // nPre, nPost, cycleType, smootherType, full-residual decision
  std::vector<double> history;
  long long totalIterations = 0;
  double lastSetupMs = 0, lastCyclesMs = 0;  // host wall clock of the last solve(): hierarchy build / cycle loop
  // coarse tail fused into one CTA (levels [tailStart, end) have <= kTailRows rows)
  static constexpr int kTailRows = 4096;
  int tailStart = -1, tailCount = 0;
  int tailGridLevels = 0;    // leading levels of the stretch worked by the whole grid; the rest by CTA 0
  DBuf<unsigned> coopBarrier;
  bool tailIsCoop = false;   // the stretch runs in the cooperative grid kernel (levels up to 139 K - 331 K rows, see Amg::buildTail)
  DBuf<char> tailLevels;
  std::vector<DBuf<int>> tailColourStarts;
  // captured (cycle [+ residual norm]) graphs
  bool useGraphs = true;
  // optional: a permutation of [0, n) giving each system row a "natural" index for the pairing
  // preference (merged coarse systems arrive in the colour-sorted order of the ranks' levels)
  DBuf<int> natHint;
  void* graphExec[2] = {nullptr, nullptr};
  long long graphLaunches[2] = {0, 0};
  // Krylov work vectors (level-0 numbering), kept between solves, and the captured graph of one BiCGStab iteration
  struct KrylovVectors {
    int n = -1; size_t ng = 0;
    DBuf<double> x, bOrig, r, rTilda, p, v, t, hat;
  } krylov;
  void* iterGraph = nullptr;
  long long iterGraphLaunches = 0;
  double iterGraphAbsTol = 0;
  int iterGraphKey[5] = {-1, -1, -1, -1, -1};   // nPre, nPost, cycleType, smootherType, precondKind
  // several right-hand sides on one matrix (the momentum system): see the end of solver.cu
  int multiNc = 0;
  DBuf<char> tailLevelsM;
  void* graphExecM = nullptr;
  long long graphLaunchesM = 0;
  int graphKeyM[6] = {-1, -1, -1, -1, -1, -1};
  bool graphWarmM = false;
  bool multiRhsSupported() const;
  void solveMulti(System* sys, int nc, const double* b3, double* delta3, int stride, int maxCycles, double relTol, double absTol,
                  double* rnorm0, double* rnorm, int* iters);
  void runMultiGraph(const std::function<void()>& body, int nc);
  void dropMultiGraph();
  void cycleOn(double* rhs);
  void runIterationGraph(const std::function<void()>& body, double absTol);
  void dropIterationGraph();

  // multi-GPU: below `mergeRows` global rows the level is all-gathered and the rest of the cycle runs
  // replicated on every rank with the single-GPU code (the reference's LinearSystemMerger idea,
  // F/LinearSystemMerger.cpp: gather coarse levels instead of exchanging halos of tiny levels)
  bool multi = false;
  int tagBase = 0;                 // profiler level tags of a nested hierarchy continue after the merged level
  // Overlap (NVLink peer transport; FVMGPU_OVERLAP=0 switches it off): the halo exchange of a pass is started
  // after its interface rows and finished before the interface rows of the next pass, with the interior rows
  // in between (rows are ordered interface-first inside every colour for this); levels below overlapMinRows
  // rows exchange in one piece (a pass there is shorter than the two extra launches).
  bool overlapExchange = false;   // measured on 2 B200s at 256^3 per GPU: 1.92 ms per cycle without, 1.99 ms with (FVMGPU_OVERLAP=1 switches it on)
  int overlapMinRows = 1000000;
  Halo* pendingHalo = nullptr;     // exchange begun, not yet finished
  double* pendingX = nullptr;
  bool exchangePerColour = false;  // true: halo exchange after every colour pass; false: after every half-sweep
  int mergedLevel = -1;            // index of the distributed level that is solved replicated
  int mergeMaxLocal = 0;           // rows per rank block in the merged numbering (padded)
  std::unique_ptr<System> mergedSys;
  std::unique_ptr<Amg> nested;
  DBuf<double> mergeSend, mergeB, mergeX;
  PeerPlan mergePlan;              // all-gather of the merged level's right-hand side over NVLink peer stores
  bool nestedLoaded = false;

  void setup(System* sys);   // AMG::createCoarseLevels
  void ensureSetup(System* sys);
  void cleanup();
  void solve(System* sys, double* rnorm0, double* rnorm, int* iters);
  void smooth(System* sys);
  void bcgstab(System* sys, int nMaxIterations, double relTol, double absTol, double* rnorm0, double* rnorm,
               int* iters);
  void cg(System* sys, int nMaxIterations, double relTol, double absTol, double* rnorm0, double* rnorm, int* iters);
  void bcgstabMulti(System* sys, int nc, const double* b3, double* delta3, int nMaxIterations, double relTol,
                    double absTol, double* rnorm0, double* rnorm, int* iters);
  void cgMulti(System* sys, int nc, const double* b3, double* delta3, int nMaxIterations, double relTol, double absTol,
               double* rnorm0, double* rnorm, int* iters);
  void jacobiSolve(System* sys, int nMaxIterations, double relTol, double absTol, double* rnorm0, double* rnorm,
                   int* iters);

  void sweeps(int nSweeps, int lvl, bool ghostsReadAfter);
  void residual(int lvl);
  double residualNorm(int lvl);
  void cycle(int cycleType, int lvl);
  void loadSystem(System* sys, const double* b_d, const double* x_d);
  void storeDelta(double* delta_d);
  void precondition(const double* rhsPerm, double* outPerm);
  void buildMerged();
  void cycleMerged(int cycleType, int lvl);
  void exchange(Level& L, double* x);
  bool overlapOn(const Level& L) const;
  void forkExchange(Level& L, double* x);
  void joinExchange();
  void buildTail();
  void runTail();
  void dropGraphs();
  void cycleGraphed(int kind);
  ~Amg() { dropGraphs(); }
};

void setDebugAggregator(fvmgpu_aggregate_fn fn, void* user);  // solver.cu
int tailTraceRead(int cap, unsigned long long* times, int* tags);

// mesh.cu / assemble.cu entry points used by capi.cu
Mesh* meshCreate(int dim, int nSelf, int nTotal, int nFaces, const int* faceCells, const int* ccRow,
                 const int* ccCol, int nGroups, const int* gOff, const int* gCnt, const int* gId, const int* gKind);
void meshSetGeometry(Mesh* m, const double* faceArea, const double* faceAreaMag, const double* faceCentroid,
                     const double* cellCentroid, const double* cellVolume, const int* ibType);
void meshComputeGeometry(Mesh* m, int nNodes, const double* nodes, const int* faceNodeOffsets, const int* faceNodes,
                         double* faceArea, double* faceAreaMag, double* faceCentroid, double* cellCentroid,
                         double* cellVolume);
void meshSetHalo(Mesh* m, int nNeigh, const int* peerRank, const int* scatterOff, const int* scatterIdx,
                 const int* gatherOff, const int* gatherIdx);
System* systemCreate(Mesh* m);
System* systemCreateRaw(int nSelf, int nGhost, const int* row, const int* col, const double* diag,
                        const double* off, const double* b);
void systemSetField(System* s, int field, const double* host, long long n, bool fill, double value);
void systemGetField(System* s, int field, double* host, long long n);
void systemSetBc(System* s, int groupId, int kind, const double* p, int np, const double* perFace);
void systemHaloExchange(System* s, int field);
void computeGradient(System* s);
void assemble(System* s, const fvmgpu_assemble_opts& o);
void postSolveUpdate(System* s);

// electric.cu
void electricField(System* s, double* E_host);
void electricDriftFlux(System* potential, System* charge, double mobility, double vsat, int nSym, const int* symGroupIds,
                       double* vel_host);

// flow.cu
struct Flow;
Flow* flowCreate(Mesh* m);
void flowDestroy(Flow* F);
void flowSetField(Flow* F, int field, const double* host, long long n, bool fill, double value);
void flowGetField(Flow* F, int field, double* host, long long n);
void flowSetBc(Flow* F, int groupId, int kind, const double* p, int np);
void flowInit(Flow* F);
void flowSetReferenceCell(Flow* F, int localCell);
void flowAssembleMomentum(Flow* F, const fvmgpu_flow_opts& o);
void flowDownloadMomentum(Flow* F, double* diag3, double* off, double* b3);
void flowSolveMomentum(Flow* F, Amg* solver, int useBcgstab, int bcgMaxIter, double bcgRel, double bcgAbs,
                       double* rnorm0, int* iters);
void flowAssembleContinuity(Flow* F, const fvmgpu_flow_opts& o);
void flowDownloadContinuity(Flow* F, double* diag, double* off, double* b, int* isBoundary);
void flowSolveContinuity(Flow* F, Amg* solver, int useBcgstab, int bcgMaxIter, double bcgRel, double bcgAbs,
                         const fvmgpu_flow_opts& o, double* rnorm0, int* iters);

}  // namespace fvmgpu
