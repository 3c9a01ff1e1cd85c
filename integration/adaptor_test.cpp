// adaptor_test.cpp -- exercises integration/fvm_gpu_adaptor.h against the UNMODIFIED reference
// classes (compiled in place by oracle/Makefile into oracle/_ref/adaptor_test; GPU needed).
//
//  1. builds an n x n quad mesh through the reference's raw Mesh ctor, MeshMetricsCalculator;
//  2. runs the reference ThermalModel twice on it, once with the reference AMG (CPU) and once with
//     `options.linearSolver = new GpuAMG` (the drop-in): same script, solver swapped -- and then
//     with GpuBCGStab;
//  3. re-does the reference's linearize + initSolve on the GPU with GpuScalarLinearizer and
//     compares CRMatrix diag / offdiag / b entry by entry with the reference's own assembly;
//  4. runs one whole outer iteration device-resident (GpuScalarLinearizer::iterateOnDevice: the matrix never
//     leaves the GPU) against ThermalModel::advance(1) of the reference;
//  5. runs a lid-driven cavity with GpuFlowModel (FlowModel::advance on the device, boundary conditions /
//     material / options read from the reference's own FlowModel object) against the reference FlowModel.
// Prints one line per check and exits non-zero on failure.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <sstream>

#define private public  // read ThermalModel<T>::Impl the way Impl::dumpMatrix does (test only)
#include <atype.h>
#include "AMG.h"
#include "CRMatrix.h"
#include "CRMatrixTranspose.h"
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
#undef private
#include "BCGStab.h"
#include "fvm_gpu_adaptor.h"

template class FlowModel<double>;
using namespace fvmgpu_adaptor;
typedef Vector<double, 3> Vec3;

static Mesh* quadMesh(int n) {
  const int np = n + 1;
  Array<Vec3> coords(np * np);
  for (int j = 0; j < np; j++)
    for (int i = 0; i < np; i++) {
      coords[i + np * j][0] = double(i) / n;
      coords[i + np * j][1] = double(j) / n + (0.3 / n) * std::sin(7.0 * i / n) * (j > 0 && j < n);
      coords[i + np * j][2] = 0;
    }
  std::vector<int> fc, fn;
  std::vector<int> gs;
  int cnt = 0;
  for (int j = 0; j < n; j++) for (int i = 0; i < n - 1; i++) { fc.push_back(i + n * j); fc.push_back(i + 1 + n * j); fn.push_back(i + 1 + np * j); fn.push_back(i + 1 + np * (j + 1)); cnt++; }
  for (int j = 0; j < n - 1; j++) for (int i = 0; i < n; i++) { fc.push_back(i + n * j); fc.push_back(i + n * (j + 1)); fn.push_back(i + 1 + np * (j + 1)); fn.push_back(i + np * (j + 1)); cnt++; }
  gs.push_back(cnt);
  int ghost = n * n;
  for (int j = 0; j < n; j++) { fc.push_back(n * j); fc.push_back(ghost++); fn.push_back(np * (j + 1)); fn.push_back(np * j); }
  gs.push_back(n);
  for (int j = 0; j < n; j++) { fc.push_back(n - 1 + n * j); fc.push_back(ghost++); fn.push_back(n + np * j); fn.push_back(n + np * (j + 1)); }
  gs.push_back(n);
  for (int i = 0; i < n; i++) { fc.push_back(i); fc.push_back(ghost++); fn.push_back(i); fn.push_back(i + 1); }
  gs.push_back(n);
  for (int i = 0; i < n; i++) { fc.push_back(i + n * (n - 1)); fc.push_back(ghost++); fn.push_back(i + 1 + np * n); fn.push_back(i + np * n); }
  gs.push_back(n);
  const int nFaces = (int)fc.size() / 2;
  Array<int> afc(2 * nFaces), afn(2 * nFaces), afnc(nFaces), ags((int)gs.size());
  for (int k = 0; k < 2 * nFaces; k++) { afc[k] = fc[k]; afn[k] = fn[k]; }
  for (int f = 0; f < nFaces; f++) afnc[f] = 2;
  for (size_t g = 0; g < gs.size(); g++) ags[(int)g] = gs[g];
  return new Mesh(2, n * n, coords, afc, afn, afnc, ags);
}

static void setBCs(ThermalModel<double>& tm) {
  tm.getBCMap()[4]->bcType = "SpecifiedTemperature";
  tm.getBCMap()[4]->find("specifiedTemperature")->second.constant = 400;
  tm.getBCMap()[3]->bcType = "SpecifiedTemperature";
  tm.getBCMap()[3]->find("specifiedTemperature")->second.constant = 300;
  tm.getBCMap()[1]->bcType = "SpecifiedHeatFlux";
  tm.getBCMap()[1]->find("specifiedHeatFlux")->second.constant = 12.0;
}

static double relL2(const Array<double>& a, const Array<double>& b) {
  double num = 0, den = 0;
  for (int i = 0; i < a.getLength(); i++) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
  return std::sqrt(num / den);
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 24;
  int failures = 0;
  try {
    Mesh* mesh = quadMesh(n);
    MeshList meshes;
    meshes.push_back(mesh);
    GeomFields geom("geom");
    MeshMetricsCalculator<double> mc(geom, meshes);
    mc.init();
    const StorageSite& cells = mesh->getCells();

    // ---- 2. solver drop-in
    Array<double> xCpu(cells.getCount()), xGpu(cells.getCount()), xBcg(cells.getCount());
    for (int variant = 0; variant < 3; variant++) {
      ThermalFields tf("therm");
      ThermalModel<double> tm(geom, tf, meshes);
      setBCs(tm);
      AMG cpu;
      GpuAMG gpu;
      GpuBCGStab bcg;
      GpuAMG pc;
      LinearSolver* s = variant == 0 ? (LinearSolver*)&cpu : variant == 1 ? (LinearSolver*)&gpu : (LinearSolver*)&bcg;
      bcg.preconditioner = &pc;
      pc.verbosity = 0;
      s->relativeTolerance = 1e-13;
      s->nMaxIterations = variant == 2 ? 200 : 5000;
      s->verbosity = 0;
      tm.getOptions().linearSolver = s;
      tm.init();
      std::stringstream sink;
      std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
      tm.advance(1);
      std::cout.rdbuf(old);
      const Array<double>& x = dynamic_cast<const Array<double>&>(tf.temperature[cells]);
      Array<double>& dst = variant == 0 ? xCpu : variant == 1 ? xGpu : xBcg;
      for (int i = 0; i < x.getLength(); i++) dst[i] = x[i];
      int its = variant == 0 ? cpu.getTotalIterations() : variant == 1 ? gpu.getTotalIterations() : bcg.getTotalIterations();
      printf("solve variant %d (%s): %d iterations\n", variant, variant == 0 ? "reference AMG, CPU" : variant == 1 ? "GpuAMG" : "GpuBCGStab+GpuAMG", its);
    }
    const double e1 = relL2(xGpu, xCpu), e2 = relL2(xBcg, xCpu);
    printf("ThermalModel.advance with GpuAMG vs reference AMG: rel L2 = %.3e %s\n", e1, e1 <= 1e-8 ? "OK" : "FAIL");
    printf("ThermalModel.advance with GpuBCGStab vs reference AMG: rel L2 = %.3e %s\n", e2, e2 <= 1e-8 ? "OK" : "FAIL");
    failures += !(e1 <= 1e-8) + !(e2 <= 1e-8);

    // ---- 3. assembly drop-in
    {
      ThermalFields tf("therm");
      ThermalModel<double> tm(geom, tf, meshes);
      setBCs(tm);
      tm.init();
      Array<double>& T = dynamic_cast<Array<double>&>(tf.temperature[cells]);
      Array<double>& k = dynamic_cast<Array<double>&>(tf.conductivity[cells]);
      for (int i = 0; i < T.getLength(); i++) { T[i] = 300 + 10 * std::sin(0.37 * i); k[i] = 1.0 + 0.5 * std::cos(0.11 * i); }
      Array<double> T0(T.getLength());
      for (int i = 0; i < T.getLength(); i++) T0[i] = T[i];
      ThermalModel<double>::Impl& impl = *tm._impl;
      LinearSystem lsRef;
      impl.initLinearization(lsRef);
      lsRef.initAssembly();
      impl.linearize(lsRef);
      lsRef.initSolve();
      for (int i = 0; i < T.getLength(); i++) T[i] = T0[i];
      LinearSystem lsGpu;
      impl.initLinearization(lsGpu);
      lsGpu.initAssembly();
      GpuMesh gm(*mesh, geom);
      GpuScalarLinearizer lin(gm);
      std::vector<GpuBC> bcs;
      GpuBC b4 = {4, FVMGPU_BC_DIRICHLET, {400, 0, 0, 0}}, b3 = {3, FVMGPU_BC_DIRICHLET, {300, 0, 0, 0}},
            b1 = {1, FVMGPU_BC_NEUMANN, {12.0, 0, 0, 0}}, b2 = {2, FVMGPU_BC_NEUMANN, {0, 0, 0, 0}};
      bcs.push_back(b4); bcs.push_back(b3); bcs.push_back(b1); bcs.push_back(b2);
      fvmgpu_assemble_opts o = {1, 0, 1, 0, 0.0, 0.0, 1, 1};
      lin.linearize(lsGpu, tf.temperature, tf.conductivity, tf.source, bcs, o);
      MultiField::ArrayIndex ti(&tf.temperature, &cells);
      ScalarMatrix& mr = dynamic_cast<ScalarMatrix&>(lsRef.getMatrix().getMatrix(ti, ti));
      ScalarMatrix& mg = dynamic_cast<ScalarMatrix&>(lsGpu.getMatrix().getMatrix(ti, ti));
      const Array<double>& br = dynamic_cast<const Array<double>&>(lsRef.getB()[ti]);
      const Array<double>& bg = dynamic_cast<const Array<double>&>(lsGpu.getB()[ti]);
      double worst = 0, scale = 0;
      for (int i = 0; i < mr.getDiag().getLength(); i++) { worst = std::max(worst, std::fabs(mr.getDiag()[i] - mg.getDiag()[i])); scale = std::max(scale, std::fabs(mr.getDiag()[i])); }
      for (int i = 0; i < mr.getOffDiag().getLength(); i++) worst = std::max(worst, std::fabs(mr.getOffDiag()[i] - mg.getOffDiag()[i]));
      double worstB = 0, scaleB = 0;
      for (int i = 0; i < br.getLength(); i++) { worstB = std::max(worstB, std::fabs(br[i] - bg[i])); scaleB = std::max(scaleB, std::fabs(br[i])); }
      const bool ok = worst <= 1e-12 * scale && worstB <= 1e-12 * scaleB;
      printf("GpuScalarLinearizer vs reference linearize+initSolve: max |dA| = %.3e (scale %.3e), max |db| = %.3e (scale %.3e) %s\n",
             worst, scale, worstB, scaleB, ok ? "OK" : "FAIL");
      failures += !ok;
    }
    // ---- 4. one outer iteration with the system resident on the device
    {
      ThermalFields tfRef("therm"), tfGpu("therm");
      ThermalModel<double> tmRef(geom, tfRef, meshes), tmGpu(geom, tfGpu, meshes);
      setBCs(tmRef);
      setBCs(tmGpu);
      AMG cpu;
      cpu.relativeTolerance = 1e-13; cpu.nMaxIterations = 5000; cpu.verbosity = 0;
      tmRef.getOptions().linearSolver = &cpu;
      tmRef.init();
      tmGpu.init();
      std::stringstream sink;
      std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
      tmRef.advance(1);
      std::cout.rdbuf(old);
      GpuMesh gm(*mesh, geom);
      GpuScalarLinearizer lin(gm);
      GpuAMG gpu;
      gpu.relativeTolerance = 1e-13; gpu.nMaxIterations = 5000; gpu.verbosity = 0;
      std::vector<GpuBC> bcs;
      GpuBC b4 = {4, FVMGPU_BC_DIRICHLET, {400, 0, 0, 0}}, b3 = {3, FVMGPU_BC_DIRICHLET, {300, 0, 0, 0}},
            b1 = {1, FVMGPU_BC_NEUMANN, {12.0, 0, 0, 0}}, b2 = {2, FVMGPU_BC_NEUMANN, {0, 0, 0, 0}};
      bcs.push_back(b4); bcs.push_back(b3); bcs.push_back(b1); bcs.push_back(b2);
      fvmgpu_assemble_opts o = {1, 0, 1, 0, 0.0, 0.0, 1, 1};
      int its = 0;
      const double r0 = lin.iterateOnDevice(gpu, tfGpu.temperature, tfGpu.conductivity, tfGpu.source, bcs, o, 0, &its);
      const double e = relL2(dynamic_cast<const Array<double>&>(tfGpu.temperature[cells]),
                             dynamic_cast<const Array<double>&>(tfRef.temperature[cells]));
      printf("device-resident outer iteration (linearize -> GpuAMG -> postSolve/update, %d cycles, r0 = %.6g) vs "
             "ThermalModel.advance(1): rel L2 = %.3e %s\n", its, r0, e, e <= 1e-8 ? "OK" : "FAIL");
      failures += !(e <= 1e-8);
    }

    // ---- 5. FlowModel::advance on the device, driven by the reference's own FlowModel object
    {
      // single-rank identity local <-> global cell maps: the -DFVM_PARALLEL build of FlowModel reads them in
      // setDirichlet (F/FlowModel_impl.h:931-968); MeshPartitioner fills them even for one part, the raw Mesh ctor not
      if (!mesh->getLocalToGlobalPtr()) {
        mesh->createLocalGlobalArray();
        Array<int>& l2g = mesh->getLocalToGlobal();
        for (int i = 0; i < cells.getCount(); i++) { l2g[i] = i; mesh->getGlobalToLocal()[i] = i; }
      }
      FlowFields ffRef("flow"), ffGpu("flow");
      FlowModel<double> fmRef(geom, ffRef, meshes), fmGpu(geom, ffGpu, meshes);
      FlowModel<double>* both[2] = {&fmRef, &fmGpu};
      AMG ms, ps;
      for (int k = 0; k < 2; k++) {
        FlowModel<double>& fm = *both[k];
        fm.getBCMap()[4]->find("specifiedXVelocity")->second.constant = 1.0;   // the lid
        fm.getVCMap()[mesh->getID()]->find("viscosity")->second.constant = 0.1;
        fm.getOptions().momentumTolerance = 1e-30;
        fm.getOptions().continuityTolerance = 1e-30;
      }
      ms.relativeTolerance = ps.relativeTolerance = 1e-13;
      ms.nMaxIterations = ps.nMaxIterations = 3000;
      ms.verbosity = ps.verbosity = 0;
      fmRef.getOptions().momentumLinearSolver = &ms;
      fmRef.getOptions().pressureLinearSolver = &ps;
      std::stringstream sink;
      std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
      fmRef.init();
      fmGpu.init();
      fmRef.advance(3);
      GpuMesh gm(*mesh, geom);
      GpuFlowModel g(gm, fmGpu, ffGpu);
      g.momentumSolver.relativeTolerance = g.pressureSolver.relativeTolerance = 1e-13;
      g.momentumSolver.nMaxIterations = g.pressureSolver.nMaxIterations = 3000;
      g.init();
      g.advance(3);
      std::cout.rdbuf(old);
      const Array<Vec3>& vr = dynamic_cast<const Array<Vec3>&>(ffRef.velocity[cells]);
      const Array<Vec3>& vg = dynamic_cast<const Array<Vec3>&>(ffGpu.velocity[cells]);
      double num = 0, den = 0;
      for (int i = 0; i < cells.getSelfCount(); i++)
        for (int k = 0; k < 3; k++) { num += (vg[i][k] - vr[i][k]) * (vg[i][k] - vr[i][k]); den += vr[i][k] * vr[i][k]; }
      const double ev = std::sqrt(num / den);
      const Array<double>& pr = dynamic_cast<const Array<double>&>(ffRef.pressure[cells]);
      const Array<double>& pg = dynamic_cast<const Array<double>&>(ffGpu.pressure[cells]);
      num = den = 0;
      for (int i = 0; i < cells.getSelfCount(); i++) { num += (pg[i] - pr[i]) * (pg[i] - pr[i]); den += pr[i] * pr[i]; }
      const double ep = std::sqrt(num / den);
      const bool ok = ev <= 1e-8 && ep <= 1e-8;
      printf("GpuFlowModel.advance(3) vs reference FlowModel.advance(3) on the lid-driven cavity: velocity rel L2 = %.3e, "
             "pressure rel L2 = %.3e %s\n", ev, ep, ok ? "OK" : "FAIL");
      failures += !ok;
    }
  } catch (std::exception& e) {
    printf("EXCEPTION: %s\n", e.what());
    return 2;
  }
  printf(failures ? "adaptor_test: FAILED\n" : "adaptor_test: OK\n");
  return failures ? 1 : 0;
}
