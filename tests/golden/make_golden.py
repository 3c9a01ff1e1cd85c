"""Generate the committed golden fixtures from the REFERENCE (run in the dev container only).

Needs /root/reference (its test meshes + golden files) and oracle/_ref/libfvmref.so (the reference
hot path compiled in place by oracle/Makefile). Writes small .npz files next to this script:

  mm226.npz        T/MatrixMarket226.dat + T/rhs226.dat as CSR-with-separate-diagonal, the
                   reference AMG's level sizes / residuals (T/testLinearSolver.out:5-11) and its
                   solution at rel-tol 1e-13
  cav32.npz        raw mesh arrays of T/cav32.cas as the reference's FluentReader numbers them,
                   reference geometry, the thermal system of T/THERMAL_MATRIX (matrix/rhs goldens
                   parsed from T/THERMAL_MATRIX/GOLDEN), the AMG history of
                   T/AMG_MERGING_THERMAL/proc1/GOLDEN/convergence.dat and the converged temperature
  hex_bcs.npz      a jittered 6x5x4 hex mesh with every thermal BC kind, random conductivity and
                   source: reference gradient / matrix / rhs before and after boundary elimination
  tet_solve.npz    a 5x5x5x6 tet mesh: reference assembled system + converged solution + heat fluxes
  flow_cavity.npz  lid-driven cavity on a jittered 12x10 quad mesh (FlowModel, NoSlipWall, lid u = 1,
                   mu = 0.1, rho = 1.3): the reference state after 6 SIMPLE iterations, its momentum
                   system assembled from that state, the post-momentum state, the pressure-correction
                   system, and the converged steady flow field (tolerances 1e-9)
  flow_symmetry.npz FlowModel in a jittered 6x5x4 hex box with two "symmetry" face groups and a moving lid:
                   momentum system from a developed state and the fields after 6 SIMPLE iterations with
                   tight inner solves
  flow_channel.npz FlowModel channel flow on a jittered 14x8 quad mesh: VelocityBoundary inlet, PressureBoundary
                   outlet, walls: momentum + pressure-correction systems from a developed state and the
                   fields after 8 SIMPLE iterations with tight inner solves
  electric_box.npz ElectricModel on a jittered 6x5x7 hex box (1 x 1 x 2 um): Poisson equation with
                   SpecifiedPotential / SpecifiedPotentialFlux / Symmetry / SpecialDielectricBoundary
                   BCs and a uniform total charge, then drift + transient charge transport of the
                   conduction-band component (nTrap = 2): reference fields after two time steps
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refapi as R  # noqa: E402
from fvm_b200 import meshgen as G  # noqa: E402

T = "/root/reference/src/fvm/test/"


def csr_from_mm(path_m, path_rhs):
    mm = np.loadtxt(path_m, skiprows=2)
    rhs = np.loadtxt(path_rhs)
    n = len(rhs)
    i = mm[:, 0].astype(int) - 1
    j = mm[:, 1].astype(int) - 1
    v = mm[:, 2]
    diag = np.zeros(n)
    diag[i[i == j]] = v[i == j]
    m = i != j
    order = np.argsort(i[m], kind="stable")
    ri, cj, vv = i[m][order], j[m][order], v[m][order]
    row = np.zeros(n + 1, np.int32)
    np.cumsum(np.bincount(ri, minlength=n), out=row[1:])
    return n, row, cj.astype(np.int32), diag, vv, -rhs  # MMReader: b = -rhs (I/MMReader.cpp:176)


def mesh_arrays(rm):
    c, g = rm.connectivity(), rm.geometry()
    return dict(dim=rm.dim, n_self=rm.n_self, n_total=rm.n_total, face_cells=c["face_cells"],
                cc_row=c["cc_row"], cc_col=c["cc_col"], pair_to_col=c["pair_to_col"],
                group_offset=c["group_offset"], group_count=c["group_count"], group_id=c["group_id"],
                group_kind=c["group_kind"], face_area=g["face_area"], face_area_mag=g["face_area_mag"],
                face_centroid=g["face_centroid"], cell_centroid=g["cell_centroid"],
                cell_volume=g["cell_volume"], ib_type=g["ib_type"])


def main():
    # ---- mm226
    n, row, col, diag, off, b = csr_from_mm(T + "MatrixMarket226.dat", T + "rhs226.dat")
    ref = R.linsolve(n, row, col, diag, off, b, R.solver_cfg(verbosity=2))
    hist = [float(l.split(":")[2].strip(" ]\n")) for l in ref["text"].splitlines() if "[test" in l]
    tight = R.linsolve(n, row, col, diag, off, b, R.solver_cfg(relativeTolerance=1e-13, nMaxIterations=500, verbosity=0))
    bcg = R.linsolve(n, row, col, diag, off, b, R.solver_cfg(kind=1, relativeTolerance=1e-13, nMaxIterations=100, verbosity=0))
    np.savez_compressed(os.path.join(HERE, "mm226.npz"), n=n, row=row, col=col, diag=diag, off=off, b=b,
                        ref_levels=np.array(ref["levels"]), ref_iters=ref["iters"], ref_rnorm0=ref["rnorm0"],
                        ref_history=np.array(hist), ref_x_tol8=ref["x"], ref_x=tight["x"],
                        ref_x_bcgstab=bcg["x"],
                        golden_text=open(T + "testLinearSolver.out").read())
    # ---- cav32 thermal
    rm = R.RefMesh.from_cas(T + "cav32.cas")
    t = R.RefThermal(rm)
    t.set_bc(3, "SpecifiedTemperature", specifiedTemperature=400)
    for g in (4, 5, 6):
        t.set_bc(g, "SpecifiedTemperature", specifiedTemperature=0)
    t.set_vc("thermalConductivity", 1.0)
    t.set_solver(R.solver_cfg(relativeTolerance=1e-9, nMaxIterations=2000, maxCoarseLevels=20, verbosity=2))
    t.init()
    a = t.assemble(1)
    text, _ = t.advance(1)
    x9 = t.field("temperature").copy()
    # tight solve for the solution-parity test
    t2 = R.RefThermal(rm)
    t2.set_bc(3, "SpecifiedTemperature", specifiedTemperature=400)
    for g in (4, 5, 6):
        t2.set_bc(g, "SpecifiedTemperature", specifiedTemperature=0)
    t2.set_solver(R.solver_cfg(relativeTolerance=1e-13, nMaxIterations=2000, verbosity=0))
    t2.init()
    t2.advance(1)
    G_ = T + "THERMAL_MATRIX/GOLDEN/"
    np.savez_compressed(os.path.join(HERE, "cav32.npz"), **mesh_arrays(rm), diag=a["diag"], off=a["offdiag"],
                        b=a["b"], x_after_bc=a["x"], ref_x_tol9=x9, ref_x=t2.field("temperature").copy(),
                        ref_text=text,
                        golden_rhs=np.loadtxt(G_ + "matrix.rhs"),
                        golden_mat=np.loadtxt(G_ + "matrix_mesh0.mat", skiprows=2),
                        golden_convergence=open(T + "AMG_MERGING_THERMAL/proc1/GOLDEN/convergence.dat").read())
    # ---- hex with every BC kind
    raw = G.hex_mesh(6, 5, 4, jitter=0.25, seed=11)
    rm = R.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes,
                            raw.face_node_count, raw.face_group_size)
    rng = np.random.default_rng(5)
    k = np.exp(rng.normal(size=rm.n_total))
    src = rng.normal(size=rm.n_total) * 100
    x0 = 300.0 + rng.uniform(-5, 5, size=rm.n_total)
    t = R.RefThermal(rm)
    t.set_bc(1, "SpecifiedTemperature", specifiedTemperature=400)
    t.set_bc(2, "SpecifiedHeatFlux", specifiedHeatFlux=25.0)
    t.set_bc(3, "Convective", convectiveCoefficient=3.0, farFieldTemperature=280.0)
    t.set_bc(4, "Radiative", surfaceEmissivity=0.8, farFieldTemperature=250.0)
    t.set_bc(5, "Mixed", convectiveCoefficient=2.0, surfaceEmissivity=0.5, farFieldTemperature=310.0)
    t.set_bc(6, "SpecifiedHeatFlux", specifiedHeatFlux=0.0)
    t.set_solver(R.solver_cfg(verbosity=0))
    t.init()
    t.field("conductivity")[:] = k
    t.field("source")[:] = src
    out = {}
    for stage in (0, 1):
        t.field("temperature")[:] = x0
        a = t.assemble(stage)
        for key in ("diag", "offdiag", "b", "x", "is_boundary"):
            out["s%d_%s" % (stage, key)] = a[key]
        out["s%d_gradient" % stage] = t.field("temperatureGradient").reshape(-1, 3).copy()
    np.savez_compressed(os.path.join(HERE, "hex_bcs.npz"), **mesh_arrays(rm), k=k, src=src, x0=x0, **out)
    # ---- tet solve
    raw = G.tet_mesh(5, 5, 5)
    rm = R.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes,
                            raw.face_node_count, raw.face_group_size)
    t = R.RefThermal(rm)
    t.set_bc(5, "SpecifiedTemperature", specifiedTemperature=300)
    t.set_bc(6, "SpecifiedTemperature", specifiedTemperature=400)
    t.set_bc(1, "SpecifiedHeatFlux", specifiedHeatFlux=7.0)
    t.set_solver(R.solver_cfg(relativeTolerance=1e-13, nMaxIterations=2000, verbosity=0))
    t.init()
    k = np.exp(0.5 * np.random.default_rng(9).normal(size=rm.n_total))
    t.field("conductivity")[:] = k
    a = t.assemble(1)
    t.advance(1)
    hf = {("hf%d" % g): t.heat_flux(int(g), int(c)) for g, c in zip(rm.connectivity()["group_id"][1:], rm.connectivity()["group_count"][1:])}
    np.savez_compressed(os.path.join(HERE, "tet_solve.npz"), **mesh_arrays(rm), k=k, diag=a["diag"], off=a["offdiag"],
                        b=a["b"], ref_x=t.field("temperature").copy(), **hf)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


def flow_golden():
    raw = G.quad_mesh(12, 10, jitter=0.2, seed=5)
    rm = R.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes, raw.face_node_count,
                            raw.face_group_size)

    def make():
        f = R.RefFlow(rm)
        for g in (1, 2, 3):
            f.set_bc(g, "NoSlipWall")
        f.set_bc(4, "NoSlipWall", specifiedXVelocity=1.0)
        f.set_vc("viscosity", 0.1)
        f.set_vc("density", 1.3)
        return f

    f = make()
    f.init()
    f.advance(6)
    out = {}
    for nm in ("velocity", "pressure", "facePressure", "massFlux", "continuityResidual"):
        out["s0_" + nm] = f.field(nm).copy()
    ms = f.momentum_system()
    out.update(mom_diag=ms["diag"], mom_off=ms["offdiag"], mom_b=ms["b"],
               mom_vgrad=f.field("velocityGradient").copy(), mom_pgrad=f.field("pressureGradient").copy())
    out["mom_rnorm"] = f.solve_momentum()
    for nm in ("velocity", "previousVelocity", "momAp", "pressureGradient", "massFlux", "pressure"):
        out["s1_" + nm] = f.field(nm).copy()
    cs = f.continuity_system()
    out.update(pp_diag=cs["diag"], pp_off=cs["offdiag"], pp_b=cs["b"], pp_is_boundary=cs["is_boundary"],
               pp_massFlux=f.field("massFlux").copy())
    f.close()
    f = make()
    f.set_option("momentumTolerance", 1e-9)
    f.set_option("continuityTolerance", 1e-9)
    f.init()
    conv, txt, _ = f.advance(600)
    assert conv
    out.update(conv_iters=len(txt.splitlines()), conv_velocity=f.field("velocity").copy(),
               conv_pressure=f.field("pressure").copy(), conv_massFlux=f.field("massFlux").copy())
    np.savez_compressed(os.path.join(HERE, "flow_cavity.npz"), **mesh_arrays(rm), **out)
    print("flow_cavity.npz: %d cells, converged in %d SIMPLE iterations" % (rm.n_self, out["conv_iters"]))


def flow_symmetry_golden():
    """3-D box with two symmetry planes (face groups 1 and 3 typed "symmetry", bcType Symmetry) and a
    moving lid: exercises the vector applySymmetryBC and the gradient / centroid reflections."""
    raw = G.hex_mesh(6, 5, 4, jitter=0.2, seed=9)
    rm = R.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes, raw.face_node_count,
                            raw.face_group_size, symmetry_groups=(1, 3))
    f = R.RefFlow(rm)
    for g in (2, 4, 5):
        f.set_bc(g, "NoSlipWall")
    f.set_bc(1, "Symmetry")
    f.set_bc(3, "Symmetry")
    f.set_bc(6, "NoSlipWall", specifiedXVelocity=1.0, specifiedYVelocity=0.3)
    f.set_vc("viscosity", 0.05)
    f.set_vc("density", 1.1)
    tight = dict(relativeTolerance=1e-14, nMaxIterations=3000, verbosity=0)
    f.set_solver(0, R.solver_cfg(**tight))
    f.set_solver(1, R.solver_cfg(**tight))
    f.init()
    out = {}
    for it in range(6):
        if it == 4:
            for nm in ("velocity", "pressure", "facePressure", "massFlux", "continuityResidual"):
                out["s0_" + nm] = f.field(nm).copy()
            ms = f.momentum_system()
            out.update(mom_diag=ms["diag"], mom_off=ms["offdiag"], mom_b=ms["b"],
                       mom_vgrad=f.field("velocityGradient").copy(), mom_pgrad=f.field("pressureGradient").copy())
        f.solve_momentum()
        f.solve_continuity()
    for nm in ("velocity", "pressure", "massFlux"):
        out["end_" + nm] = f.field(nm).copy()
    np.savez_compressed(os.path.join(HERE, "flow_symmetry.npz"), **mesh_arrays(rm), **out)
    print("flow_symmetry.npz: %d cells, group kinds %s" % (rm.n_self, rm.connectivity()["group_kind"]))


def flow_channel_golden():
    """Channel: VelocityBoundary inlet (u = 1), PressureBoundary outlet (p = 0.5), two walls."""
    raw = G.quad_mesh(14, 8, lx=2.0, ly=1.0, jitter=0.15, seed=4)
    rm = R.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes, raw.face_node_count,
                            raw.face_group_size)
    f = R.RefFlow(rm)
    f.set_bc(1, "VelocityBoundary", specifiedXVelocity=1.0)
    f.set_bc(2, "PressureBoundary", specifiedPressure=0.5)
    f.set_bc(3, "NoSlipWall")
    f.set_bc(4, "NoSlipWall")
    f.set_vc("viscosity", 0.05)
    f.set_vc("density", 1.2)
    tight = dict(relativeTolerance=1e-14, nMaxIterations=3000, verbosity=0)
    f.set_solver(0, R.solver_cfg(**tight))
    f.set_solver(1, R.solver_cfg(**tight))
    f.init()
    out = {}
    for it in range(8):
        if it == 3:
            for nm in ("velocity", "pressure", "facePressure", "massFlux", "continuityResidual"):
                out["s0_" + nm] = f.field(nm).copy()
            ms = f.momentum_system()
            out.update(mom_diag=ms["diag"], mom_off=ms["offdiag"], mom_b=ms["b"])
        f.solve_momentum()
        if it == 3:
            for nm in ("velocity", "previousVelocity", "momAp", "pressureGradient", "massFlux", "pressure"):
                out["s1_" + nm] = f.field(nm).copy()
            cs = f.continuity_system()
            out.update(pp_diag=cs["diag"], pp_off=cs["offdiag"], pp_b=cs["b"], pp_is_boundary=cs["is_boundary"])
            f.field("massFlux")[:] = out["s1_massFlux"]   # continuity_system advanced the face fluxes: restore
        f.solve_continuity()
    for nm in ("velocity", "pressure", "facePressure", "massFlux"):
        out["end_" + nm] = f.field(nm).copy()
    np.savez_compressed(os.path.join(HERE, "flow_channel.npz"), **mesh_arrays(rm), **out)
    print("flow_channel.npz: %d cells" % rm.n_self)


def flow_slip_golden():
    """Lid-driven box whose lid and floor are SlipJump walls (F/FlowModelSlipJump.h): a low operating
    pressure makes the mean free path comparable to the wall distance, so the slip is not negligible."""
    raw = G.quad_mesh(11, 9, jitter=0.2, seed=6)
    rm = R.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes, raw.face_node_count,
                            raw.face_group_size)
    f = R.RefFlow(rm)
    f.set_bc(1, "NoSlipWall")
    f.set_bc(2, "NoSlipWall")
    f.set_bc(3, "SlipJump", accomodationCoefficient=0.9)
    f.set_bc(4, "SlipJump", specifiedXVelocity=1.0, accomodationCoefficient=0.8)
    f.set_vc("viscosity", 0.08)
    f.set_vc("density", 1.2)
    f.set_option("operatingPressure", 900.0)
    f.set_option("operatingTemperature", 320.0)
    tight = dict(relativeTolerance=1e-14, nMaxIterations=3000, verbosity=0)
    f.set_solver(0, R.solver_cfg(**tight))
    f.set_solver(1, R.solver_cfg(**tight))
    f.init()
    out = {}
    for it in range(6):
        if it == 3:
            for nm in ("velocity", "pressure", "facePressure", "massFlux", "continuityResidual"):
                out["s0_" + nm] = f.field(nm).copy()
            ms = f.momentum_system()
            out.update(mom_diag=ms["diag"], mom_off=ms["offdiag"], mom_b=ms["b"])
        f.solve_momentum()
        f.solve_continuity()
    for nm in ("velocity", "pressure", "massFlux"):
        out["end_" + nm] = f.field(nm).copy()
    np.savez_compressed(os.path.join(HERE, "flow_slip.npz"), **mesh_arrays(rm), **out)
    v = out["end_velocity"].reshape(-1, 3)
    print("flow_slip.npz: %d cells, max |u| on the slip-wall ghosts %.4f" % (rm.n_self, np.abs(v[rm.n_self:, 0]).max()))


def electric_golden():
    raw = G.hex_mesh(6, 5, 7, lx=1e-6, ly=1e-6, lz=2e-6, jitter=0.15, seed=2)
    rm = R.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes, raw.face_node_count,
                            raw.face_group_size)
    tight = dict(relativeTolerance=1e-13, nMaxIterations=2000, verbosity=0)
    e = R.RefElectric(rm)
    e.set_bc(5, "SpecifiedPotential", specifiedPotential=0.0)
    e.set_bc(6, "SpecifiedPotential", specifiedPotential=100.0)
    e.set_bc(1, "Symmetry")
    e.set_bc(2, "Symmetry")
    e.set_bc(3, "SpecifiedPotentialFlux", specifiedPotentialFlux=1e-3)
    e.set_bc(4, "SpecialDielectricBoundary", specifiedPotential=20.0)
    e.set_option("drift_enable", 1)
    e.set_option("initialTotalCharge", 1e18)
    e.set_option("timeStep", 1e-12)
    e.set_constant("nTrap", 2)
    e.set_constant("electron_mobility", 1e-3)
    e.set_constant("electron_saturation_velocity", 1e5)
    e.set_solver(0, R.solver_cfg(**tight))
    e.set_solver(1, R.solver_cfg(**tight))
    e.init()
    charge0 = 1e15 * (1 + np.arange(raw.n_cells) % 7)
    e.field("charge").reshape(-1, 3)[:raw.n_cells, 2] = charge0
    e.field("chargeN1").reshape(-1, 3)[:] = e.field("charge").reshape(-1, 3)
    out = dict(charge0=charge0)
    for step in range(2):
        _, txt = e.advance(1)
        for nm in ("potential", "electric_field", "electron_velocity", "convectionFlux", "charge"):
            out["s%d_%s" % (step, nm)] = e.field(nm).copy()
        out["s%d_text" % step] = np.array(txt)
        e.update_time()
    np.savez_compressed(os.path.join(HERE, "electric_box.npz"), **mesh_arrays(rm), nodes=raw.nodes,
                        face_nodes=raw.face_nodes, face_node_count=raw.face_node_count,
                        face_group_size=raw.face_group_size, **out)
    print("electric_box.npz: %d cells" % rm.n_self)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "flowchan":
        flow_channel_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "flowslip":
        flow_slip_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "flowsym":
        flow_symmetry_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "electric":
        electric_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "flow":
        flow_golden()
        sys.exit(0)
    main()
