"""bench.py's secondary workloads (BASELINE.json configs[2] and configs[4]); the headline stays `--workload thermal`.

  cavity        lid-driven cavity, FlowModelA (SIMPLE: momentum + Rhie-Chow pressure correction, both solved by AMG
                with the reference's default inner tolerances) on a synthetic n x n quad mesh (default 2048^2).
                One STEP = one SIMPLE iteration. Single GPU.
  electric-tet  ElectricModelA (electrostatics + drift / transient charge transport, nTrap = 2; BCGStab preconditioned
                by one AMG cycle, as the reference's own dielectric-charging script) on the unit box cut into 6
                jittered tetrahedra per hex (default 96^3 x 6 = 5.3 M tets per GPU-octet... see --size), partitioned
                into blocks (what coordinate bisection gives on a uniform box), one part per GPU.
                One STEP = one time step (advance(1) + updateTime).

Same JSON contract as the headline line: value = device-timed (CUDA events around assembly + solve + update of every
equation, fields resident), e2e = wall clock around the public advance() with host numpy fields (H2D / D2H inside),
roofline of the level-0 Gauss-Seidel pass from a profiled extra step, cpu_baseline = the reference's own model
(oracle/_ref) on a bounded sample, parity = a small case of the same model against the reference.
"""
import contextlib
import io
import json
import os
import time

import numpy as np

import bench as B


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _dist():
    import torch
    import torch.distributed as dist
    from fvm_b200 import capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.default_lib()
    if world > 1:
        capi.init_comm_from_torch(lib)
    return lib, rank, world, local


def _allmax(v, world):
    if world == 1:
        return float(v)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _allsum(vals, world):
    if world == 1:
        return [float(v) for v in vals]
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v) for v in vals], device="cuda", dtype=torch.float64)
    dist.all_reduce(t)
    return [float(v) for v in t.cpu()]


def _barrier(lib, world):
    lib.synchronize()
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()


def _roofline(recs, step_ms, nnz_per_row):
    """Level-0 Gauss-Seidel launches of all systems of the step (one launch = one colour class)."""
    peak, src = B.hbm_peak()
    by = {}
    for r in recs:
        d = by.setdefault(r["name"], dict(launches=0, ms=0.0))
        d["launches"] += r["launches"]; d["ms"] += r["ms"]
    total = sum(d["ms"] for d in by.values()) or 1.0
    table = sorted(({"kernel": k, "launches": v["launches"], "ms": round(v["ms"], 3), "share": round(v["ms"] / total, 4)}
                    for k, v in by.items()), key=lambda t: -t["ms"])[:12]
    picked = [r for r in recs if r["name"] == "GsRows" and r.get("level", -1) == 0]
    launches = sum(r["launches"] for r in picked)
    ms = sum(r["ms"] for r in picked)
    if not launches:
        return None, table
    per_row = 36.0 + 12.0 * nnz_per_row
    total_bytes = sum(r["launches"] * r["rows"] * per_row for r in picked)
    achieved = total_bytes / (ms * 1e-3) / 1e9
    return ({"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
             "kernel": "k_rows<GsRows> at AMG level 0 (one colour of the multicolour Gauss-Seidel sweep), all systems of the step",
             "peak_source": src, "bytes_per_launch": total_bytes / launches, "bytes_per_row": per_row,
             "rows_per_launch": sum(r["launches"] * r["rows"] for r in picked) / launches,
             "mean_launch_ms": ms / launches, "launches": launches, "share_of_step": ms / max(step_ms, 1e-12)}, table)


# ----------------------------------------------------------------------------- cavity
def _cavity_model(lib, n, mu):
    from fvm_b200 import meshgen as G, models as M
    raw = G.quad_mesh(n, n)
    mesh = M.Mesh(raw)
    geom = M.GeomFields("geom")
    M.MeshMetricsCalculatorA(geom, [mesh], lib=lib).init()
    ff = M.FlowFields("flow")
    fm = M.FlowModelA(geom, ff, [mesh], lib=lib)
    fm.getBCMap()[4]["specifiedXVelocity"] = 1.0          # the lid (y = top) moves along x
    fm.getVCMap()[mesh.getID()]["viscosity"] = mu
    fm.getOptions().momentumTolerance = 1e-30              # never "converged": every step is a full SIMPLE iteration
    fm.getOptions().continuityTolerance = 1e-30
    fm.init()
    return raw, mesh, ff, fm


def _cavity_reference(n, mu, iters, tight, pressure_kind=0):
    """oracle/_ref: the reference's FlowModel<double> on the same cavity. Returns (seconds per iteration, fields).
    pressure_kind: refapi solver kind of the pressure-correction solver when tight (0 = AMG, 1 = BCGStab + AMG)."""
    from fvm_b200 import meshgen as G
    from oracle import refapi as R
    raw = G.quad_mesh(n, n)
    rm = R.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes, raw.face_node_count,
                            raw.face_group_size)
    f = R.RefFlow(rm)
    for g in (1, 2, 3):
        f.set_bc(g, "NoSlipWall")
    f.set_bc(4, "NoSlipWall", specifiedXVelocity=1.0)
    f.set_vc("viscosity", mu)
    f.set_vc("density", 1.0)
    f.set_option("momentumTolerance", 1e-30)
    f.set_option("continuityTolerance", 1e-30)
    if tight:
        cfg = dict(relativeTolerance=1e-13, nMaxIterations=3000, verbosity=0)
        f.set_solver(0, R.solver_cfg(**cfg))
        f.set_solver(1, R.solver_cfg(kind=pressure_kind, **cfg))
    f.init()
    t0 = time.perf_counter()
    f.advance(iters)
    dt = (time.perf_counter() - t0) / iters
    out = dict(velocity=f.field("velocity").copy(), pressure=f.field("pressure").copy(), n_cells=raw.n_cells)
    f.close()
    return dt, out


def run_cavity(args):
    from fvm_b200 import models as M
    from oracle import refapi
    lib, rank, world, local = _dist()
    if world != 1:
        raise SystemExit("--workload cavity runs on one GPU (FlowModelA on mesh parts is covered by tests/test_multigpu.py)")
    n = args.n or 2048
    mu = 0.01
    t0 = time.time()
    raw, mesh, ff, fm = _cavity_model(lib, n, mu)
    setup_s = time.time() - t0
    ncells = raw.n_cells
    with _quiet():
        fm.advance(args.warmup)
    lib.flush_l2()
    sampler = B.ClockSampler(local)
    sampler.start()
    l0 = lib.counters()[0]
    k0 = len(fm.timings)
    _barrier(lib, world)
    with _quiet():
        fm.advance(args.steps)          # fields stay on the device between the iterations of one call
    _barrier(lib, world)
    l1 = lib.counters()[0]
    clocks = sampler.stop()
    tm = fm.timings[k0:]
    keys = ("momentum_assemble_ms", "momentum_solve_ms", "continuity_assemble_ms", "continuity_solve_ms")
    step_ms = [sum(t[k] for k in keys) for t in tm]
    total_ms = float(sum(step_ms))
    # e2e: one public advance(1) per step -- host fields up, one SIMPLE iteration, host fields back
    e2e_t, h0 = [], lib.counters()
    for i in range(max(1, min(args.steps, 3)) + 1):
        lib.synchronize()
        t0 = time.perf_counter()
        with _quiet():
            fm.advance(1)
        lib.synchronize()
        if i > 0:
            e2e_t.append(time.perf_counter() - t0)
        else:
            h0 = lib.counters()
    h1 = lib.counters()
    e2e_s = float(np.mean(e2e_t))
    # parity: a 64^2 cavity, 3 SIMPLE iterations with converged inner solves, against the reference's FlowModel
    parity = {"oracle_check": {"oracle": "unavailable"}}
    if refapi.available() and args.parity_size > 0:
        pn = min(args.parity_size, 96)
        _, ref = _cavity_reference(pn, mu, 3, tight=True)
        _, m2, f2, fm2 = _cavity_model(lib, pn, mu)
        for nm in ("momentumLinearSolver", "pressureLinearSolver"):
            s = M.AMG()
            s.relativeTolerance, s.nMaxIterations, s.verbosity = 1e-13, 3000, 0
            setattr(fm2.getOptions(), nm, s)
        with _quiet():
            fm2.advance(3)
        c2 = m2.getCells()
        nn = ref["n_cells"]
        v = np.asarray(f2.velocity[c2]).reshape(-1, 3)[:nn]
        vr = ref["velocity"].reshape(-1, 3)[:nn]
        pr = np.asarray(f2.pressure[c2])[:nn]
        ev = float(np.linalg.norm(v - vr) / np.linalg.norm(vr))
        ep = float(np.linalg.norm(pr - ref["pressure"][:nn]) / max(np.linalg.norm(ref["pressure"][:nn]), 1e-300))
        parity = {"oracle_check": {"oracle": "reference (oracle/_ref FlowModel<double>)",
                                   "case": "%d^2 cavity, 3 SIMPLE iterations, inner solves to rel 1e-13 on both sides" % pn,
                                   "velocity_rel_l2": ev, "pressure_rel_l2": ep, "bar": 1e-8, "pass": bool(ev <= 1e-8 and ep <= 1e-8)}}
    # roofline: a profiled extra iteration
    roof = table = None
    if not args.no_profile:
        fl = fm._flows[mesh.getID()]
        lib.profile_begin()
        kp = len(fm.timings)
        with _quiet():
            fm.advance(1)
        recs = lib.profile_end(cap=8192)
        roof, table = _roofline(recs, sum(fm.timings[kp][k] for k in keys), 4.0)
    cpu = None
    if refapi.available() and not args.no_cpu_baseline:
        rn = args.ref_n or 512
        dt, _ = _cavity_reference(rn, mu, 3, tight=False)
        cpu = {"value": rn * rn / dt, "unit": B.UNIT, "cores": 1, "kind": "reference",
               "sample": "%d^2 cavity (%d cells), 3 SIMPLE iterations of the reference's FlowModel<double> with its default "
                         "solvers, one core" % (rn, rn * rn), "seconds_per_step": dt}
    last = tm[-1]
    out = {"metric": B.METRIC, "value": ncells * args.steps / max(total_ms * 1e-3, 1e-12), "unit": B.UNIT, "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "lid-driven cavity, FlowModel SIMPLE (momentum + Rhie-Chow pressure correction), %dx%d quads "
                                  "(%d cells), rho=1, mu=%g, lid u=1, URFs 0.7/0.3, AMG inner solves with the reference's "
                                  "default tolerances; one SIMPLE iteration per step" % (n, n, ncells, mu),
                      "cells_per_gpu": ncells, "parallelism": "single",
                      "l2": "matrix + fields of one system (~0.6 GB) exceed the 126 MB L2; L2 flushed after warm-up"},
           "step_ms": [round(v, 3) for v in step_ms],
           "phase_ms": {k: float(np.mean([t[k] for t in tm])) for k in keys},
           "inner_cycles": {"momentum": [int(i) for i in last["momentum_iterations"]], "pressure": int(last["pressure_iterations"])},
           "residual": {"momentum": [float(v) for v in last["momentum_norm"]], "continuity": float(last["continuity_norm"])},
           "gpu_launches": int(l1 - l0),
           "e2e": {"value": ncells / e2e_s, "unit": B.UNIT, "h2d_bytes_per_step": int((h1[1] - h0[1]) / len(e2e_t)),
                   "d2h_bytes_per_step": int((h1[2] - h0[2]) / len(e2e_t)), "seconds_per_step": e2e_s},
           "parity": parity, "clocks": clocks, "mesh_setup_s": setup_s}
    if roof:
        out["roofline"], out["kernel_profile"] = roof, table
    if cpu:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))


# ----------------------------------------------------------------------------- electric, tets
E_BOX = 1e-6   # metres: the box is 1 x 1 x 1 micron (the scale of the reference's dielectric-charging case)


def _electric_setup(em, M, tol, iters):
    bc = em.getBCMap()
    for gid, b in bc.items():
        b.bcType = "Symmetry"
    if 5 in bc:
        bc[5].bcType = "SpecifiedPotential"; bc[5]["specifiedPotential"] = 0.0
    if 6 in bc:
        bc[6].bcType = "SpecifiedPotential"; bc[6]["specifiedPotential"] = 100.0
    o = em.getOptions()
    o.drift_enable = True
    o["initialTotalCharge"] = 1e18
    o["timeStep"] = 1e-12
    c = em.getConstants()
    c["nTrap"] = 2
    c["electron_mobility"] = 1e-3
    c["electron_saturation_velocity"] = 1e5
    for nm in ("electrostaticsLinearSolver", "chargetransportLinearSolver"):
        pc = M.AMG()
        pc.verbosity = 0
        s = M.BCGStab()
        s.preconditioner = pc
        s.relativeTolerance, s.nMaxIterations, s.verbosity = tol, iters, 0
        setattr(o, nm, s)


def _electric_model(lib, raw, tol=1e-8, iters=100, uniform_field=False):
    """uniform_field: no space charge and the potential initialised with its exact (linear) solution, so that the
    electron velocity is the same in every cell (see the parity note in run_electric)."""
    from fvm_b200 import models as M
    mesh = M.Mesh(raw)
    geom = M.GeomFields("geom")
    M.MeshMetricsCalculatorA(geom, [mesh], lib=lib).init()
    ef = M.ElectricFields("elec")
    em = M.ElectricModelA(geom, ef, [mesh], lib=lib)
    _electric_setup(em, M, tol, iters)
    em.init()
    cells = mesh.getCells()
    if uniform_field:
        ef.total_charge[cells][:] = 0.0
        ef.potential[cells][:] = 100.0 * np.asarray(geom.coordinate[cells])[:, 2] / E_BOX
    gids = raw.cell_global if "cell_global" in raw else np.arange(raw.n_total)
    own = np.arange(raw.n_total) < raw.n_cells
    ef.charge[cells][:, 2] = np.where(own, 1e15 * (1 + np.maximum(gids, 0) % 7), 0.0)
    ef.chargeN1[cells][:] = ef.charge[cells]
    return mesh, ef, em


def _electric_reference(raw, steps, tol, iters=2000, kind=1, uniform_field=False):
    """oracle/_ref: the reference's ElectricModel<double> on the same (single-partition) mesh."""
    from oracle import refapi as R
    rm = R.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes, raw.face_node_count,
                            raw.face_group_size)
    e = R.RefElectric(rm)
    for g in (1, 2, 3, 4):
        e.set_bc(g, "Symmetry")
    e.set_bc(5, "SpecifiedPotential", specifiedPotential=0.0)
    e.set_bc(6, "SpecifiedPotential", specifiedPotential=100.0)
    e.set_option("drift_enable", 1)
    e.set_option("initialTotalCharge", 1e18)
    e.set_option("timeStep", 1e-12)
    e.set_constant("nTrap", 2)
    e.set_constant("electron_mobility", 1e-3)
    e.set_constant("electron_saturation_velocity", 1e5)
    cfg = dict(kind=kind, relativeTolerance=tol, nMaxIterations=iters, verbosity=0)
    e.set_solver(0, R.solver_cfg(**cfg))
    e.set_solver(1, R.solver_cfg(**cfg))
    e.init()
    if uniform_field:
        from fvm_b200 import meshgen as G
        e.field("total_charge")[:] = 0.0
        e.field("potential")[:] = 100.0 * G.metrics(raw)["cell_centroid"][:, 2] / E_BOX
    e.field("charge").reshape(-1, 3)[:raw.n_cells, 2] = 1e15 * (1 + np.arange(raw.n_cells) % 7)
    e.field("chargeN1").reshape(-1, 3)[:] = e.field("charge").reshape(-1, 3)
    t0 = time.perf_counter()
    for _ in range(steps):
        e.advance(1)
        e.update_time()
    dt = (time.perf_counter() - t0) / steps
    out = dict(potential=e.field("potential").copy(), charge=e.field("charge").reshape(-1, 3).copy())
    e.close()
    return dt, out


def run_electric(args):
    import torch
    import torch.distributed as dist
    from fvm_b200 import meshgen as G, partition as P
    from oracle import refapi
    lib, rank, world, local = _dist()
    n = args.n or 96                      # hexes per side of the GLOBAL box (6 tets each)
    t0 = time.time()
    if world == 1:
        raw = G.tet_mesh(n, n, n, lx=E_BOX, ly=E_BOX, lz=E_BOX)
    else:
        raw = P.tet_block(n, n, n, rank, world, lx=E_BOX, ly=E_BOX, lz=E_BOX)
    mesh, ef, em = _electric_model(lib, raw)
    setup_s = time.time() - t0
    ncells = raw.n_cells
    total_cells = int(_allsum([ncells], world)[0])

    def step():
        with _quiet():
            em.advance(1)
        em.updateTime()

    for _ in range(args.warmup):
        step()
    lib.flush_l2()
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.counters()
    c0 = lib.comm_collectives()
    k0 = len(em.timings)
    e2e_t = []
    _barrier(lib, world)
    for _ in range(args.steps):
        t1 = time.perf_counter()
        step()
        lib.synchronize()
        e2e_t.append(time.perf_counter() - t1)
    _barrier(lib, world)
    l1 = lib.counters()
    c1 = lib.comm_collectives()
    clocks = sampler.stop() if rank == 0 else None
    tm = em.timings[k0:]
    step_ms = [t.get("electrostatics_ms", 0.0) + t.get("charge_ms", 0.0) for t in tm]
    total_ms = _allmax(sum(step_ms), world)
    e2e_s = _allmax(float(np.mean(e2e_t)), world)
    # parity: the same model on a small box, THIS run's ranks as one partitioned problem, against the reference.
    # The reference's updateConvectionFlux gives boundary face k of a non-symmetry group the velocity of cell c0 of face
    # k of the whole mesh (F/ElectricModel_impl.h:1070-1088; reproduced in csrc/electric.cu): the charge then depends on
    # the face numbering, hence on the partition, in the reference as here. On several ranks the charge is therefore
    # compared in a second run whose field is uniform (potential started from its exact linear solution, no space
    # charge: every cell has the same velocity); the potential is compared in the workload's own configuration.
    parity = {"oracle_check": {"oracle": "unavailable"}}
    if args.parity_size > 0:
        pn = max(4, min(args.parity_size, 16))
        praw = G.tet_mesh(pn, pn, pn, lx=E_BOX, ly=E_BOX, lz=E_BOX)
        nt = praw.n_total
        cases = [False] if world == 1 else [False, True]      # uniform_field of _electric_model
        ref_pack = np.zeros(2 * nt * len(cases) + 1)
        if rank == 0 and refapi.available():
            for i, uf in enumerate(cases):
                _, ref = _electric_reference(praw, 2, 1e-13, kind=0, uniform_field=uf)
                ref_pack[2 * nt * i:2 * nt * i + nt] = ref["potential"]
                ref_pack[2 * nt * i + nt:2 * nt * (i + 1)] = ref["charge"][:, 2]
            ref_pack[-1] = 1
        if world > 1:
            pack = torch.from_numpy(ref_pack).cuda()
            dist.broadcast(pack, 0)
            ref_pack = pack.cpu().numpy()
            ploc = P.tet_block(pn, pn, pn, rank, world, lx=E_BOX, ly=E_BOX, lz=E_BOX)
        else:
            ploc = praw
        if int(ref_pack[-1]):
            own = (ploc.cell_global if "cell_global" in ploc else np.arange(ploc.n_total))[:ploc.n_cells]
            errs = []
            for i, uf in enumerate(cases):
                ref_pot, ref_chg = ref_pack[2 * nt * i:2 * nt * i + nt], ref_pack[2 * nt * i + nt:2 * nt * (i + 1)]
                m2, f2, e2 = _electric_model(lib, ploc, tol=1e-13, iters=500, uniform_field=uf)
                for _ in range(2):
                    with _quiet():
                        e2.advance(1)
                    e2.updateTime()
                c2 = m2.getCells()
                pot = np.asarray(f2.potential[c2])[:ploc.n_cells]
                chg = np.asarray(f2.charge[c2])[:ploc.n_cells, 2]
                s = _allsum([((pot - ref_pot[own]) ** 2).sum(), (ref_pot[own] ** 2).sum(),
                             ((chg - ref_chg[own]) ** 2).sum(), (ref_chg[own] ** 2).sum()], world)
                errs.append((float(np.sqrt(s[0] / s[1])), float(np.sqrt(s[2] / s[3]))))
            ep, ec = errs[0][0], errs[-1][1]
            oc = {"oracle": "reference (oracle/_ref ElectricModel<double>, single partition)",
                  "case": "%d^3 x 6 jittered tets in %d part(s), 2 time steps, solvers to rel 1e-13" % (pn, world),
                  "potential_rel_l2": ep, "charge_rel_l2": ec, "bar": 1e-8, "pass": bool(ep <= 1e-8 and ec <= 1e-8)}
            if world > 1:
                oc["charge_case"] = ("uniform field (linear initial potential, no space charge): the reference's boundary "
                                     "drift flux depends on the face numbering, F/ElectricModel_impl.h:1070-1088")
                oc["charge_rel_l2_workload_configuration"] = errs[0][1]
            parity = {"oracle_check": oc}
    roof = table = None
    if not args.no_profile:
        lib.profile_begin()
        kp = len(em.timings)
        step()
        recs = lib.profile_end(cap=8192)
        t = em.timings[kp]
        roof, table = _roofline(recs, t.get("electrostatics_ms", 0.0) + t.get("charge_ms", 0.0), 4.0)
    cpu = None
    if rank == 0 and refapi.available() and not args.no_cpu_baseline:
        rn = args.ref_n or 24
        rraw = G.tet_mesh(rn, rn, rn, lx=E_BOX, ly=E_BOX, lz=E_BOX)
        dt, _ = _electric_reference(rraw, 1, 1e-8, iters=100, kind=1)
        cpu = {"value": rraw.n_cells / dt, "unit": B.UNIT, "cores": 1, "kind": "reference",
               "sample": "%d^3 x 6 tets (%d cells), one time step of the reference's ElectricModel<double>, BCGStab + AMG to "
                         "rel 1e-8, one core" % (rn, rraw.n_cells), "seconds_per_step": dt}
    if rank != 0:
        return
    last = tm[-1]
    out = {"metric": B.METRIC, "value": total_cells * args.steps / max(total_ms * 1e-3, 1e-12), "unit": B.UNIT,
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "ElectricModel charge transport: electrostatics (eps_r 7.9, 100 V across z, symmetry elsewhere) + "
                                  "drift / transient transport of the charge vector (nTrap = 2), %d^3 hexes cut into 6 jittered "
                                  "tetrahedra (%d cells), BCGStab preconditioned by one AMG cycle to rel 1e-8 for every system; one "
                                  "time step per step" % (n, total_cells),
                      "cells_per_gpu": ncells,
                      "parallelism": ("block partition %s (coordinate bisection of the uniform box), one part per GPU, NVLink "
                                      "peer-memory halo exchange, all-reduced dots and norms" % "x".join(map(str, P.block_dims(world))))
                      if world > 1 else "single",
                      "l2": "matrix + vectors of one system exceed the 126 MB L2; L2 flushed after warm-up"},
           "step_ms": [round(v, 3) for v in step_ms],
           "phase_ms": {"electrostatics_ms": float(np.mean([t.get("electrostatics_ms", 0.0) for t in tm])),
                        "charge_ms": float(np.mean([t.get("charge_ms", 0.0) for t in tm]))},
           "krylov_iterations_electrostatics": int(last.get("electrostatics_iterations", 0)),
           "gpu_launches": int(l1[0] - l0[0]),
           "e2e": {"value": total_cells / e2e_s, "unit": B.UNIT, "h2d_bytes_per_step": int((l1[1] - l0[1]) / args.steps),
                   "d2h_bytes_per_step": int((l1[2] - l0[2]) / args.steps), "seconds_per_step": e2e_s},
           "parity": parity, "clocks": clocks, "mesh_setup_s": setup_s}
    if world > 1:
        out["collectives_per_step"] = int((c1 - c0) / args.steps)
    if roof:
        out["roofline"], out["kernel_profile"] = roof, table
    if cpu:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
