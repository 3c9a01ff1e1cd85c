"""Worker of tests/test_multirank.py: one process per rank (torch.distributed, gloo, CPU).

Runs the PRODUCT's multi-rank code path -- partitioned meshes, halo exchange around the assembly,
distributed AMG / BCGStab with per-level ghost maps and the merged (replicated) coarse level -- in
the test-only host simulator build of the same sources (tests/hostsim, FVMGPU_HOSTSIM), whose
transport is a set of callbacks implemented here with torch.distributed. The solution of every
rank is compared with the single-partition oracle on the global mesh (north_star: 1e-8 rel L2).
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

from fvm_b200 import build, capi as X, meshgen as G, partition as P

EXCH = C.CFUNCTYPE(None, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double))
ALLR = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int)
ALLG = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_longlong)


def _exchange(n, peer, so, sc, send, ro, rc, recv):
    reqs, landing = [], []
    for i in range(n):
        if sc[i]:
            t = torch.from_numpy(np.ctypeslib.as_array(send, shape=(so[i] + sc[i],))[so[i]:so[i] + sc[i]].copy())
            reqs.append(dist.isend(t, int(peer[i])))
        if rc[i]:
            buf = torch.empty(rc[i], dtype=torch.float64)
            reqs.append(dist.irecv(buf, int(peer[i])))
            landing.append((buf, ro[i], rc[i]))
    for r in reqs:
        r.wait()
    for buf, o, c in landing:
        np.ctypeslib.as_array(recv, shape=(o + c,))[o:o + c] = buf.numpy()


def _allreduce(data, n):
    a = np.ctypeslib.as_array(data, shape=(n,))
    t = torch.from_numpy(a.copy())
    dist.all_reduce(t)
    a[:] = t.numpy()


def _allgather(send, recv, nbytes):
    world = dist.get_world_size()
    s = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_ubyte)), shape=(nbytes,)).copy()
    outs = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(outs, torch.from_numpy(s))
    r = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_ubyte)), shape=(nbytes * world,))
    for k, o in enumerate(outs):
        r[k * nbytes:(k + 1) * nbytes] = o.numpy()


_keep = (EXCH(_exchange), ALLR(_allreduce), ALLG(_allgather))


def rejoin(lib, world, rank):
    """Join the library's communicator again after a single-partition run (hostsim: callbacks; GPU: a fresh
    NCCL communicator bootstrapped through torch.distributed)."""
    if os.environ.get("FVM_WORKER_GPU") == "1":
        X.init_comm_from_torch(lib)
    else:
        lib.comm_init(world, rank)


def make_case(name):
    if name == "hex_slabs":
        raw = G.hex_mesh(10, 9, 12, jitter=0.15, seed=3)
        method = "slabs"
    elif name == "hex_box":      # uniform, even dimensions: the hierarchy must stay structured on every rank
        raw = G.hex_mesh(16, 16, 16)
        method = "slabs"
    elif name == "tet_rcb":
        raw = G.tet_mesh(6, 5, 7, jitter=0.2, seed=7)
        method = "rcb"
    else:
        raise ValueError(name)
    return raw, method


def main():
    case, solver_kind = sys.argv[1], sys.argv[2]
    merge_rows = sys.argv[3] if len(sys.argv) > 3 else None
    if merge_rows:
        os.environ["FVMGPU_MERGE_ROWS"] = merge_rows
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if os.environ.get("FVM_WORKER_GPU") == "1":
        # the product library on one B200 per rank, halo exchange over NCCL (tests/test_multigpu.py)
        lib = X.default_lib()
        X.init_comm_from_torch(lib)
    else:
        lib = X.Lib(build.HOSTSIM_LIB)
        lib.init(0)
        lib.dll.fvmgpu_hostsim_set_comm.restype = C.c_int
        lib.dll.fvmgpu_hostsim_set_comm(*_keep)
        lib.comm_init(world, rank)

    if solver_kind == "electric_bench":
        # bench.py --workload electric-tet in miniature: the bench's own model set-up (Symmetry walls, BCGStab + AMG
        # for both systems, its charge initial condition) on the block partition it uses (partition.tet_block),
        # against the single-partition run of the same code.
        # The reference's updateConvectionFlux gives boundary face k of a non-symmetry group the velocity of cell c0 of
        # face k OF THE WHOLE MESH (F/ElectricModel_impl.h:1070-1088, reproduced): its result depends on the face
        # numbering, i.e. on the partition -- in the reference as here. The partitioned run can therefore be compared
        # with the single-partition one (i) in the potential, always, (ii) in the charge when the field is uniform (the
        # velocity is then the same in every cell): potential initialised with the exact linear profile and no
        # space charge. With the non-uniform field the interface fluxes are checked against their definition instead.
        import contextlib
        import io
        import bench_workloads as W
        pn = int(case)
        kw = dict(lx=W.E_BOX, ly=W.E_BOX, lz=W.E_BOX)
        graw = G.tet_mesh(pn, pn, pn, **kw)

        def run(mesh_raw, uniform_field):
            m2, f2, e2 = W._electric_model(lib, mesh_raw, tol=1e-13, iters=500, uniform_field=uniform_field)
            for _ in range(2):
                with contextlib.redirect_stdout(io.StringIO()):
                    e2.advance(1)
                e2.updateTime()
            c2 = m2.getCells()
            run.model = (m2, f2, e2)
            return np.asarray(f2.potential[c2]).copy(), np.asarray(f2.charge[c2])[:, 2].copy()

        lib.comm_destroy()
        ref = {u: run(graw, u) for u in (False, True)}
        rejoin(lib, world, rank)
        ploc = P.tet_block(pn, pn, pn, rank, world, **kw) if world > 1 else graw
        own = (ploc.cell_global if "cell_global" in ploc else np.arange(ploc.n_total))[:ploc.n_cells]
        nv = ploc.n_cells
        pot, _ = run(ploc, False)
        # interface faces: flux = 0.5 (v_own + v_ghost) . A with the owner's velocity in the ghost cell
        m2, f2, e2 = run.model
        vel = np.asarray(f2.electron_velocity[m2.getCells()])
        flux = np.asarray(f2.convectionFlux[m2.getFaces()])
        area = np.asarray(e2.geom.area[m2.getFaces()])
        fc = np.asarray(ploc.face_cells).reshape(-1, 2)
        iface_err, n_iface = 0.0, 0
        for gi in range(len(ploc.group_id)):
            if int(ploc.group_kind[gi]) != X.GROUP_INTERFACE:
                continue
            o, c = int(ploc.group_offset[gi]), int(ploc.group_count[gi])
            want = 0.5 * (np.einsum("ij,ij->i", vel[fc[o:o + c, 0]], area[o:o + c]) +
                          np.einsum("ij,ij->i", vel[fc[o:o + c, 1]], area[o:o + c]))
            iface_err = max(iface_err, float(np.abs(flux[o:o + c] - want).max() / np.abs(want).max()))
            n_iface += c
        _, chg = run(ploc, True)
        t = torch.tensor([((pot[:nv] - ref[False][0][own]) ** 2).sum(), (ref[False][0][own] ** 2).sum(),
                          ((chg[:nv] - ref[True][1][own]) ** 2).sum(), (ref[True][1][own] ** 2).sum()])
        dist.all_reduce(t)
        out = dict(rank=rank, world=world, n_self=int(nv), pot_rel_l2=float(torch.sqrt(t[0] / t[1])),
                   chg_rel_l2=float(torch.sqrt(t[2] / t[3])), iface_flux_err=iface_err, n_iface=n_iface)
        with open(os.path.join(os.environ["FVM_RESULT_DIR"], "rank%d.json" % rank), "w") as fh:
            json.dump(out, fh)
        dist.barrier()
        dist.destroy_process_group()
        return

    if solver_kind == "thermgold":
        # T/PARALLEL_TESTS CAVITY_*_JACOBISOLVER: the golden (iteration count + last residual of a Jacobi-smoothed
        # solve to rel 1e-5) is the same file for every rank count the reference registered (1 ... 47)
        import contextlib
        import io
        from fvm_b200 import importers, models as M
        cas, golden = {"cav32": ("cav32.cas", "QUAD_1024"), "tri894": ("tri_894.cas", "TRI_894"),
                       "tetra8k": ("cav_tetra.cas", "TETRA_8K")}[case]
        fc = importers.FluentCase("/root/reference/src/fvm/test/" + cas)
        fc.read()
        raw0 = fc.getMeshList()[0].raw
        geo0 = G.metrics(raw0)
        part = (P.assign_slabs(raw0.n_cells, world) if case == "cav32"
                else P.assign_rcb(geo0["cell_centroid"][:raw0.n_cells], world))
        loc = P.partition_mesh(raw0, geo0, part, rank)
        mesh = M.Mesh(loc)
        geomf = M.GeomFields("geom")
        M.MeshMetricsCalculatorA(geomf, [mesh], lib=lib).init()
        tf = M.ThermalFields("therm")
        tm = M.ThermalModelA(geomf, tf, [mesh], lib=lib)
        bcm = tm.getBCMap()
        if 3 in bcm:
            bcm[3].bcType = "SpecifiedTemperature"; bcm[3].setVar("specifiedTemperature", 400)
        for gid in (4, 5, 6):
            if gid in bcm:
                bcm[gid].bcType = "SpecifiedTemperature"; bcm[gid].setVar("specifiedTemperature", 0)
        for vc in tm.getVCMap().values():
            vc.setVar("thermalConductivity", 1.0)
        sv = M.AMG()
        sv.smootherType, sv.maxCoarseLevels = 1, 0
        sv.relativeTolerance, sv.nMaxIterations, sv.verbosity = 1e-5, 20000, 0
        tm.getOptions().linearSolver = sv
        tm.init()
        with contextlib.redirect_stdout(io.StringIO()):
            tm.advance(1)
        gold = open("/root/reference/src/fvm/test/PARALLEL_TESTS/SOLVER_JACOBI/%s/proc2/GOLDEN/convergence.dat" % golden)
        last = gold.read().splitlines()[1]
        out = dict(rank=rank, world=world, n_self=int(loc.n_cells), peers=[int(p) for p in loc.halo["peers"]],
                   err_diag=0.0, err_b=0.0, rel_l2=0.0, ghost_err=0.0, r0=1.0, r=0.0, iters=int(sv.lastIterations),
                   levels=[], collectives=lib.comm_collectives(), golden_last=last,
                   ours_last="%d: [therm.temperature : %g]" % (sv.lastIterations, sv.lastResidual))
        with open(os.path.join(os.environ["FVM_RESULT_DIR"], "rank%d.json" % rank), "w") as fh:
            json.dump(out, fh)
        dist.barrier()
        dist.destroy_process_group()
        return

    if solver_kind == "flowgold":
        # T/PARALLEL_CAVITY_JACOBI: cav32.cas, lid u = 1, both systems relaxed by Jacobi sweeps on the fine level only
        # (AMG with smootherType = JACOBI, maxCoarseLevels = 0, rel 1e-1 / 200). Jacobi with a ghost refresh after every
        # pass does not depend on the partition -- the reference's goldens for 1, 4, 16 and 64 ranks are the same file --
        # so this rank count must reproduce that file too.
        import contextlib
        import io
        import re
        from fvm_b200 import importers, models as M
        cas = "/root/reference/src/fvm/test/cav32.cas"
        gold_path = "/root/reference/src/fvm/test/PARALLEL_CAVITY_JACOBI/PROC4/GOLDEN/convergence.dat"
        fc = importers.FluentCase(cas)
        fc.read()
        raw0 = fc.getMeshList()[0].raw
        geo0 = G.metrics(raw0)
        loc = P.partition_mesh(raw0, geo0, P.assign_slabs(raw0.n_cells, world), rank)
        mesh = M.Mesh(loc)
        geomf = M.GeomFields("geom")
        M.MeshMetricsCalculatorA(geomf, [mesh], lib=lib).init()
        ff = M.FlowFields("flow")
        fm = M.FlowModelA(geomf, ff, [mesh], lib=lib)
        bcm = fm.getBCMap()
        for gid, bc in bcm.items():
            bc.bcType = "NoSlipWall"
        if 3 in bcm:
            bcm[3].setVar("specifiedXVelocity", 1)
        for vc in fm.getVCMap().values():
            vc.setVar("density", 1.0); vc.setVar("viscosity", 0.1)
        fo = fm.getOptions()
        for nm in ("momentumLinearSolver", "pressureLinearSolver"):
            sv = M.AMG()
            sv.smootherType, sv.maxCoarseLevels = 1, 0
            sv.relativeTolerance, sv.nMaxIterations, sv.verbosity = 1e-1, 200, 0
            setattr(fo, nm, sv)
        fo.momentumTolerance = fo.continuityTolerance = 1e-5
        fm.init()
        with contextlib.redirect_stdout(io.StringIO()):
            fm.advance(10)
        ours = np.array([[t["momentum_norm"][0], t["momentum_norm"][1], t["continuity_norm"]] for t in fm.timings])
        gold = np.array([[float(x) for x in re.findall(r"[-+]?\d+\.?\d*(?:e[-+]?\d+)?", l.split(":", 1)[1])]
                         for l in open(gold_path).read().splitlines()])[:, [0, 1, 3]]
        dev = np.abs(ours - gold) / np.maximum(np.abs(gold), 1e-300)
        dev[0, 1] = 0.0
        out = dict(rank=rank, world=world, n_self=int(loc.n_cells), peers=[int(p) for p in loc.halo["peers"]],
                   err_diag=0.0, err_b=0.0, rel_l2=0.0, ghost_err=0.0, r0=1.0, r=0.0, iters=0, levels=[],
                   collectives=lib.comm_collectives(), golden_dev=float(dev.max()), rows=int(len(ours)),
                   y0=float(ours[0, 1]))
        with open(os.path.join(os.environ["FVM_RESULT_DIR"], "rank%d.json" % rank), "w") as fh:
            json.dump(out, fh)
        dist.barrier()
        dist.destroy_process_group()
        return

    raw, method = make_case(case)
    geo = G.metrics(raw)
    part = (P.assign_slabs(raw.n_cells, world) if method == "slabs"
            else P.assign_rcb(geo["cell_centroid"][:raw.n_cells], world))
    loc = P.partition_mesh(raw, geo, part, rank)
    rng = np.random.default_rng(11)
    k_glob = np.exp(0.5 * rng.normal(size=raw.n_total))
    if case == "hex_box":
        k_glob[:] = 1.0
    bcs = {5: ("dirichlet", 300.0), 6: ("dirichlet", 400.0), 1: ("neumann", 5.0)}

    # ---- the checker: single-partition oracle on the global mesh
    from oracle import port
    conn = dict(zip(("cc_row", "cc_col"), G.connectivity(raw)))
    conn.update(face_cells=raw.face_cells, group_offset=raw.group_offset, group_count=raw.group_count,
                group_id=raw.group_id, group_kind=raw.group_kind)
    g2 = dict(geo)
    g2["ib_type"] = np.full(raw.n_total, -1, np.int32)
    ref = port.thermal_reference(raw, conn, g2, k_glob, bcs, x0=300.0, tol=1e-13)

    if solver_kind == "electric":
        # ElectricModelA (electrostatics + drift / transient charge transport) on this rank's part of a
        # tet mesh, two time steps; checked against the single-partition run of the same code, which
        # tests/test_electric.py pins to the reference's ElectricModel
        import contextlib
        import io
        from fvm_b200 import models as M

        def run(mesh_raw, use_lib):
            mesh = M.Mesh(mesh_raw)
            geomf = M.GeomFields("geom")
            M.MeshMetricsCalculatorA(geomf, [mesh], lib=use_lib).init()
            ef = M.ElectricFields("elec")
            em = M.ElectricModelA(geomf, ef, [mesh], lib=use_lib)
            bcm = em.getBCMap()
            for gid, bc in bcm.items():
                bc.bcType = "SpecifiedPotentialFlux"
                bc["specifiedPotentialFlux"] = 0.0
            if 5 in bcm:
                bcm[5].bcType = "SpecifiedPotential"; bcm[5]["specifiedPotential"] = 0.0
            if 6 in bcm:
                bcm[6].bcType = "SpecifiedPotential"; bcm[6]["specifiedPotential"] = 50.0
            o = em.getOptions()
            o.drift_enable = True
            o["initialTotalCharge"] = 1e10
            o["timeStep"] = 1e-6
            c = em.getConstants()
            c["nTrap"] = 2; c["electron_mobility"] = 1e-4; c["electron_saturation_velocity"] = 1e9
            for nm in ("electrostaticsLinearSolver", "chargetransportLinearSolver"):
                sv = M.AMG()
                sv.relativeTolerance, sv.nMaxIterations, sv.verbosity = 1e-13, 3000, 0
                setattr(o, nm, sv)
            em.init()
            cells = mesh.getCells()
            gids = mesh_raw.cell_global if "cell_global" in mesh_raw else np.arange(mesh_raw.n_total)
            ef.charge[cells][:, 2] = np.where(gids < raw.n_cells, 1e12 * (1 + gids % 5), 0.0)
            ef.chargeN1[cells][:] = ef.charge[cells]
            for _ in range(2):
                with contextlib.redirect_stdout(io.StringIO()):
                    em.advance(1)
                em.updateTime()
            return ef.potential[cells].copy(), ef.charge[cells].copy()

        # single-partition run of the same code BEFORE this process joins the communicator
        lib.comm_destroy()
        pot_ref, chg_ref = run(raw, lib)
        rejoin(lib, world, rank)
        pot, chg = run(loc, lib)
        own = loc.cell_global[:loc.n_cells]
        num = float(((pot[:loc.n_cells] - pot_ref[own]) ** 2).sum()) + float(((chg[:loc.n_cells, 2] / 1e12 - chg_ref[own, 2] / 1e12) ** 2).sum())
        den = float((pot_ref[own] ** 2).sum()) + float(((chg_ref[own, 2] / 1e12) ** 2).sum())
        t = torch.tensor([num, den])
        dist.all_reduce(t)
        out = dict(rank=rank, world=world, n_self=int(loc.n_cells), peers=[int(p) for p in loc.halo["peers"]],
                   err_diag=0.0, err_b=0.0, rel_l2=float(np.sqrt(float(t[0]) / float(t[1]))), ghost_err=0.0,
                   r0=1.0, r=0.0, iters=0, levels=[], collectives=lib.comm_collectives())
        with open(os.path.join(os.environ["FVM_RESULT_DIR"], "rank%d.json" % rank), "w") as fh:
            json.dump(out, fh)
        dist.barrier()
        dist.destroy_process_group()
        return

    if solver_kind == "flow":
        # FlowModelA (SIMPLE: momentum + Rhie-Chow pressure correction) on this rank's part of a lid-driven
        # box, three outer iterations with tight inner solves; checked against the single-partition run of
        # the same code, which tests/test_flow.py pins to the reference's FlowModel
        import contextlib
        import io
        from fvm_b200 import models as M

        def run(mesh_raw, use_lib):
            mesh = M.Mesh(mesh_raw)
            geomf = M.GeomFields("geom")
            M.MeshMetricsCalculatorA(geomf, [mesh], lib=use_lib).init()
            ff = M.FlowFields("flow")
            fm = M.FlowModelA(geomf, ff, [mesh], lib=use_lib)
            bcm = fm.getBCMap()
            for gid, bc in bcm.items():
                bc.bcType = "NoSlipWall"
            if 4 in bcm:
                bcm[4]["specifiedXVelocity"] = 1.0     # the lid (y = top) moves along x
                bcm[4]["specifiedZVelocity"] = 0.3
            o = fm.getOptions()
            for nm in ("momentumLinearSolver", "pressureLinearSolver"):
                sv = M.AMG()
                sv.relativeTolerance, sv.nMaxIterations, sv.verbosity = 1e-13, 3000, 0
                setattr(o, nm, sv)
            vc = fm.getVCMap()[mesh.getID()]
            vc["viscosity"] = 0.05
            fm.init()
            with contextlib.redirect_stdout(io.StringIO()):
                fm.advance(3)
            cells, faces = mesh.getCells(), mesh.getFaces()
            return ff.velocity[cells].copy(), ff.pressure[cells].copy(), ff.massFlux[faces].copy()

        lib.comm_destroy()
        v_ref, p_ref, _ = run(raw, lib)
        rejoin(lib, world, rank)
        v, pr, mf = run(loc, lib)
        own = loc.cell_global[:loc.n_cells]
        num = float(((v[:loc.n_cells] - v_ref[own]) ** 2).sum()) + float(((pr[:loc.n_cells] - p_ref[own]) ** 2).sum())
        den = float((v_ref[own] ** 2).sum()) + float((p_ref[own] ** 2).sum())
        t = torch.tensor([num, den])
        dist.all_reduce(t)
        gi = loc.halo["gather_idx"]
        ghost_err = float(np.abs(v[gi] - v_ref[loc.cell_global[gi]]).max()) if len(gi) else 0.0
        out = dict(rank=rank, world=world, n_self=int(loc.n_cells), peers=[int(p) for p in loc.halo["peers"]],
                   err_diag=0.0, err_b=0.0, rel_l2=float(np.sqrt(float(t[0]) / float(t[1]))), ghost_err=ghost_err,
                   r0=1.0, r=0.0, iters=0, levels=[], collectives=lib.comm_collectives(),
                   vmax=float(np.abs(v_ref[:raw.n_cells]).max()))   # interior cells only
        with open(os.path.join(os.environ["FVM_RESULT_DIR"], "rank%d.json" % rank), "w") as fh:
            json.dump(out, fh)
        dist.barrier()
        dist.destroy_process_group()
        return

    if solver_kind == "species":
        # SpeciesModelA (two species: one diffusing with a source, one convected) with BDF2 on this rank's mesh part,
        # three time steps; checked against the single-partition run of the same code, which tests/test_species.py
        # pins to the reference's SpeciesModel
        import contextlib
        import io
        from fvm_b200 import models as M

        def run(mesh_raw, geo_raw):
            mesh = M.Mesh(mesh_raw)
            geomf = M.GeomFields("geom")
            M.MeshMetricsCalculatorA(geomf, [mesh], lib=lib).init()
            sm = M.SpeciesModelA(geomf, [mesh], 2, lib=lib)
            for m in range(2):
                bcm = sm.getBCMap(m)
                for gid, bc in bcm.items():
                    bc.bcType = "Symmetry"
                if 5 in bcm:
                    bcm[5].bcType = "SpecifiedMassFraction"; bcm[5]["specifiedMassFraction"] = 0.1 + m
                if 6 in bcm:
                    bcm[6].bcType = "SpecifiedMassFraction"; bcm[6]["specifiedMassFraction"] = 0.9 - 0.5 * m
                if 1 in bcm:
                    bcm[1].bcType = "SpecifiedMassFlux"; bcm[1]["specifiedMassFlux"] = 0.2
                sm.getVCMap(m)[mesh.getID()]["massDiffusivity"] = 0.6 + m
                sm.getVCMap(m)[mesh.getID()]["initialMassFraction"] = 0.5
            o = sm.getOptions()
            o.transient, o.timeDiscretizationOrder = True, 2
            o["timeStep"] = 0.05
            sv = M.AMG()
            sv.relativeTolerance, sv.nMaxIterations, sv.verbosity = 1e-13, 3000, 0
            o.linearSolver = sv
            sm.init()
            gids = mesh_raw.cell_global if "cell_global" in mesh_raw else np.arange(mesh_raw.n_total)
            area = np.asarray(geomf.area[mesh.getFaces()])
            sm.getSpeciesFields(1).convectionFlux[mesh.getFaces()][:] = area @ np.array([0.3, -0.2, 0.7])
            sm.getSpeciesFields(0).source[mesh.getCells()][:] = np.where(gids >= 0, 1.0 + (np.maximum(gids, 0) % 5), 0.0)
            for _ in range(3):
                with contextlib.redirect_stdout(io.StringIO()):
                    sm.advance(2)
                sm.updateTime()
            return [np.asarray(sm.getSpeciesFields(m).massFraction[mesh.getCells()]).copy() for m in range(2)]

        lib.comm_destroy()
        ref_x = run(raw, geo)
        rejoin(lib, world, rank)
        x = run(loc, loc.geometry)
        own = loc.cell_global[:loc.n_cells]
        num = sum(float(((x[m][:loc.n_cells] - ref_x[m][own]) ** 2).sum()) for m in range(2))
        den = sum(float((ref_x[m][own] ** 2).sum()) for m in range(2))
        t = torch.tensor([num, den])
        dist.all_reduce(t)
        out = dict(rank=rank, world=world, n_self=int(loc.n_cells), peers=[int(p) for p in loc.halo["peers"]],
                   err_diag=0.0, err_b=0.0, rel_l2=float(np.sqrt(float(t[0]) / float(t[1]))), ghost_err=0.0,
                   r0=1.0, r=0.0, iters=0, levels=[], collectives=lib.comm_collectives())
        with open(os.path.join(os.environ["FVM_RESULT_DIR"], "rank%d.json" % rank), "w") as fh:
            json.dump(out, fh)
        dist.barrier()
        dist.destroy_process_group()
        return

    if solver_kind == "model":
        # the public (reference-mirroring) Python API on this rank's mesh: same script as single rank
        import contextlib
        import io
        from fvm_b200 import models as M
        ref = port.thermal_reference(raw, conn, g2, np.ones(raw.n_total), bcs, x0=300.0, tol=1e-13)
        mesh = M.Mesh(loc)
        geomf = M.GeomFields("geom")
        M.MeshMetricsCalculatorA(geomf, [mesh], lib=lib).init()
        tf = M.ThermalFields("therm")
        tm = M.ThermalModelA(geomf, tf, [mesh], lib=lib)
        bcm = tm.getBCMap()
        for gid, (kind, v) in bcs.items():
            if gid in bcm:
                if kind == "dirichlet":
                    bcm[gid].bcType = "SpecifiedTemperature"; bcm[gid]["specifiedTemperature"] = v
                else:
                    bcm[gid].bcType = "SpecifiedHeatFlux"; bcm[gid]["specifiedHeatFlux"] = v
        sv = M.AMG()
        sv.relativeTolerance, sv.nMaxIterations, sv.verbosity = 1e-13, 3000, 0
        tm.getOptions().linearSolver = sv
        tm.getOptions()["initialTemperature"] = 300.0
        tm.init()
        with contextlib.redirect_stdout(io.StringIO()):
            tm.advance(1)
        x = tf.temperature[mesh.getCells()]
        own = loc.cell_global[:loc.n_cells]
        t = torch.tensor([float(((x[:loc.n_cells] - ref["x"][own]) ** 2).sum()), float((ref["x"][own] ** 2).sum())])
        dist.all_reduce(t)
        gi = loc.halo["gather_idx"]
        out = dict(rank=rank, world=world, n_self=int(loc.n_cells), peers=[int(p) for p in loc.halo["peers"]],
                   err_diag=0.0, err_b=0.0, rel_l2=float(np.sqrt(float(t[0]) / float(t[1]))),
                   ghost_err=float(np.abs(x[gi] - ref["x"][loc.cell_global[gi]]).max()), r0=1.0, r=0.0,
                   iters=int(sv.lastIterations), levels=[], collectives=lib.comm_collectives())
        with open(os.path.join(os.environ["FVM_RESULT_DIR"], "rank%d.json" % rank), "w") as fh:
            json.dump(out, fh)
        dist.barrier()
        dist.destroy_process_group()
        return

    # ---- this rank's part through the library
    row, col = G.connectivity(loc)
    dm = X.DeviceMesh(lib, loc.dim, loc.n_cells, loc.n_total, loc.face_cells, row, col, loc.group_offset,
                      loc.group_count, loc.group_id, loc.group_kind)
    ge = loc.geometry
    dm.set_geometry(ge["face_area"], ge["face_area_mag"], ge["cell_centroid"], ge["cell_volume"],
                    face_centroid=ge["face_centroid"], ib_type=np.full(loc.n_total, -1, np.int32))
    h = loc.halo
    dm.set_halo(h["peers"], h["scatter_off"], h["scatter_idx"], h["gather_off"], h["gather_idx"])
    ds = X.DeviceSystem(lib, dm)
    ds.fill_field(X.FIELD_X, 300.0)
    ds.set_field(X.FIELD_DIFFUSIVITY, k_glob[loc.cell_global])
    kinds = {"dirichlet": X.BC_DIRICHLET, "neumann": X.BC_NEUMANN}
    present = set(int(i) for i in loc.group_id)
    for gid in range(1, 7):
        if gid in present:
            kind, v = bcs.get(gid, ("neumann", 0.0))
            ds.set_bc(gid, kinds[kind], [v])
    # MultiField::sync on its own (fvmgpu_system_halo_exchange): every interface ghost receives its owner's value
    probe = np.where(np.arange(loc.n_total) < loc.n_cells, loc.cell_global.astype(float) + 0.25, -1.0)
    ds.set_field(X.FIELD_DIFFUSIVITY, probe)
    ds.halo_exchange(X.FIELD_DIFFUSIVITY)
    got = ds.get_field(X.FIELD_DIFFUSIVITY)
    sync_err = float(np.abs(got[h["gather_idx"]] - (loc.cell_global[h["gather_idx"]] + 0.25)).max()) if len(h["gather_idx"]) else 0.0
    ds.set_field(X.FIELD_DIFFUSIVITY, k_glob[loc.cell_global])
    ds.assemble()
    a = ds.download()
    own = loc.cell_global[:loc.n_cells]
    scale_d, scale_b = np.abs(ref["diag"]).max(), np.abs(ref["b"]).max()
    err_diag = float(np.abs(a["diag"][:loc.n_cells] - ref["diag"][own]).max() / scale_d)
    err_b = float(np.abs(a["b"][:loc.n_cells] - ref["b"][own]).max() / scale_b)

    o = lib.default_amg_opts()
    o.relativeTolerance, o.nMaxIterations = 1e-13, 3000
    amg = X.DeviceAMG(lib, o)
    if solver_kind == "bcgstab":
        r0, r, it = amg.bcgstab(ds, 300, 1e-13, 1e-50)
    else:
        if solver_kind == "group4":
            o.coarseGroupSize = 4
            amg.set_opts(o)
        elif solver_kind == "jacobi_w":
            o.smootherType, o.cycleType = X.SMOOTHER_JACOBI, X.CYCLE_W
            amg.set_opts(o)
        r0, r, it = amg.solve(ds)
    levels = amg.levels()
    ds.post_solve_update()
    x = ds.get_field(X.FIELD_X)
    num = float(((x[:loc.n_cells] - ref["x"][own]) ** 2).sum())
    den = float((ref["x"][own] ** 2).sum())
    # interface ghosts hold the owner's converged values after updateSolution's sync
    gi = h["gather_idx"]
    ghost_err = float(np.abs(x[gi] - ref["x"][loc.cell_global[gi]]).max()) if len(gi) else 0.0
    t = torch.tensor([num, den])
    dist.all_reduce(t)
    out = dict(rank=rank, world=world, n_self=int(loc.n_cells), peers=[int(p) for p in h["peers"]],
               err_diag=err_diag, err_b=err_b, rel_l2=float(np.sqrt(float(t[0]) / float(t[1]))), ghost_err=ghost_err,
               r0=r0, r=r, iters=it, levels=[int(s) for s in levels["sizes"]], sync_err=sync_err,
               colours=[int(c) for c in levels["colours"]],
               collectives=lib.comm_collectives())
    with open(os.path.join(os.environ["FVM_RESULT_DIR"], "rank%d.json" % rank), "w") as fh:
        json.dump(out, fh)
    amg.close(); ds.close(); dm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
