#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200 (contract: see the task / DESIGN.md §4).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size CELLS_PER_SIDE]

One STEP = one ThermalModel outer iteration on the synthetic mesh: gradient + fused assembly
(+BCs + boundary elimination) + AMG hierarchy setup + AMG V-cycles to rel-tol 1e-8 + postSolve /
updateSolution -- the statement sequence of ThermalModel::Impl::advance (F/ThermalModel_impl.h:
424-456) with the linear solver converged to the model's own tolerance, i.e. "time to converge".

metric  fp64_cell_updates_per_s = (cells x outer iterations) / time of the timed steps
        (SURVEY §8d(i) extended to the whole path: every cell of the mesh is taken through one full
        assembly + solve + update). Also reported: time_to_converge_s, AMG cycles, per-phase ms,
        solver row-updates/s.
value   device-timed (CUDA events on the library's compute stream), fields resident in HBM.
e2e     the same step through the public API (fvm_b200.models.ThermalModelA.advance) with HOST
        numpy arrays: H2D of temperature / conductivity / source and D2H of temperature + boundary
        heat fluxes inside the timed region (wall clock around the call, device synchronised).
roofline  dominant kernel = the level-0 multicolour Gauss-Seidel pass (GsRows); achieved =
        algorithmic bytes per launch (36 B/row + 12 B/nnz) / mean launch duration from
        CUDA events bracketing every launch of one extra profiled step.
cpu_baseline / --impl reference: the reference's own C++ (oracle/_ref, compiled in place) on the
        box's host cores on a bounded sample of the same workload (smaller mesh, same BCs/solver).

parity  printed with every line, for every N: (1) the converged field of the timed workload against the exact
        discrete solution of this workload (uniform hexes, T fixed on z = 0 / top: linear in z), relative L2 over
        all ranks; (2) a 64^3 jittered-hex problem with random conductivity solved by ALL ranks of this run as one
        partitioned problem (same code path: halo exchange, distributed hierarchy) against the reference's own
        C++ (oracle/_ref, rank 0, single partition): assembled diag / b of every rank's rows (bar 1e-12) and the
        converged temperature, both solvers at rel 1e-13 (bar 1e-8 relative L2).

N>1 (torchrun): ONE problem, z-slab partition, one part per GPU: weak scaling, 256^3 cells per
        GPU (8 GPUs = the 512^3 mesh), value = all cells / max over ranks of the time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fp64_cell_updates_per_s"
UNIT = "cell-updates/s"
REL_TOL = 1e-8
T_HOT, T_COLD, T_INIT = 400.0, 300.0, 300.0


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # NB: no option of this script may be a prefix of a torchrun option ("--n" would be swallowed by
    # `python -m torch.distributed.run ... bench.py --n 128` as an ambiguous --nnodes/--nproc-per-node)
    p.add_argument("--size", dest="n", type=int, default=0,
                   help="cells per side of the per-GPU hex block (default 256: 256^3 cells per GPU)")
    p.add_argument("--workload", choices=("thermal", "cavity", "electric-tet"), default="thermal",
                   help="thermal: the headline (BASELINE configs[1] / [3]); cavity: configs[2], FlowModel SIMPLE on a 2048^2 quad "
                        "mesh; electric-tet: configs[4], ElectricModel on a tet box partitioned over the GPUs (bench_workloads.py)")
    p.add_argument("--global-size", type=int, default=0,
                   help="STRONG scaling: one fixed G^3 hex box (e.g. 512: BASELINE configs[3]) cut into z-slabs over the GPUs "
                        "instead of --size cells per side per GPU; the line then says scaling = strong")
    p.add_argument("--mesh", choices=("hex", "tet"), default="hex",
                   help="tet: the same box cut into 6 jittered tetrahedra per hex (unstructured numbering of the "
                        "coarse levels; single GPU; not the headline workload)")
    p.add_argument("--krylov", action="store_true",
                   help="NOT the headline: solve with the reference's BCGStab preconditioned by one AMG cycle "
                        "(F/BCGStab.cpp) instead of stand-alone AMG cycles; both arms honour it")
    p.add_argument("--ref-n", type=int, default=0,
                   help="cells per side of the CPU sample mesh (default: 128 for --impl reference, 96 for the "
                        "cpu_baseline leg of the GPU arm)")
    p.add_argument("--parity-size", type=int, default=64,
                   help="cells per side of the partitioned oracle-parity problem every run solves (0: skip)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-profile", action="store_true")
    p.add_argument("--_worker", action="store_true", help=argparse.SUPPRESS)
    return p.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 4 + k and r[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- workload
GLOBAL_SIZE = 0


def global_dims(n, world):
    """Weak scaling: n^3 cells per GPU; the box doubles along x, then y, then z with the GPU count
    (8 GPUs at n = 256: the 512^3 mesh of BASELINE.json configs[3]); z-slab partition.
    --global-size G: the box is G^3 whatever the GPU count (strong scaling)."""
    if GLOBAL_SIZE:
        return [GLOBAL_SIZE, GLOBAL_SIZE, GLOBAL_SIZE]
    dims = [n, n, n]
    k, axis = world, 0
    while k > 1:
        if k % 2:
            raise ValueError("--gpus must be a power of two")
        dims[axis % 3] *= 2
        axis += 1
        k //= 2
    return dims


KRYLOV = False
MESH = "hex"
WORKLOAD = "thermal"


def build_case(n, lib, rank=0, world=1):
    """Box of uniform cubic hexes (h = 1/n), k = 1, T = 400 on z = top, T = 300 on z = 0, zero flux
    elsewhere, T0 = 300. world > 1: this rank's z-slab of the global mesh, built
    directly with the reference partitioner's local numbering (fvm_b200.partition.hex_slab)."""
    from fvm_b200 import meshgen as G, models as M, partition as P
    if MESH == "tet":
        if world != 1:
            raise SystemExit("--mesh tet runs on one GPU (use tests/test_multigpu.py for partitioned tets)")
        raw = G.tet_mesh(n, n, n)
    elif world == 1:
        g = GLOBAL_SIZE or n
        raw = G.hex_mesh(g, g, g)
    else:
        nx, ny, nz = global_dims(n, world)
        unit = GLOBAL_SIZE or n
        raw = P.hex_slab(nx, ny, nz, rank, world, nx / unit, ny / unit, nz / unit, lib=lib)
    meshes = [M.Mesh(raw)]
    geom = M.GeomFields("geom")
    M.MeshMetricsCalculatorA(geom, meshes, lib=lib).init()
    fields = M.ThermalFields("therm")
    model = M.ThermalModelA(geom, fields, meshes, lib=lib)
    bc = model.getBCMap()
    if 5 in bc:
        bc[5].bcType = "SpecifiedTemperature"; bc[5]["specifiedTemperature"] = T_COLD
    if 6 in bc:
        bc[6].bcType = "SpecifiedTemperature"; bc[6]["specifiedTemperature"] = T_HOT
    solver = M.AMG()
    solver.relativeTolerance = REL_TOL
    solver.nMaxIterations = 20000
    solver.verbosity = 0
    if KRYLOV:
        top = M.BCGStab()
        top.preconditioner = solver
        top.relativeTolerance, top.nMaxIterations, top.verbosity = REL_TOL, 2000, 0
        solver = top
    model.getOptions().linearSolver = solver
    model.getOptions()["initialTemperature"] = T_INIT
    model.init()
    return raw, meshes[0], fields, model, solver, geom


def device_step(lib, model, mesh, ls, solver):
    """One outer iteration with all fields already resident on the device. Returns phase times."""
    from fvm_b200 import capi
    ls.fill_field(capi.FIELD_X, T_INIT)          # reset the unknown (device-side fill, no host copy)
    lib.timer_start(2)
    lib.timer_start(3)
    model._assemble(ls)
    t_asm = lib.timer_stop(3)
    lib.timer_start(3)
    if KRYLOV:
        dev = solver.preconditioner._device(lib)
        r0, r, it = dev.bcgstab(ls, solver.nMaxIterations, solver.relativeTolerance, solver.absoluteTolerance)
    else:
        dev = solver._device(lib)
        r0, r, it = dev.solve(ls)
    t_solve = lib.timer_stop(3)
    levels = dev.levels()
    timing = dev.last_timing()
    dev.cleanup()
    lib.timer_start(3)
    ls.post_solve_update()
    t_upd = lib.timer_stop(3)
    t_all = lib.timer_stop(2)
    return dict(total_ms=t_all, assemble_ms=t_asm, solve_ms=t_solve, update_ms=t_upd, cycles=it, rnorm0=r0,
                rnorm=r, levels=levels, setup_ms=timing["setup_ms"], cycles_ms=timing["cycles_ms"])


def run_ours(args):
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
        capi.init_comm_from_torch(lib)   # the library's own NCCL communicator (halo exchange, all-reduce)
    n = args.n or 256
    t0 = time.time()
    raw, mesh, fields, model, solver, geom = build_case(n, lib, rank, world)
    ls = model._systems[mesh.getID()]
    setup_s = time.time() - t0
    ncells = raw.n_cells
    model._upload(mesh, ls)

    def barrier():
        lib.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step(lib, model, mesh, ls, solver)
        lib.flush_l2()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0, _, _ = lib.counters()
    c0 = lib.comm_collectives()
    barrier()
    lib.timer_start(0)
    steps = []
    for _ in range(args.steps):
        steps.append(device_step(lib, model, mesh, ls, solver))
    total_ms = lib.timer_stop(0)
    barrier()
    l1, _, _ = lib.counters()
    c1 = lib.comm_collectives()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    # ---- e2e through the public API with host buffers
    cells = mesh.getCells()
    e2e_times = []
    h0 = lib.counters()
    for i in range(max(1, min(args.steps, 2)) + 1):
        fields.temperature[cells][:] = T_INIT
        model._initialNorm = None
        lib.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            model.advance(1)
        lib.synchronize()
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_times.append(dt)
        if i == 0:
            h0 = lib.counters()
    h1 = lib.counters()
    ne2e = len(e2e_times)
    e2e_s = float(np.mean(e2e_times))
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    x = fields.temperature[cells]
    # ---- parity (collective: all ranks)
    zc = np.asarray(geom.coordinate[cells])[:, 2]
    lz = global_dims(n, world)[2] / (GLOBAL_SIZE or n) if MESH == "hex" else 1.0
    parity = parity_block(lib, rank, world, np.asarray(x), zc, lz, ncells, getattr(args, "parity_size", 64))
    # ---- profiled step for the roofline of the dominant kernel
    roof = None
    prof_table = None
    if not args.no_profile:   # every rank takes the step (it contains collectives); rank 0 reports
        lib.profile_begin()
        ps = device_step(lib, model, mesh, ls, solver)
        recs = lib.profile_end(cap=8192)
        roof, prof_table = roofline(recs, ps)
        try:  # full per-(kernel, rows) records for offline analysis (scratch, not part of the contract)
            if rank != 0:
                raise OSError("rank 0 writes")
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "profile_records_n%d.json" % n), "w") as fh:
                json.dump(dict(records=recs, levels=ps["levels"], cycles=ps["cycles"], total_ms=ps["total_ms"]), fh)
        except OSError:
            pass
    # ---- cpu baseline
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_sample(args.ref_n or 96, threads=1, steps=1)   # ~15 s of single-core reference work
    if rank != 0:
        return
    ms_per_step = total_ms / args.steps
    last = steps[-1]
    sizes, nnzs = last["levels"]["sizes"], last["levels"]["nnz"]
    cyc = last["cycles"]
    # row visits by smoother / residual passes in the solve
    # (a sweep with k colour classes visits (2k-1)/k of a level's rows; the level-0 residual (k-1)/k)
    cols_ = [max(int(k), 1) for k in last["levels"]["colours"]]
    row_visits = cyc * (sum(sz * (2 * k - 1) / k for sz, k in zip(sizes, cols_)) + sizes[0] * (cols_[0] - 1) / cols_[0]) + sizes[0]
    mesh_desc = "structured hex mesh" if MESH == "hex" else "box cut into jittered tetrahedra (unstructured tet mesh)"
    out = {
        "metric": METRIC, "value": ncells * world * args.steps / max(total_ms * 1e-3, 1e-12), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong" if GLOBAL_SIZE else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": ("3D steady thermal diffusion, %s " + mesh_desc + " (%d cells, %d per GPU), k=1, "
                                "T=400/300 on z=top/z=0, %sAMG (V-cycle, multicolour GS, nPre 0 / nPost 1, group 2) "
                                "to rel 1e-8, one outer iteration per step")
                               % ("x".join(str(d) for d in global_dims(n, world)), ncells * world, ncells,
                                  "BCGStab preconditioned by one cycle of " if KRYLOV else ""),
                   "cells_per_gpu": ncells, "l2": "inputs (>= 1.8 GB of matrix per pass) exceed the 126 MB L2; "
                                                  "L2 flushed between warm-up steps",
                   "parallelism": ("z-slab domain decomposition, one part per GPU; halo values stored straight into the "
                                   "neighbour's memory over NVLink (device-initiated, one flag per exchange) after every "
                                   "half-sweep, norms all-reduced the same way, coarse levels merged and solved replicated")
                   if world > 1 else "single"},
        "time_to_converge_s": ms_per_step * 1e-3, "step_ms": [round(s_["total_ms"], 3) for s_ in steps],
        "solve_split_ms": {"hierarchy_setup": [round(s_["setup_ms"], 2) for s_ in steps],
                           "cycles": [round(s_["cycles_ms"], 2) for s_ in steps]},
        "amg_cycles": cyc, "amg_levels": len(sizes), "level_sizes": sizes[:6], "level_colours": last["levels"]["colours"][:6],
        "phase_ms": {k: float(np.mean([s[k] for s in steps])) for k in ("assemble_ms", "solve_ms", "update_ms")},
        "residual": [last["rnorm0"], last["rnorm"]],
        "solver_row_updates_per_s": row_visits / max(last["solve_ms"] * 1e-3, 1e-12),
        "assembly_cells_per_s": ncells / max(float(np.mean([s["assemble_ms"] for s in steps])) * 1e-3, 1e-12),
        "gpu_launches": int(l1 - l0),
        "e2e": {"value": ncells * world / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": int((h1[1] - h0[1]) / ne2e), "d2h_bytes_per_step": int((h1[2] - h0[2]) / ne2e),
                "seconds_per_step": e2e_s},
        "solution_check": {"min": float(x.min()), "max": float(x.max()), "mean": float(x[:ncells].mean())},
        "parity": parity,
        "clocks": clocks, "mesh_setup_s": setup_s,
    }
    if world > 1:
        out["collectives_per_step"] = int((c1 - c0) / args.steps)
    if not KRYLOV:
        out["solve_hbm"] = cycle_hbm(last, float(np.mean([s_["cycles_ms"] for s_ in steps])))
    if roof:
        out["roofline"] = roof
        out["kernel_profile"] = prof_table
    if cpu:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))


def parity_block(lib, rank, world, x, zc, lz, n_own, n):
    """The `parity` object of the JSON line (see the module docstring). Collective: every rank calls it."""
    import torch
    import torch.distributed as dist
    from fvm_b200 import capi as X, meshgen as G, partition as P

    def allsum(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return [float(v) for v in t.cpu()]

    out = {}
    # (1) the timed workload against its exact discrete solution (the face-centred two-point flux is exact for a
    #     field that is linear in z on uniform hexes, so the linear profile IS the discrete solution)
    if MESH == "hex":
        exact = T_COLD + (T_HOT - T_COLD) * zc[:n_own] / lz
        num, den = allsum([((x[:n_own] - exact) ** 2).sum(), (exact ** 2).sum()])
        out["workload_vs_exact_solution_rel_l2"] = float(np.sqrt(num / den))
        out["workload_solver_rel_tol"] = REL_TOL
    if n <= 0:
        return out
    # (2) partitioned n^3 problem of this run's ranks against the reference on rank 0
    raw = G.hex_mesh(n, n, n, jitter=0.15, seed=1)
    geo = G.metrics(raw)
    k_glob = np.exp(0.5 * np.random.default_rng(11).normal(size=raw.n_total))
    bcs = {5: ("dirichlet", T_COLD), 6: ("dirichlet", T_HOT), 1: ("neumann", 5.0)}
    ref_x = np.zeros(raw.n_total)
    ref_diag = np.zeros(raw.n_total)
    ref_b = np.zeros(raw.n_total)
    kind = "none"
    ref_cycles = -1
    if rank == 0:
        from oracle import refapi
        if refapi.available():
            kind = "reference (oracle/_ref)"
            rm = refapi.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes,
                                         raw.face_node_count, raw.face_group_size)
            t = refapi.RefThermal(rm)
            t.set_bc(5, "SpecifiedTemperature", specifiedTemperature=T_COLD)
            t.set_bc(6, "SpecifiedTemperature", specifiedTemperature=T_HOT)
            t.set_bc(1, "SpecifiedHeatFlux", specifiedHeatFlux=5.0)
            t.set_solver(refapi.solver_cfg(relativeTolerance=1e-13, nMaxIterations=5000, verbosity=0))
            t.init()
            t.field("conductivity")[:] = k_glob
            a = t.assemble(1)
            t.field("temperature")[:] = T_INIT
            r = t.advance_timed()
            ref_cycles = int(r["cycles"])
            ref_x[:] = t.field("temperature")
            ref_diag[:], ref_b[:] = a["diag"], a["b"]
            t.close()
        else:
            from oracle import port
            kind = "port (oracle/fvm_oracle.c)"
            conn = dict(zip(("cc_row", "cc_col"), G.connectivity(raw)))
            conn.update(face_cells=raw.face_cells, group_offset=raw.group_offset, group_count=raw.group_count,
                        group_id=raw.group_id, group_kind=raw.group_kind)
            g2 = dict(geo)
            g2["ib_type"] = np.full(raw.n_total, -1, np.int32)
            rr = port.thermal_reference(raw, conn, g2, k_glob, bcs, x0=T_INIT, tol=1e-13)
            ref_x[:], ref_diag[:], ref_b[:] = rr["x"], rr["diag"], rr["b"]
    if world > 1:
        pack = torch.from_numpy(np.concatenate([ref_x, ref_diag, ref_b])).cuda()
        dist.broadcast(pack, 0)
        pack = pack.cpu().numpy()
        nt = raw.n_total
        ref_x, ref_diag, ref_b = pack[:nt], pack[nt:2 * nt], pack[2 * nt:]
        loc = P.partition_mesh(raw, geo, P.assign_slabs(raw.n_cells, world), rank)
        cell_global, ge = loc.cell_global, loc.geometry
    else:
        loc = raw
        cell_global, ge = np.arange(raw.n_total), geo
    row, col = G.connectivity(loc)
    dm = X.DeviceMesh(lib, loc.dim, loc.n_cells, loc.n_total, loc.face_cells, row, col, loc.group_offset,
                      loc.group_count, loc.group_id, loc.group_kind)
    dm.set_geometry(ge["face_area"], ge["face_area_mag"], ge["cell_centroid"], ge["cell_volume"],
                    face_centroid=ge["face_centroid"], ib_type=np.full(loc.n_total, -1, np.int32))
    if world > 1:
        h = loc.halo
        dm.set_halo(h["peers"], h["scatter_off"], h["scatter_idx"], h["gather_off"], h["gather_idx"])
    ds = X.DeviceSystem(lib, dm)
    ds.fill_field(X.FIELD_X, T_INIT)
    ds.set_field(X.FIELD_DIFFUSIVITY, k_glob[cell_global])
    kinds = {"dirichlet": X.BC_DIRICHLET, "neumann": X.BC_NEUMANN}
    for gid in sorted(set(int(i) for i in loc.group_id)):
        if 1 <= gid <= 6:
            knd, v = bcs.get(gid, ("neumann", 0.0))
            ds.set_bc(gid, kinds[knd], [v])
    ds.assemble()
    a = ds.download()
    own = cell_global[:loc.n_cells]
    ed = float(np.abs(a["diag"][:loc.n_cells] - ref_diag[own]).max())
    eb = float(np.abs(a["b"][:loc.n_cells] - ref_b[own]).max())
    o = lib.default_amg_opts()
    o.relativeTolerance, o.nMaxIterations = 1e-13, 5000
    amg = X.DeviceAMG(lib, o)
    if KRYLOV:
        r0, r, it = amg.bcgstab(ds, 500, 1e-13, 1e-50)
    else:
        r0, r, it = amg.solve(ds)
    ds.post_solve_update()
    xs = ds.get_field(X.FIELD_X)
    num, den = allsum([((xs[:loc.n_cells] - ref_x[own]) ** 2).sum(), (ref_x[own] ** 2).sum()])
    if world > 1:
        t = torch.tensor([ed, eb], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ed, eb = float(t[0]), float(t[1])
    amg.close(); ds.close(); dm.close()
    sd, sb = float(np.abs(ref_diag).max()), float(np.abs(ref_b).max())
    out["oracle_check"] = {
        "oracle": kind, "mesh": "%d^3 jittered hexes, random conductivity, %d part(s)" % (n, world),
        "assembly_max_rel_diff": {"diag": ed / max(sd, 1e-300), "b": eb / max(sb, 1e-300), "bar": 1e-12},
        "solution_rel_l2": float(np.sqrt(num / max(den, 1e-300))), "solution_bar": 1e-8,
        "solver_rel_tol_both": 1e-13, "cycles": int(it), "reference_cycles": ref_cycles}
    if kind == "none":
        out["oracle_check"] = {"oracle": "unavailable (oracle/_ref and the C port are missing)"}
    else:
        oc = out["oracle_check"]
        oc["pass"] = bool(oc["assembly_max_rel_diff"]["diag"] <= 1e-12 and oc["assembly_max_rel_diff"]["b"] <= 1e-12
                          and oc["solution_rel_l2"] <= 1e-8)
    return out


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json (sustained: inside a long step)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cycle_hbm(step, cycles_ms):
    """Whole-solve HBM fraction of ONE GPU: algorithmic bytes of a V(0,1) cycle with
    multicolour Gauss-Seidel summed over this rank's levels, times the cycles, over the wall time of the
    cycle loop. Row pass = 36 B per row + 12 B per stored off-diagonal entry (10.125 B where the level
    stores 16-bit column offsets: the bytes of the format actually read, so that the fraction stays a utilisation); a sweep with k
    colour classes visits (2k-1)/k of the rows (forward + reverse, repeated class skipped); the level-0
    residual skips the class relaxed last; restriction and prolongation move 32 B and 28 B per fine row."""
    lv = step["levels"]
    sizes, nnzs, cols = lv["sizes"], lv["nnz"], lv["colours"]
    cbytes = lv.get("col_bytes") or [4.0] * len(sizes)     # column-index bytes per entry as stored (4, or 2.125 compressed)
    total = 0.0
    for l, (n, nnz, k) in enumerate(zip(sizes, nnzs, cols)):
        k = max(int(k), 1)
        row_pass = 36.0 * n + (8.0 + cbytes[l]) * nnz
        total += row_pass * (2 * k - 1) / k                    # post-sweep
        if l == 0:
            total += row_pass * (k - 1) / k + 8.0 * n          # residual + 1-norm (last class skipped), r written
        if l + 1 < len(sizes):
            total += 12.0 * n + 20.0 * sizes[l + 1]            # restriction: r, member list | offsets, b, x = 0
            total += 20.0 * n + 8.0 * sizes[l + 1]             # prolongation: ci, x read + write | coarse x
    peak, src = hbm_peak()
    cyc = max(int(step["cycles"]), 1)
    achieved = total * cyc / max(cycles_ms * 1e-3, 1e-12) / 1e9
    return {"bytes_per_cycle": total, "ms_per_cycle": cycles_ms / cyc, "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "peak_source": src,
            "scope": "cycle loop of one GPU (all levels, graph launches, convergence checks); hierarchy build excluded"}


def roofline(recs, step):
    """Dominant kernel: GsRows launches on level 0 (rows = a level-0 colour)."""
    peak, src = hbm_peak()
    sizes, nnzs = step["levels"]["sizes"], step["levels"]["nnz"]
    n0, nnz0 = sizes[0], nnzs[0]
    by = {}
    for r in recs:
        d = by.setdefault(r["name"], dict(launches=0, ms=0.0))
        d["launches"] += r["launches"]
        d["ms"] += r["ms"]
    total = sum(d["ms"] for d in by.values())
    table = sorted(({"kernel": k, "launches": v["launches"], "ms": round(v["ms"], 3),
                     "share": round(v["ms"] / total, 4)} for k, v in by.items()), key=lambda t: -t["ms"])[:12]
    # level-0 Gauss-Seidel launches (the profiler tags every solver launch with its AMG level). One
    # launch = one colour = `rows` rows; algorithmic bytes of a launch (one SpMV-class row
    # pass): 36 B per row + 12 B per off-diagonal entry, the level's nnz apportioned by rows.
    picked = [r for r in recs if r["name"] == "GsRows" and r.get("level", -1) == 0]
    launches = sum(r["launches"] for r in picked)
    ms = sum(r["ms"] for r in picked)
    if not launches:
        return None, table
    col_bytes0 = (step["levels"].get("col_bytes") or [4.0])[0]
    per_row = 36.0 + (8.0 + col_bytes0) * nnz0 / n0
    total_bytes = sum(r["launches"] * r["rows"] * per_row for r in picked)
    achieved = total_bytes / (ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "kernel": "k_rows<GsRows> at AMG level 0 (one colour of the multicolour Gauss-Seidel sweep)",
            "peak_source": src, "bytes_per_launch": total_bytes / launches,
            "bytes_per_row": per_row, "column_index_bytes_per_entry": col_bytes0, "rows_per_launch": sum(r["launches"] * r["rows"] for r in picked) / launches,
            "mean_launch_ms": ms / launches, "launches": launches,
            "share_of_step": ms / step["total_ms"]}
    # DRAM traffic of that kernel from the committed `ncu --set full` capture (per launch), when the
    # capture was taken on this workload size
    tpath = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        g = t.get("GsRows_level0", {})
        if t.get("workload_cells") == n0 and g.get("rows_per_launch") == int(round(roof["rows_per_launch"])):
            roof["traffic"] = g["dram_bytes_read"] + g["dram_bytes_write"]
            roof["traffic_source"] = t["source"]
        elif t.get("workload_cells") == n0 and t.get("note"):
            roof["traffic_note"] = t["note"]
    return roof, table


# ----------------------------------------------------------------------------- CPU arms
def _ref_worker(n, steps):
    """One single-rank run of the reference C++ (oracle/_ref) or, if absent, of the C port."""
    from fvm_b200 import meshgen as G
    from oracle import refapi
    if WORKLOAD == "cavity":
        import bench_workloads as W
        dt, ref = W._cavity_reference(n, 0.01, max(steps, 1), tight=False)
        return dict(kind="reference", cells=ref["n_cells"], runs=[dict(seconds=dt, cycles=-1)] * max(steps, 1))
    if WORKLOAD == "electric-tet":
        import bench_workloads as W
        rraw = G.tet_mesh(n, n, n, lx=W.E_BOX, ly=W.E_BOX, lz=W.E_BOX)
        dt, _ = W._electric_reference(rraw, max(steps, 1), 1e-8, iters=100, kind=1)
        return dict(kind="reference", cells=rraw.n_cells, runs=[dict(seconds=dt, cycles=-1)] * max(steps, 1))
    raw = G.hex_mesh(n, n, n)
    out = []
    if refapi.available():
        rm = refapi.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes,
                                     raw.face_node_count, raw.face_group_size)
        kind = "reference"
        for _ in range(steps):
            t = refapi.RefThermal(rm)
            t.set_bc(5, "SpecifiedTemperature", specifiedTemperature=T_COLD)
            t.set_bc(6, "SpecifiedTemperature", specifiedTemperature=T_HOT)
            t.set_solver(refapi.solver_cfg(kind=1 if KRYLOV else 0, relativeTolerance=REL_TOL,
                                           nMaxIterations=2000 if KRYLOV else 20000, verbosity=0))
            t.set_option("initialTemperature", T_INIT)
            t.init()
            r = t.advance_timed()
            out.append(dict(seconds=r["assemble_s"] + r["solve_s"] + r["update_s"], cycles=r["cycles"],
                            assemble_s=r["assemble_s"], solve_s=r["solve_s"]))
            t.close()
    else:
        from oracle import port
        kind = "port"
        conn = dict(zip(("cc_row", "cc_col"), G.connectivity(raw)))
        conn.update(face_cells=raw.face_cells, group_offset=raw.group_offset, group_count=raw.group_count,
                    group_id=raw.group_id, group_kind=raw.group_kind)
        geo = G.metrics(raw)
        for _ in range(steps):
            t0 = time.perf_counter()
            port.thermal_reference(raw, conn, geo, np.ones(raw.n_total), {5: ("dirichlet", T_COLD), 6: ("dirichlet", T_HOT)},
                                   x0=T_INIT, tol=REL_TOL)
            out.append(dict(seconds=time.perf_counter() - t0, cycles=-1))
    return dict(kind=kind, cells=raw.n_cells, runs=out)


def cpu_sample(n, threads, steps):
    """`threads` independent single-rank processes of the reference, each on its own n^3 mesh."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--_worker", "--ref-n", str(n),
           "--steps", str(steps), "--workload", WORKLOAD] + (["--krylov"] if KRYLOV else [])
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    t0 = time.perf_counter()
    procs = [subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True, env=env) for _ in range(threads)]
    res = []
    for p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("reference worker failed")
        res.append(json.loads(out.strip().splitlines()[-1]))
    wall = time.perf_counter() - t0
    cells = res[0]["cells"]
    per_proc = [sum(r["seconds"] for r in rr["runs"]) for rr in res]
    value = threads * cells * steps / max(per_proc)
    return {"value": value, "unit": UNIT, "cores": threads, "kind": res[0]["kind"],
            "sample": "%d^3 hex mesh (%d cells), same BCs / solver / tolerance as the GPU workload, %d step(s) per "
                      "process, %d independent single-rank process(es) (no MPI runtime in the image: throughput "
                      "proxy without halo cost)" % (n, cells, steps, threads),
            "seconds_per_step": max(per_proc) / steps, "cycles": res[0]["runs"][-1]["cycles"],
            "phase_s": {k: res[0]["runs"][-1].get(k) for k in ("assemble_s", "solve_s")}, "wall_s": wall}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample: the reference needs O(n) more cycles per doubling of the mesh (165 at 64^3, ~330 at
    # 128^3, ~700 at 256^3), so the sample is taken as large as a few minutes allow (128^3, <= 2 steps)
    n = args.ref_n or {"thermal": 128, "cavity": 512, "electric-tet": 32}[WORKLOAD]
    for _ in range(min(args.warmup, 1)):
        cpu_sample(min(n, 32), threads, 1)
    steps = max(1, min(args.steps, 2))
    cpu = cpu_sample(n, threads, steps)
    out = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
           "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": cpu["seconds_per_step"] * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": {"thermal": "3D steady thermal diffusion on a structured hex mesh, k=1, T=400/300 on z=1/z=0, "
                                              "AMG (V-cycle, GS, nPre 0 / nPost 1, group 2) to rel 1e-8, one outer iteration per step",
                                   "cavity": "lid-driven cavity, FlowModel SIMPLE with the reference's default solvers, one SIMPLE "
                                             "iteration per step",
                                   "electric-tet": "ElectricModel electrostatics + drift / transient charge transport on jittered "
                                                   "tets, BCGStab + AMG to rel 1e-8, one time step per step"}[WORKLOAD]
                                  + "; CPU arm: bounded sample " + cpu["sample"]},
           "cpu_baseline": cpu, "gpu_launches": 0,
           "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    KRYLOV = bool(a.krylov)
    MESH = a.mesh
    WORKLOAD = a.workload
    GLOBAL_SIZE = a.global_size
    if a._worker:
        print(json.dumps(_ref_worker(a.ref_n, a.steps)))
    elif a.impl == "reference":
        run_reference(a)
    elif a.workload == "thermal":
        run_ours(a)
    else:
        # the workload module imports THIS module by name: make `bench` resolve to the running script
        sys.modules.setdefault("bench", sys.modules["__main__"])
        import bench_workloads
        {"cavity": bench_workloads.run_cavity, "electric-tet": bench_workloads.run_electric}[a.workload](a)
