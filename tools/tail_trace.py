#!/usr/bin/env python
"""Per-phase timing of the fused coarse-level V-cycle kernel (k_coop_vcycle): where do its ~0.4 ms go?

  FVMGPU_TAIL_TRACE=1 python tools/tail_trace.py [cells_per_side]

Solves the bench workload with eager launches (profiling mode), then prints the time between consecutive barrier
stamps of the LAST fused-kernel launch, grouped by level of the fused stretch and phase kind."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FVMGPU_TAIL_TRACE"] = "1"
import bench  # noqa: E402
from fvm_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    lib = capi.default_lib()
    raw, mesh, fields, model, solver, geom = bench.build_case(n, lib)
    ls = model._systems[mesh.getID()]
    model._upload(mesh, ls)
    bench.device_step(lib, model, mesh, ls, solver)
    lib.profile_begin()
    st = bench.device_step(lib, model, mesh, ls, solver)
    lib.profile_end(cap=8192)
    t, g = lib.tail_trace()
    sizes = st["levels"]["sizes"]
    cols = st["levels"]["colours"]
    fixed = os.environ.get("FVMGPU_COOP_ROWS")
    first = next(i for i, s in enumerate(sizes) if i > 0 and s <= (int(fixed) if fixed else 75000 + 32000 * min(cols[i], 8)))
    print("levels", sizes, "fused stretch starts at level", first, "stamps", len(t))
    if len(t) < 2:
        return
    dt = np.diff(t.astype(np.int64)) / 1e3
    kinds = {1: "restrict", 2: "residual", 3: "prolong"}
    rows = {}
    for d, tag in zip(dt, g[1:]):
        lvl, k = tag >> 8, tag & 0xff
        name = kinds.get(k, "gs%s c%d" % ("+otf" if k & 0x40 else "", k & 15) if k & 0x10 else ("jacobi" if k & 0x20 else "other"))
        if lvl == 0xff:
            name, lvl = "hand-back barrier", -1
        rows.setdefault((lvl, name), []).append(d)
    total = 0.0
    for (lvl, name), v in sorted(rows.items()):
        nrows = sizes[first + lvl] if 0 <= lvl < len(sizes) - first else 0
        print("  stretch level %3d (%8d rows)  %-18s  n=%d  %.2f us" % (lvl, nrows, name, len(v), float(np.sum(v))))
        total += float(np.sum(v))
    print("total between first and last stamp: %.1f us" % total)


if __name__ == "__main__":
    main()
