"""Debug aid: run (part of) the GPU test-suite in THIS process, then solve mm226 twice and report where the
two runs start to differ (hierarchy, colours, residual history, solution)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import pytest  # noqa: E402

from fvm_b200 import capi as X  # noqa: E402

files = sys.argv[1:] or ["tests"]
rc = pytest.main(["-q", "-m", "gpu", "-k", "not deterministic and not fused", "-p", "no:cacheprovider"] + files)
print("suite rc", rc, flush=True)
lib = X.default_lib()
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "mm226.npz")))
runs = []
for k in range(4):
    ds = X.DeviceSystem(lib, raw=(int(g["n"]), 0, g["row"], g["col"], g["diag"], g["off"], g["b"]))
    amg = X.DeviceAMG(lib)
    r0, r, it = amg.solve(ds)
    runs.append(dict(x=ds.get_field(X.FIELD_DELTA), lv=amg.levels(), h=amg.history(), it=it))
    amg.close(); ds.close()
for k in range(1, 4):
    a, b = runs[0], runs[k]
    same_h = [i for i, (u, v) in enumerate(zip(a["h"], b["h"])) if u != v]
    print("run", k, "x equal", np.array_equal(a["x"], b["x"]), "levels equal", a["lv"]["sizes"] == b["lv"]["sizes"],
          "colours equal", a["lv"]["colours"] == b["lv"]["colours"], "iters", a["it"], b["it"],
          "first differing history index", same_h[:1], "max|dx|", float(np.abs(a["x"] - b["x"]).max()))
    if a["lv"]["sizes"] != b["lv"]["sizes"] or a["lv"]["colours"] != b["lv"]["colours"]:
        print("   ", a["lv"]["sizes"], a["lv"]["colours"]); print("   ", b["lv"]["sizes"], b["lv"]["colours"])
