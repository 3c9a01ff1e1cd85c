"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python tools/summarize_launches.py launches.csv "<command that was profiled>" > profiles/<name>.txt"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    m = re.search(r"k_rows<fvmgpu::(\w+)", name) or re.search(r"k_reduce1<\d+, *fvmgpu::(\w+)", name)
    if m:
        return m.group(1)
    return re.sub(r"\(.*", "", name)[:60]


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((short(r["Kernel Name"]), v * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for k, us in rows:
        agg[k][0] += 1
        agg[k][1] += us
    total = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none: %s" % cmd)
    print("# cold-cache, serialised per-launch times: compare SHARES with bench.py's kernel_profile, not absolutes")
    print("# %d launches, %.2f ms total" % (len(rows), total / 1e3))
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%-44s launches %6d  total %10.1f us  share %.4f  mean %.2f us" % (k, n, us, us / total, us / n))


if __name__ == "__main__":
    main()
