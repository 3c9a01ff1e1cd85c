#!/bin/bash
# one GPU: whole GPU suite, then the workloads with the default settings and with FVMGPU_COL16=0
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_e.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_e.log
run() {  # name, env, args
  env $2 python bench.py $3 --steps 3 --warmup 2 --no-cpu-baseline --no-profile --parity-size 0 > gpurun_out/c16b_$1.json 2>gpurun_out/c16b_$1.err
  python - <<PY
import json
try:
    p=json.loads(open("gpurun_out/c16b_$1.json").read().strip().splitlines()[-1])
    print("$1", "$2", round(p["ms_per_step"],2), p.get("solve_split_ms"), p.get("phase_ms"), p.get("amg_cycles"), (p.get("solve_hbm") or {}).get("frac"))
except Exception as e: print("$1 failed", e)
PY
}
run hex_on "FVMGPU_COL16=1" ""
run hex_off "FVMGPU_COL16=0" ""
run tet_on "FVMGPU_COL16=1" "--mesh tet --size 96"
run tet_off "FVMGPU_COL16=0" "--mesh tet --size 96"
run etet_on "FVMGPU_COL16=1" "--workload electric-tet --size 64"
run etet_off "FVMGPU_COL16=0" "--workload electric-tet --size 64"
run cav_on "FVMGPU_COL16=1" "--workload cavity"
run cav_off "FVMGPU_COL16=0" "--workload cavity"
