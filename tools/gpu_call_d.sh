#!/bin/bash
# one GPU: whole GPU suite with the 16-bit column copy, A/B of FVMGPU_COL16 on the workloads
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu_d.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_d.log
run() {  # name, env, args
  env $2 python bench.py $3 --steps 3 --warmup 2 --no-cpu-baseline --no-profile --parity-size 0 > gpurun_out/c16_$1.json 2>gpurun_out/c16_$1.err
  python - <<PY
import json
try:
    p=json.loads(open("gpurun_out/c16_$1.json").read().strip().splitlines()[-1])
    print("$1", "$2", round(p["ms_per_step"],2), p.get("solve_split_ms"), p.get("phase_ms"), p.get("amg_cycles"), (p.get("solve_hbm") or {}).get("frac"))
except Exception as e: print("$1 failed", e)
PY
}
run hex_on "FVMGPU_COL16=1 FVMGPU_COL16_REPORT=1" ""
grep fvmgpu gpurun_out/c16_hex_on.err | sort | uniq -c | head -30
run hex_off "FVMGPU_COL16=0" ""
run hex_on2 "FVMGPU_COL16=1" ""
run krylov_on "FVMGPU_COL16=1" "--krylov"
run krylov_off "FVMGPU_COL16=0" "--krylov"
run tet_on "FVMGPU_COL16=1 FVMGPU_COL16_REPORT=1" "--mesh tet --size 96"
grep fvmgpu gpurun_out/c16_tet_on.err | sort | uniq -c | head -30
run tet_off "FVMGPU_COL16=0" "--mesh tet --size 96"
run etet_on "FVMGPU_COL16=1" "--workload electric-tet --size 64"
run etet_off "FVMGPU_COL16=0" "--workload electric-tet --size 64"
run cav_on "FVMGPU_COL16=1" "--workload cavity"
run cav_off "FVMGPU_COL16=0" "--workload cavity"
