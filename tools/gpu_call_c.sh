#!/bin/bash
# one GPU: batched fused kernels (FVMGPU_TAIL_BATCH) x cooperative row limit, conditional-graph probe
timeout 900 python -m pytest tests/test_gpu_sizes.py tests/test_solver.py tests/test_flow.py -m gpu -q -x > gpurun_out/r2_pytest_gpu_c.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_c.log
timeout 120 tools/cond_graph_probe > gpurun_out/r2_cond_graph_probe.log 2>&1; cat gpurun_out/r2_cond_graph_probe.log
run() {  # name, env, args
  env $2 python bench.py $3 --steps 3 --warmup 2 --no-cpu-baseline --no-profile --parity-size 0 > gpurun_out/cr_$1.json 2>gpurun_out/cr_$1.err
  python - <<PY
import json
try:
    p=json.loads(open("gpurun_out/cr_$1.json").read().strip().splitlines()[-1])
    print("$1", "$2", round(p["ms_per_step"],2), p.get("solve_split_ms"), p.get("phase_ms"), p.get("amg_cycles"))
except Exception as e: print("$1 failed", e)
PY
}
for U in 1 2 4; do for CR in 150000 300000 1200000; do run hex_u${U}_$CR "FVMGPU_TAIL_BATCH=$U FVMGPU_COOP_ROWS=$CR" ""; done; done
for U in 1 2 4; do for CR in 300000 1200000; do run etet_u${U}_$CR "FVMGPU_TAIL_BATCH=$U FVMGPU_COOP_ROWS=$CR" "--workload electric-tet --size 64"; done; done
for U in 1 2 4; do run tet_u${U} "FVMGPU_TAIL_BATCH=$U FVMGPU_COOP_ROWS=300000" "--mesh tet --size 96"; done
for U in 1 2 4; do run cav_u${U} "FVMGPU_TAIL_BATCH=$U" "--workload cavity"; done
