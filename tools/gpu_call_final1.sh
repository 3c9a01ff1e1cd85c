#!/bin/bash
# one GPU, last call of the round: smoke, the electric / size tests, the default bench line, ncu traffic of the level-0 kernels
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 200 python bench.py > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err; python - <<PY
import json
try:
    p=json.loads(open("gpurun_out/r2_bench_1gpu_final.json").read().strip().splitlines()[-1])
    print({k:p.get(k) for k in ("value","ms_per_step","amg_cycles","solve_split_ms","roofline","solve_hbm")})
except Exception as e: print("bench failed", e)
PY
timeout 150 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'GsRows|ResidualRows|InjectRows|CorrectRows' -s 18 -c 10 -f -o gpurun_out/r2_prof_level0_col16 python tools/ncu_level0.py 256 1 > gpurun_out/ncu_full2.log 2>&1
echo "capture rc=$?"; tail -2 gpurun_out/ncu_full2.log
timeout 120 python -m pytest tests/test_electric.py tests/test_gpu_sizes.py -m gpu -q > gpurun_out/r2_pytest_gpu_final_subset.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu_final_subset.log
