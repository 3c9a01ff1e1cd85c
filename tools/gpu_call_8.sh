#!/bin/bash
# 8 GPUs: the driver's own command (thermal 512^3, weak scaling) and configs[4] (ElectricModel, 203^3 x 6 tets)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_8gpu_final.json 2> gpurun_out/r2_bench_8gpu_final.err || tail -5 gpurun_out/r2_bench_8gpu_final.err
timeout 1200 $TR bench.py --gpus 8 --workload electric-tet --size 203 --steps 2 --warmup 1 > gpurun_out/r2_bench_8gpu_electric_tet203.json 2> gpurun_out/r2_bench_8gpu_electric_tet203.err || tail -5 gpurun_out/r2_bench_8gpu_electric_tet203.err
python - <<PY
import json
for f in ("r2_bench_8gpu_final","r2_bench_8gpu_electric_tet203"):
    try:
        p=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, {k:p.get(k) for k in ("value","ms_per_step","amg_cycles","solve_split_ms","phase_ms","parity","e2e","cpu_baseline")})
    except Exception as e: print(f, "failed", e)
PY
