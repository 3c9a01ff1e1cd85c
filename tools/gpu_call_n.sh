#!/bin/bash
# usage: tools/gpu_call_n.sh N  -- scaling point N of the two multi-GPU workloads (+ the 2-GPU tests at N=2)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
declare -A ESIZE=([1]=101 [2]=128 [4]=161 [8]=203)
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/r2_pytest_multigpu_2gpu_final.log 2>&1; tail -3 gpurun_out/r2_pytest_multigpu_2gpu_final.log
fi
if [ "$N" != "8" ]; then
  timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu_final.json 2> gpurun_out/r2_bench_${N}gpu_final.err || tail -5 gpurun_out/r2_bench_${N}gpu_final.err
fi
timeout 1200 $TR bench.py --gpus $N --workload electric-tet --size ${ESIZE[$N]} --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu_electric_tet${ESIZE[$N]}.json 2> gpurun_out/r2_bench_${N}gpu_electric_tet${ESIZE[$N]}.err || tail -5 gpurun_out/r2_bench_${N}gpu_electric_tet${ESIZE[$N]}.err
python - <<PY
import json
for f in ("r2_bench_${N}gpu_final","r2_bench_${N}gpu_electric_tet${ESIZE[$N]}"):
    try:
        p=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, {k:p.get(k) for k in ("value","ms_per_step","amg_cycles","solve_split_ms","phase_ms","parity","e2e")})
    except Exception as e: print(f, "failed", e)
PY
