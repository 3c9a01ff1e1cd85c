#!/bin/bash
# usage: tools/bench2.sh NGPU TAG [ENV=VAL ...] -- runs bench.py on NGPU GPUs with extra environment, prints a summary
N=$1; TAG=$2; shift 2
env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-2} --warmup 1 --no-cpu-baseline $EXTRA > gpurun_out/r2_bench_${N}gpu_$TAG.json 2> gpurun_out/r2_bench_${N}gpu_$TAG.err || tail -5 gpurun_out/r2_bench_${N}gpu_$TAG.err
python - <<PY
import json
try:
    p=json.loads(open("gpurun_out/r2_bench_${N}gpu_$TAG.json").read().strip().splitlines()[-1])
    print("$TAG", {k:p.get(k) for k in ("value","ms_per_step","amg_cycles","solve_split_ms","collectives_per_step","gpu_launches","residual")})
    for t in p.get("kernel_profile",[])[:9]: print("   ",t)
except Exception as e:
    print("$TAG failed", e)
PY
