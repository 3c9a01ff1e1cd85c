#!/bin/bash
# 8 GPUs: coarse-level merge threshold sweep on the thermal workload
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
for MR in 524288 1048576 65536; do
  FVMGPU_MERGE_ROWS=$MR timeout 300 $TR bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu-baseline --no-profile --parity-size 0 > gpurun_out/mr_$MR.json 2> gpurun_out/mr_$MR.err || tail -3 gpurun_out/mr_$MR.err
  python - <<PY
import json
try:
    p=json.loads(open("gpurun_out/mr_$MR.json").read().strip().splitlines()[-1])
    print("merge rows $MR", round(p["ms_per_step"],1), p.get("amg_cycles"), p.get("solve_split_ms"))
except Exception as e: print("merge rows $MR failed", e)
PY
done
