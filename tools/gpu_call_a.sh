#!/bin/bash
# one GPU: tests, launch list, full captures of the level-0 kernels, single-CTA threshold experiment
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_1gpu.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_1gpu.log
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --parity-size 0"
$CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2500 --csv --log-file gpurun_out/r2_launches_256cubed.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2_launches_256cubed.csv)"
python tools/ncu_level0.py 256 1 > gpurun_out/plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'GsRows|ResidualRows|InjectRows|CorrectRows' -s 6 -c 8 -f -o gpurun_out/r2_prof_level0 python tools/ncu_level0.py 256 1 > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/r2_prof_level0.ncu-rep
for TR in 4096 16384 65536; do
  FVMGPU_TAIL_ROWS=$TR python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-profile --parity-size 0 > gpurun_out/tr_hex_$TR.json 2>/dev/null
  FVMGPU_TAIL_ROWS=$TR python bench.py --workload electric-tet --size 64 --steps 3 --warmup 2 --no-cpu-baseline --no-profile --parity-size 0 > gpurun_out/tr_tet_$TR.json 2>/dev/null
  python - <<PY
import json
for f in ("hex","tet"):
    try:
        p=json.loads(open("gpurun_out/tr_%s_$TR.json"%f).read().strip().splitlines()[-1])
        print("tail rows $TR", f, round(p["ms_per_step"],2), p.get("solve_split_ms"), p.get("phase_ms"))
    except Exception as e: print("tail rows $TR", f, "failed", e)
PY
done
