#!/bin/bash
# one GPU: new parity tests, full ncu captures (level-0 solver kernels, assembly kernels), cooperative-kernel row limit on tets / quads
timeout 900 python -m pytest tests/test_gpu_sizes.py tests/test_adaptor_dropin.py tests/test_flow.py -m gpu -q > gpurun_out/r2_pytest_gpu_sizes.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_sizes.log
python tools/ncu_level0.py 256 1 > gpurun_out/plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'GsRows|ResidualRows|InjectRows|CorrectRows' -s 9 -c 10 -f -o gpurun_out/r2_prof_level0 python tools/ncu_level0.py 256 1 > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/r2_prof_level0.ncu-rep
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --parity-size 0"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'AssembleRows|GradientRows' -s 2 -c 2 -f -o gpurun_out/r2_prof_assembly $CMD > gpurun_out/ncu_asm.log 2>&1
echo "assembly capture rc=$?"; ls -la gpurun_out/r2_prof_assembly.ncu-rep
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
for CR in 300000 600000; do run hex_$CR FVMGPU_COOP_ROWS=$CR ""; done
for CR in 75000 150000 300000 600000 1200000; do run tet_$CR FVMGPU_COOP_ROWS=$CR "--mesh tet --size 96"; done
for CR in 150000 300000 1200000; do run etet_$CR FVMGPU_COOP_ROWS=$CR "--workload electric-tet --size 64"; done
for CR in 150000 400000 1200000; do run cav_$CR FVMGPU_COOP_ROWS=$CR "--workload cavity"; done
