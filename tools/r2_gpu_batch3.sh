#!/bin/bash
# round 2, GPU batch 3: many-files tests, plan tests (flat combo follow-up), full ncu captures of configs 2..5
mkdir -p gpurun_out
python -m pytest tests/test_gpu_many.py tests/test_gpu_plans.py tests/test_gpu_handlers.py -x -q 2>&1 | tail -15 > gpurun_out/r2_many_test.log
cat gpurun_out/r2_many_test.log
for c in 2 3 4 5; do
  python tools/profile_config.py $c 20000000 3 > gpurun_out/r2_prof_c$c.txt 2>&1 || { echo "profile_config $c FAILED"; tail -5 gpurun_out/r2_prof_c$c.txt; continue; }
  cat gpurun_out/r2_prof_c$c.txt
done
ncu --set full --clock-control none --import-source on -k regex:'spec_single_kernel_u' --launch-skip 2 -c 1 -f -o gpurun_out/r2_c2_full python tools/profile_config.py 2 20000000 1 > gpurun_out/r2_c2_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'spec_dual|dual_deferred' --launch-skip 4 -c 2 -f -o gpurun_out/r2_c3_full python tools/profile_config.py 3 20000000 1 > gpurun_out/r2_c3_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'spec_combo|combo_deferred' --launch-skip 4 -c 2 -f -o gpurun_out/r2_c4_full python tools/profile_config.py 4 20000000 1 > gpurun_out/r2_c4_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'spec_random' --launch-skip 2 -c 1 -f -o gpurun_out/r2_c5_full python tools/profile_config.py 5 20000000 1 > gpurun_out/r2_c5_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
