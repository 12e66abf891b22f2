#!/bin/bash
# round 2, GPU batch 2: many-files / multi-device tests, launch lists and full ncu captures of configs 3/4/5 (and 2), full gpu suite
mkdir -p gpurun_out
python -m pytest tests/test_gpu_many.py -x -q 2>&1 | tail -15 > gpurun_out/r2_many_test.log
cat gpurun_out/r2_many_test.log
for c in 2 3 4 5; do
  python tools/profile_config.py $c 20000000 3 > gpurun_out/r2_prof_c$c.txt 2>&1 || { echo "profile_config $c FAILED"; tail -5 gpurun_out/r2_prof_c$c.txt; continue; }
  cat gpurun_out/r2_prof_c$c.txt
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_c${c}_launches.csv python tools/profile_config.py $c 20000000 1 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'spec_|deferred|dual_pe_kernel|combo_kernel|random_kernel' --launch-skip 12 -c 8 -f -o gpurun_out/r2_c${c}_full python tools/profile_config.py $c 20000000 1 > gpurun_out/r2_c${c}_ncu.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_gpu_suite.log
cat gpurun_out/r2_gpu_suite.log
