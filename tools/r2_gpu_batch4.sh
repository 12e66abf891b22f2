#!/bin/bash
# round 2, GPU batch 4: random kernel with deferred inserts, pipelined combo kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_plans.py tests/test_gpu_handlers.py tests/test_gpu_multi.py tests/test_gpu_many.py tests/test_gpu_properties.py -x -q 2>&1 | tail -15 > gpurun_out/r2_b4_test.log
cat gpurun_out/r2_b4_test.log
for c in 3 4 5; do
  python tools/profile_config.py $c 20000000 3 2>&1 | tail -1
done
python tools/profile_config.py 5 200000000 3 2>&1 | tail -1
python tools/profile_config.py 4 100000000 3 2>&1 | tail -1
for c in 4 5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_b4_c${c}_launches.csv python tools/profile_config.py $c 20000000 1 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:'spec_combo' --launch-skip 2 -c 1 -f -o gpurun_out/r2_b4_c4_full python tools/profile_config.py 4 20000000 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'spec_random|random_insert' --launch-skip 4 -c 2 -f -o gpurun_out/r2_b4_c5_full python tools/profile_config.py 5 20000000 1 > /dev/null 2>&1
ls -la gpurun_out/r2_b4*.ncu-rep
