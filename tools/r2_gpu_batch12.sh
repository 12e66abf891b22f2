#!/bin/bash
mkdir -p gpurun_out
SCG_RANDOM_PARTS=4 timeout 900 python -m pytest tests/test_gpu_plans.py tests/test_gpu_handlers.py -x -q -k "random" 2>&1 | tail -2
python tools/profile_config.py 5 200000000 3 2>&1 | tail -1
SCG_RANDOM_PARTS=8 python tools/profile_config.py 5 200000000 3 2>&1 | tail -1
SCG_RANDOM_PARTS=0 python tools/profile_config.py 5 200000000 3 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_b12_c5_launches.csv python tools/profile_config.py 5 200000000 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:'random_count_parts' --launch-skip 8 -c 1 -f -o gpurun_out/r2_b12_parts_full python tools/profile_config.py 5 200000000 1 > /dev/null 2>&1
