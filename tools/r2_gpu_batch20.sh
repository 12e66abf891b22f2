#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_plans.py tests/test_gpu_handlers.py tests/test_gpu_segmented.py tests/test_golden_fixtures.py -x -q 2>&1 | tail -3
python tools/profile_config.py 3 20000000 3 2>&1 | tail -1
SCG_DUAL_NO_FLAT=1 python tools/profile_config.py 3 20000000 3 2>&1 | tail -1
python tools/profile_config.py 3 100000000 3 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_b20_c3_launches.csv python tools/profile_config.py 3 20000000 1 > /dev/null 2>&1
