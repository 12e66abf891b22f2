#!/bin/bash
# round 2, GPU batch 5: block-gzip on the device (tests + throughput), random kernel with quarter-full table and 4-way insert follow-up
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bgzf.py -x -q 2>&1 | tail -25 > gpurun_out/r2_b5_bgzf_test.log
cat gpurun_out/r2_b5_bgzf_test.log
timeout 600 python tools/bgzf_bench.py 4000000 6 2>&1 | tail -14
timeout 900 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_plans.py tests/test_gpu_handlers.py -x -q 2>&1 | tail -8
for c in 5; do
  python tools/profile_config.py $c 20000000 3 2>&1 | tail -1
done
python tools/profile_config.py 5 200000000 3 2>&1 | tail -1
SCG_L2_FETCH=64 python tools/profile_config.py 5 200000000 3 2>&1 | tail -1
python tools/profile_config.py 2 50000000 3 2>&1 | tail -1
SCG_L2_FETCH=64 python tools/profile_config.py 2 50000000 3 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_b5_c5_launches.csv python tools/profile_config.py 5 200000000 1 > /dev/null 2>&1
