#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bgzf.py -x -q 2>&1 | tail -2
SCG_RANDOM_PARTS=4 timeout 900 python -m pytest tests/test_gpu_plans.py tests/test_gpu_handlers.py tests/test_gpu_multi.py -x -q -k "random" 2>&1 | tail -2
timeout 600 python tools/bgzf_bench.py 8000000 6 2>&1 | tail -9
python tools/profile_config.py 5 200000000 3 2>&1 | tail -1
