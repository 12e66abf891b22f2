#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_bgzf.py tests/test_gpu_single.py tests/test_gpu_handlers.py -x -q 2>&1 | tail -3
timeout 600 python tools/bgzf_bench.py 8000000 6 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_b16_e2e_launches.csv python tools/bgzf_bench.py 1000000 6 > /dev/null 2>&1
