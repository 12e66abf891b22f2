#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bgzf.py -x -q 2>&1 | tail -3
timeout 600 python tools/bgzf_bench.py 8000000 6 2>&1 | tail -9
timeout 600 python tools/bgzf_bench.py 1000000 6 2>&1 | tail -2
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
