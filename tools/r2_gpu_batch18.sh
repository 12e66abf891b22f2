#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bgzf.py -x -q 2>&1 | tail -3
SCG_INFLATE_LANES=32 timeout 900 python -m pytest tests/test_gpu_bgzf.py -x -q -k "inflate" 2>&1 | tail -2
timeout 600 python tools/bgzf_bench.py 8000000 6 2>&1 | tail -9
SCG_INFLATE_LANES=32 timeout 600 python tools/bgzf_bench.py 8000000 6 2>&1 | tail -9 | grep -E "256 MiB|block gzip"
