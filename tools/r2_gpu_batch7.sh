#!/bin/bash
# round 2, GPU batch 7: inflate with 16-bit tables (36 warps/SM), block-gzip chunks issued 7 ahead on 4 streams
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bgzf.py tests/test_gpu_ingest.py -x -q 2>&1 | tail -6
timeout 600 python tools/bgzf_bench.py 8000000 6 2>&1 | tail -12
SCG_INGEST_CHUNK=33554432 SCG_BGZF_SLOTS=12 timeout 600 python tools/bgzf_bench.py 8000000 6 2>&1 | tail -1
SCG_INGEST_CHUNK=134217728 SCG_BGZF_SLOTS=6 timeout 600 python tools/bgzf_bench.py 8000000 6 2>&1 | tail -1
