#!/bin/bash
# round 2, GPU batch 6: inflate kernel assembling its batches in shared memory
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bgzf.py -x -q 2>&1 | tail -25 > gpurun_out/r2_b6_bgzf_test.log
tail -5 gpurun_out/r2_b6_bgzf_test.log
timeout 600 python tools/bgzf_bench.py 4000000 6 2>&1 | tail -14
SCG_INGEST_CHUNK=134217728 timeout 600 python tools/bgzf_bench.py 4000000 6 2>&1 | tail -2
SCG_INGEST_CHUNK=33554432 timeout 600 python tools/bgzf_bench.py 4000000 6 2>&1 | tail -2
SCG_BGZF_NO_CRC=1 timeout 600 python tools/bgzf_bench.py 4000000 6 2>&1 | tail -10
python tools/profile_config.py 5 200000000 3 2>&1 | tail -1
