#!/bin/bash
# block-gzip pipeline knobs: chunk size x slots in flight
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_many.py -x -q 2>&1 | tail -2
for cfg in "67108864 8" "134217728 6" "134217728 8" "201326592 5" "268435456 4" "268435456 6"; do
  set -- $cfg
  echo "chunk $1 slots $2: $(SCG_INGEST_CHUNK=$1 SCG_BGZF_SLOTS=$2 timeout 600 python tools/bgzf_bench.py 8000000 6 2>&1 | tail -1 | cut -c1-60)"
done
