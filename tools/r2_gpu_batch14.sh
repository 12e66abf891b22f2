#!/bin/bash
mkdir -p gpurun_out/cubin_pred0 gpurun_out/cubin_pred1
SCG_SPEC_PRED=1 timeout 900 python -m pytest tests/test_gpu_single.py tests/test_gpu_resident.py -x -q 2>&1 | tail -2
for k in 1 2; do
  SCG_SPEC_PRED=0 SCG_DUMP_CUBIN=gpurun_out/cubin_pred0 python tools/profile_config.py 2 100000000 5 2>&1 | tail -1
  SCG_SPEC_PRED=1 SCG_DUMP_CUBIN=gpurun_out/cubin_pred1 python tools/profile_config.py 2 100000000 5 2>&1 | tail -1
done
ls gpurun_out/cubin_pred1
