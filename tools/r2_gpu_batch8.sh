#!/bin/bash
mkdir -p gpurun_out
python tools/inflate_only.py 256 6 2
ncu --set full --clock-control none --import-source on -k regex:'inflate_kernel' --launch-skip 1 -c 1 -f -o gpurun_out/r2_inflate_full python tools/inflate_only.py 256 6 2 > gpurun_out/r2_inflate_ncu.log 2>&1
ls -la gpurun_out/r2_inflate_full.ncu-rep
