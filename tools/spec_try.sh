#!/bin/bash
# Compiles the specialised single-barcode kernel for the BASELINE configs[1] template with nvcc (as NVRTC would at
# run time) and prints registers / spills / opcode histogram.  usage: tools/spec_try.sh [-DSPEC_MIN_BLOCKS=5 ...]
cd "$(dirname "$0")/../screencounter_b200/csrc"
echo '#include "spec_single.cuh"' > /tmp/spec_try.cu
nvcc -I. -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xptxas -v -cubin -DSPEC_CUSTOM -DSPEC_T=44 \
  '-DSPEC_FBASES="CAGCTACGTACG--------------------CCAGCTCGATCG"' '-DSPEC_RBASES="CGATCGAGCTGG--------------------CGTACGTAGCTG"' \
  -DSPEC_FWD=1 -DSPEC_REV=1 -DSPEC_W=3 -DSPEC_NB=1 -DSPEC_CB=1 -DSPEC_MM=1 -DSPEC_MAXMM=1 -DSPEC_USE_FIRST=1 -DSPEC_FSTART=12 -DSPEC_RSTART=12 \
  -DSPEC_KEYLEN=20 -DSPEC_NAME=spec_single_kernel "$@" -o /tmp/spec_try.cubin /tmp/spec_try.cu 2>&1 | grep -v "^$" | grep -B1 -A3 "error\|entry function" | tail -14
python ../../tools/sass_hist.py /tmp/spec_try.cubin
