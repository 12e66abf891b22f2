#!/bin/bash
# Sweeps the tuning knobs of the specialised single-barcode kernel (device-timed value only).
# usage: tools/sweep_single.sh "<min_blocks list>" "<stages list>" [reads]
reads=${3:-50000000}
for mb in $1; do for st in $2; do
  SCG_SPEC_MIN_BLOCKS=$mb SCG_SPEC_STAGES=$st python bench.py --steps 5 --warmup 3 --reads $reads --e2e-reads 500000 --cpu-reads 100000 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('min_blocks=$mb stages=$st  %.2f G reads/s  frac %.3f  kernel_ms %.3f' % (d['value']/1e9, d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']))"
done; done
