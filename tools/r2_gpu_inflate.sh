#!/bin/bash
# the device inflater: throughput with 32 / 16 / 8 lanes per member on 1 GiB of text (16448 members), then the end-to-end figure
mkdir -p gpurun_out
for l in 32 16 8; do
  SCG_INFLATE_LANES=$l python tools/inflate_only.py 1024 6 3 > gpurun_out/infl_big_lanes$l.log 2>&1
done
python bench.py --steps 3 --warmup 3 > gpurun_out/infl_bench32.json 2> gpurun_out/infl_bench32.err
tail -n 3 gpurun_out/infl_big_lanes*.log
python - <<'P'
import json
d=json.loads(open('gpurun_out/infl_bench32.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e']['raw_text']['value'], d['e2e']['stages_s'])
P
