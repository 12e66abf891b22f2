#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_bgzf.py -x -q 2>&1 | tail -3
for c in 2 3; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2_b17_bench_c$c.json 2> gpurun_out/r2_b17_bench_c$c.err || tail -5 gpurun_out/r2_b17_bench_c$c.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_b17_bench_c$c.json").read())
print("config $c: %.2f G/s frac %.3f e2e %.1f M/s (%s) raw %.1f M/s" % (d["value"]/1e9, d["roofline"]["frac"], d["e2e"]["value"]/1e6, d["e2e"]["input"][:20], d["e2e"]["raw_text"]["value"]/1e6))
PY
done
