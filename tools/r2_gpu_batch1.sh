#!/bin/bash
# round 2, GPU batch 1: plan parity tests, table tests, every config's bench at a reduced size, then at the stated size
mkdir -p gpurun_out
python -m pytest tests/test_gpu_plans.py tests/test_gpu_multi.py -x -q 2>&1 | tail -25 > gpurun_out/r2_plans_test.log
cat gpurun_out/r2_plans_test.log
for c in 1 3 4 5; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --reads 20000000 > gpurun_out/r2_small_c$c.json 2> gpurun_out/r2_small_c$c.err || { echo "config $c small FAILED"; tail -15 gpurun_out/r2_small_c$c.err; }
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_small_c$c.json").read())
    print("config $c small:", "%.2f G/s" % (d["value"]/1e9), "frac %.3f" % d["roofline"]["frac"], "e2e %.1f M/s" % (d["e2e"]["value"]/1e6), "cold %.3f s" % d["e2e"]["cold"]["seconds"], d["roofline"]["kernel"][:90])
except Exception as e:
    print("config $c small: no line", e)
PY
done
for c in 1 3 4 5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2_full_c$c.json 2> gpurun_out/r2_full_c$c.err || { echo "config $c full FAILED"; tail -15 gpurun_out/r2_full_c$c.err; }
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_full_c$c.json").read())
    print("config $c full:", "%.2f G/s" % (d["value"]/1e9), "frac %.3f" % d["roofline"]["frac"], "e2e %.1f M/s" % (d["e2e"]["value"]/1e6), "cpu %.2f M/s" % (d["cpu_baseline"]["value"]/1e6), "harvest", d["e2e"]["stages_s"].get("harvest_s"))
except Exception as e:
    print("config $c full: no line", e)
PY
done
SCG_COMBO_FORCE_SPARSE=1 timeout 600 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r2_full_c4_hash.json 2> gpurun_out/r2_full_c4_hash.err; tail -c 600 gpurun_out/r2_full_c4_hash.json
