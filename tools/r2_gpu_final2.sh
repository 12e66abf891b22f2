#!/bin/bash
# round 2, last single-GPU pass: the whole GPU suite, then every config at its stated size
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_final2_tests.log
cat gpurun_out/r2_final2_tests.log
for c in 2 1 3 4 5; do
  timeout 900 python bench.py --config $c > gpurun_out/r2_final2_c$c.json 2> gpurun_out/r2_final2_c$c.err || { echo "config $c FAILED"; tail -8 gpurun_out/r2_final2_c$c.err; }
done
python - <<PY
import json
for c in (1,2,3,4,5):
    try:
        d=json.loads(open("gpurun_out/r2_final2_c%d.json"%c).read().strip().splitlines()[-1])
        print("config %d: %.2f G/s frac %.3f | e2e %.1f M/s raw %.1f M/s cold %.3f s | cpu %.2f M/s" % (c, d["value"]/1e9, d["roofline"]["frac"], d["e2e"]["value"]/1e6, d["e2e"]["raw_text"]["value"]/1e6, d["e2e"]["cold"]["seconds"], d["cpu_baseline"]["value"]/1e6))
    except Exception as e:
        print("config", c, "no line", e)
PY
