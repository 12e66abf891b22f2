#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bgzf.py -x -q 2>&1 | tail -3
timeout 600 python tools/bgzf_bench.py 8000000 6 2>&1 | tail -9
SCG_INGEST_CHUNK=268435456 SCG_BGZF_SLOTS=4 timeout 600 python tools/bgzf_bench.py 8000000 6 2>&1 | tail -1
timeout 900 python bench.py --config 2 --steps 5 --warmup 3 > gpurun_out/r2_b9_bench_c2.json 2> gpurun_out/r2_b9_bench_c2.err || tail -5 gpurun_out/r2_b9_bench_c2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_b9_bench_c2.json").read())
print("config 2: %.2f G/s frac %.3f e2e %.1f M/s bgzf %.1f M/s" % (d["value"]/1e9, d["roofline"]["frac"], d["e2e"]["value"]/1e6, d["e2e"]["block_gzip"]["value"]/1e6))
PY
