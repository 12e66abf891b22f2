#!/bin/bash
# round 2, GPU batch 10: random barcodes counted part by part
mkdir -p gpurun_out
SCG_RANDOM_PARTS=4 timeout 900 python -m pytest tests/test_gpu_plans.py tests/test_gpu_handlers.py tests/test_gpu_multi.py -x -q -k "random" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_plans.py tests/test_gpu_handlers.py tests/test_gpu_ingest.py -x -q -k "random" 2>&1 | tail -4
python tools/profile_config.py 5 20000000 3 2>&1 | tail -1
python tools/profile_config.py 5 200000000 3 2>&1 | tail -1
SCG_RANDOM_PARTS=16 python tools/profile_config.py 5 200000000 3 2>&1 | tail -1
SCG_RANDOM_PARTS=0 python tools/profile_config.py 5 200000000 3 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_b10_c5_launches.csv python tools/profile_config.py 5 200000000 1 > /dev/null 2>&1
timeout 900 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r2_b10_bench_c5.json 2> gpurun_out/r2_b10_bench_c5.err || tail -5 gpurun_out/r2_b10_bench_c5.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_b10_bench_c5.json").read())
print("config 5 bench: %.2f G/s frac %.3f e2e %.1f M/s bgzf %.1f M/s" % (d["value"]/1e9, d["roofline"]["frac"], d["e2e"]["value"]/1e6, d["e2e"]["block_gzip"]["value"]/1e6))
PY
