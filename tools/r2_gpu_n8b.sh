#!/bin/bash
# round 2, eight GPUs: the literal BASELINE sizes split over the GPUs (strong scaling) for configs 2 and 5
mkdir -p gpurun_out
for args in "--config 5 --scaling strong" "--config 2 --scaling strong"; do
  tag=$(echo $args | tr -d ' -')
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 $args > gpurun_out/r2_n8_$tag.json 2> gpurun_out/r2_n8_$tag.err || { echo "bench $args FAILED"; tail -12 gpurun_out/r2_n8_$tag.err; }
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_n8_$tag.json").read())
    print("$args:", "%.2f G/s" % (d["value"]/1e9), d["scaling"], "ms/step %.3f" % d["ms_per_step"], "e2e %.1f M/s" % (d["e2e"]["value"]/1e6), "nccl_check", d.get("nccl_check"), d["config"]["parallelism"][:200])
except Exception as e:
    print("$args: no line", e)
PY
done
