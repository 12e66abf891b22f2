#!/bin/bash
# round 2, two GPUs: NCCL merge test, multi-device C-ABI tests, bench at N=2 for configs 2 (weak + strong) and 5
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_many.py -x -q 2>&1 | tail -8
for args in "--config 2" "--config 2 --scaling strong" "--config 5 --reads 100000000" "--config 4 --reads 50000000"; do
  tag=$(echo $args | tr -d ' -')
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 $args > gpurun_out/r2_n2_$tag.json 2> gpurun_out/r2_n2_$tag.err || { echo "bench $args FAILED"; tail -12 gpurun_out/r2_n2_$tag.err; }
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_n2_$tag.json").read())
    print("$args:", "%.2f G/s" % (d["value"]/1e9), d["scaling"], "e2e %.1f M/s" % (d["e2e"]["value"]/1e6), "bgzf", d["e2e"].get("block_gzip") and "%.1f M/s" % (d["e2e"]["block_gzip"]["value"]/1e6), "nccl_check", d.get("nccl_check"))
except Exception as e:
    print("$args: no line", e)
PY
done
