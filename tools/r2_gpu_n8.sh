#!/bin/bash
# round 2, eight GPUs: the headline config, weak scaling, block-gzip vs raw-text end to end
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_n8_config2.json 2> gpurun_out/r2_n8_config2.err || { echo FAILED; tail -20 gpurun_out/r2_n8_config2.err; }
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_n8_config2.json").read())
print("N=8 config 2: %.1f G/s e2e %.1f M/s raw %.1f M/s nccl %s" % (d["value"]/1e9, d["e2e"]["value"]/1e6, d["e2e"]["raw_text"]["value"]/1e6, d.get("nccl_check")))
PY
