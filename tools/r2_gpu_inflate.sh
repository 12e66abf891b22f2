#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bgzf.py -x -q 2>&1 | tail -4
timeout 300 python tools/inflate_only.py 256 6 3 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/infl_launches.csv python tools/inflate_only.py 256 6 2 > /dev/null 2>&1
grep -i "inflate_kernel\|crc_kernel" gpurun_out/infl_launches.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -4
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/sp_$name.json 2> gpurun_out/sp_$name.err
  python - "$name" <<'P'
import json,sys
try:
    d=json.loads(open("gpurun_out/sp_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "e2e %.1f M/s"%(d["e2e"]["value"]/1e6), d["e2e"]["stages_s"]["total_s"])
except Exception as e:
    print(sys.argv[1], "failed", e)
P
}
run default A=1
run split_first2 SCG_INFLATE_ROUTE=split
run default_again A=1
