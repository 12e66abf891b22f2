#!/bin/bash
# round 2, final single-GPU sweep: every config at its stated size, the reference arm, the launch list of the headline command
mkdir -p gpurun_out
for c in 1 4 5; do
  timeout 900 python bench.py --config $c > gpurun_out/r2_final_c$c.json 2> gpurun_out/r2_final_c$c.err || { echo "config $c FAILED"; tail -8 gpurun_out/r2_final_c$c.err; }
done
timeout 900 python bench.py > gpurun_out/r2_final_c2.json 2> gpurun_out/r2_final_c2.err || { echo "config 2 FAILED"; tail -8 gpurun_out/r2_final_c2.err; }
timeout 900 python bench.py --config 3 > gpurun_out/r2_final_c3.json 2> gpurun_out/r2_final_c3.err || { echo "config 3 FAILED"; tail -8 gpurun_out/r2_final_c3.err; }
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err || { echo "reference arm FAILED"; tail -8 gpurun_out/r2_final_reference.err; }
python - <<PY
import json
for c in (1,2,3,4,5):
    try:
        d=json.loads(open("gpurun_out/r2_final_c%d.json"%c).read())
        print("config %d: %.2f G/s frac %.3f | e2e %.1f M/s raw %.1f M/s cold %.3f s | cpu %.2f M/s" % (c, d["value"]/1e9, d["roofline"]["frac"], d["e2e"]["value"]/1e6, d["e2e"]["raw_text"]["value"]/1e6, d["e2e"]["cold"]["seconds"], d["cpu_baseline"]["value"]/1e6))
    except Exception as e:
        print("config", c, "no line", e)
try:
    d=json.loads(open("gpurun_out/r2_final_reference.json").read()); print("reference arm: %.2f M/s" % (d["value"]/1e6), d.get("cpu_baseline",{}).get("mode"))
except Exception as e:
    print("reference: no line", e)
PY
SCG_BENCH_NO_BGZF=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --cpu-reads 100000 --e2e-reads 1000000 > gpurun_out/r2_final_ncu.log 2>&1
ncu --set full --clock-control none -k regex:'spec_random|random_count_parts' --launch-skip 4 -c 2 -f -o gpurun_out/r2_final_c5_full python tools/profile_config.py 5 200000000 1 > /dev/null 2>&1
ls -la gpurun_out/r2_final_c5_full.ncu-rep
