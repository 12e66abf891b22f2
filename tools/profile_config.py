#!/usr/bin/env python
"""Runs the resident hot path of one BASELINE config a few times (for ncu captures, launch lists and knob sweeps).
usage: profile_config.py <config 1..5> [units] [passes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from screencounter_b200 import rcpp

config = int(sys.argv[1])
wl = bench.make_workload(config)
units = int(sys.argv[2]) if len(sys.argv) > 2 else min(wl.default_units, 20_000_000)
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 3
reads = wl.resident(0, units, 0)
wl.make_plan(0, units)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    wl.reset(s)
    wl.run(reads, s)
torch.cuda.synchronize()
before = rcpp.kernel_launches(0)
wl.reset(s)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(passes):
    wl.run(reads, s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / passes
launches = (rcpp.kernel_launches(0) - before) // passes
print("config %d: %s: %d %s, %.3f ms per pass (%d launches), %.2f G/s, frac %.3f" % (
    config, wl.plan.kernel, units, wl.what, ms, launches, units / ms / 1e6, wl.bytes_per_unit * units / ms / 1e6 / bench.measured_peaks()[0]))
