#!/usr/bin/env python
"""Runs the resident single-barcode path a few times on the bench workload (for ncu captures and knob sweeps).
usage: profile_single.py [n_reads] [launches] [read_len]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from screencounter_b200.device import SynthSpec, SinglePlan

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
read_len = int(sys.argv[3]) if len(sys.argv) > 3 else 75
lib = bench.make_library()
spec = SynthSpec(bench.TEMPLATE, [lib], seed=42, read_len=read_len, strand=2)
reads = spec.on_device(0, n, device=0)
plan = SinglePlan(bench.TEMPLATE, 2, lib, 1, True, device=0)
counts = torch.zeros(len(lib), dtype=torch.int32, device="cuda")
index = torch.empty(n, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    plan.run(reads, counts.data_ptr(), index_ptr=index.data_ptr(), stream=s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(launches):
    plan.run(reads, counts.data_ptr(), index_ptr=index.data_ptr(), stream=s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / launches
print("%s: %d reads, %.3f ms per pass, %.2f G reads/s, matched %d" % (plan.kernel, n, ms, n / ms / 1e6, int((index >= 0).sum())))
