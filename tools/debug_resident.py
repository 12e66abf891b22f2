#!/usr/bin/env python
"""Resident single-barcode path against the oracle, read by read, on the bench workload."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from screencounter_b200.device import SynthSpec, SinglePlan
from oracle import kref

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
lib = bench.make_library()
spec = SynthSpec(bench.TEMPLATE, [lib], seed=42, read_len=75, strand=2)
reads = spec.on_device(0, n, device=0)
plan = SinglePlan(bench.TEMPLATE, 2, lib, 1, True, device=0)
counts = torch.zeros(len(lib), dtype=torch.int32, device="cuda")
index = torch.full((n,), -7, dtype=torch.int32, device="cuda")
plan.run(reads, counts.data_ptr(), index_ptr=index.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("kernel:", plan.kernel)
idx = index.cpu().numpy()
cnt = counts.cpu().numpy()
text = spec.fastq(0, n)
want, _ = kref.trace_single(text, bench.TEMPLATE, 2, lib, 1, True)
bad = np.nonzero(idx != want)[0]
print("reads", n, "mismatching reads", len(bad), "unwritten", int((idx == -7).sum()))
print("counts ok", np.array_equal(cnt, np.bincount(want[want >= 0], minlength=len(lib))), "sum", cnt.sum(), (want >= 0).sum())
for b in bad[:20]:
    print("  read", b, "tile", b // 32, "lane", b % 32, "got", idx[b], "want", want[b])
if len(bad):
    tiles = bad // 32
    print("distinct tiles", len(np.unique(tiles)), "first tiles", np.unique(tiles)[:20])
