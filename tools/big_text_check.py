#!/usr/bin/env python
"""A FASTQ text larger than 4 GiB through the device-side reader: counts must equal those of the same reads generated
directly on the device (offsets past 2^32, many chunk slots reused).  usage: big_text_check.py [reads]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from screencounter_b200 import rcpp
from screencounter_b200.device import SynthSpec, SinglePlan

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30_000_000
lib = bench.make_library()
spec = SynthSpec(bench.TEMPLATE, [lib], seed=42, read_len=75, strand=2)
text = spec.fastq_pinned(0, n)
print("text bytes", text.size, "(> 2^32: %s)" % (text.size > 2 ** 32))
t0 = time.perf_counter()
counts, total = rcpp.count_single_barcodes(text, bench.TEMPLATE, 2, lib, 1, True, 16)
dt = time.perf_counter() - t0
print("e2e %.1f M reads/s" % (n / dt / 1e6), rcpp.timing()["reader"])
reads = spec.on_device(0, n)
plan = SinglePlan(bench.TEMPLATE, 2, lib, 1, True)
resident = torch.zeros(len(lib), dtype=torch.int32, device="cuda")
plan.run(reads, resident.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
assert total == n
assert np.array_equal(counts, resident.cpu().numpy()), "counts differ"
print("counts equal:", int(counts.sum()), "matched of", n)
