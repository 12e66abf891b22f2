#!/usr/bin/env python
"""Runs the device inflater alone on a block-gzip image of the benchmark's FASTQ text (for ncu captures).
usage: inflate_only.py [MiB of text] [level] [repeats]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from screencounter_b200 import rcpp

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
wl = bench.make_workload(2)
members = (mb << 20) // 65280
n = members * 65280 // 157 + 1
text = wl.texts(0, n, pinned=False)[0]
raw = np.frombuffer(text, dtype=np.uint8)[: members * 65280]
image = rcpp.bgzf_compress(raw, level=level)
for _ in range(reps):
    out, ms = rcpp.bgzf_inflate(image)
assert np.array_equal(out, raw)
print("%d MiB of text, %d members, ratio %.2f: %.3f ms = %.1f GB/s" % (mb, members, raw.size / image.size, ms, raw.size / ms / 1e6))
