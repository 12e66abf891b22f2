#!/usr/bin/env python
"""One end-to-end counting call on a block-gzip image of the benchmark's FASTQ text in page-locked memory, for launch lists
(`ncu --metrics gpu__time_duration.sum`): two calls, the second is the one to read.  usage: e2e_launches.py [reads]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from screencounter_b200 import rcpp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
wl = bench.make_workload(2)
text = wl.texts(0, n, pinned=True, device=0)[0]
image = rcpp.bgzf_compress(text.array[: text.size], level=6)
pinned = rcpp.PinnedText.from_bytes(image.tobytes())
threads = len(os.sched_getaffinity(0)) or 1
for _ in range(2):
    wl.ours([pinned], threads, 0)
print(rcpp.timing(0))
