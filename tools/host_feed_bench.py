#!/usr/bin/env python
"""Host feed rate (FASTQ text -> records -> tile-planar packed words) against the number of host
threads; no device involved.  usage: host_feed_bench.py [reads]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from screencounter_b200._lib import lib
from screencounter_b200.device import SynthSpec
from screencounter_b200.rcpp import _Src

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
library = bench.make_library()
text = SynthSpec(bench.TEMPLATE, [library], seed=42, read_len=75, strand=2).fastq(0, n)
print("cores", os.cpu_count(), "reads", n, "text MB", len(text) >> 20)
for threads in (1, 2, 4, 8, 16, 32):
    best = 1e9
    for _ in range(3):
        src = _Src(text)
        nr, nb = C.c_longlong(), C.c_longlong()
        t0 = time.perf_counter()
        assert lib().scg_host_pack_roundtrip(src.ref(), threads, None, None, C.byref(nr), C.byref(nb)) == 0
        best = min(best, time.perf_counter() - t0)
    print("threads %2d: %.3f s  %.1f M reads/s  %.2f GB/s of text" % (threads, best, n / best / 1e6, len(text) / best / 1e9))
