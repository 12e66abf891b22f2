#!/usr/bin/env python
"""ONE process, ONE counting call, several GPUs behind it (scg_ctx_create_multi): config 2's workload from raw text and from a
block-gzip image in page-locked memory.  usage: multi_ctx_bench.py [reads] [n_devices]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from screencounter_b200 import rcpp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
ndev = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = bench.make_workload(2)
text = wl.texts(0, n, pinned=True, device=0)[0]
image = rcpp.PinnedText.from_bytes(rcpp.bgzf_compress(text.array[: text.size], level=6).tobytes(), device=0)
threads = len(os.sched_getaffinity(0)) or 1
ref = None
for devices in [0] + [tuple(range(k)) for k in range(2, ndev + 1)]:
    for name, src in (("raw text", text), ("block gzip", image)):
        for _ in range(2):
            res = wl.ours([src], threads, devices)
        t0 = time.perf_counter()
        for _ in range(3):
            res = wl.ours([src], threads, devices)
        dt = (time.perf_counter() - t0) / 3
        if ref is None:
            ref = res
        assert wl.same(ref, res), "result differs"
        print("devices %-12s %-10s: %7.1f M reads/s (%.1f ms)  %s" % (devices, name, n / dt / 1e6, dt * 1e3, rcpp.timing(devices)["kernel"][-12:]))
