#!/usr/bin/env python
"""Block-gzip input on the device: throughput of the inflate kernels alone, and the end-to-end counting call on a
block-gzip image in page-locked memory beside the same text uncompressed.
usage: bgzf_bench.py [reads] [level]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from screencounter_b200 import rcpp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
wl = bench.make_workload(2)
text = wl.texts(0, n, pinned=True, device=0)[0]
raw = text.array[: text.size]
t0 = time.perf_counter()
image = rcpp.bgzf_compress(raw, level=level)
print("compressed %d reads: %.1f MB text -> %.1f MB image (ratio %.2f, level %d) in %.1f s" % (
    n, raw.size / 1e6, image.size / 1e6, raw.size / image.size, level, time.perf_counter() - t0))
# the inflate kernels alone, on growing prefixes (whole members)
for mb in (8, 32, 64, 128, 256, 512):
    want = mb << 20
    if want > raw.size:
        break
    # cut the image after the member that ends the prefix: members are 65280 bytes of text
    members = want // 65280
    sub_text = members * 65280
    sub = rcpp.bgzf_compress(raw[:sub_text], level=level)
    best = None
    for _ in range(3):
        out, ms = rcpp.bgzf_inflate(sub)
        best = ms if best is None else min(best, ms)
    assert np.array_equal(out, raw[:sub_text])
    print("inflate + CRC kernels: %4d MiB of text (%5d members): %.3f ms = %.1f GB/s of text" % (mb, members, best, sub_text / best / 1e6))
pinned = rcpp.PinnedText.from_bytes(image.tobytes())
threads = len(os.sched_getaffinity(0)) or 1
for name, src in (("raw text", text), ("block gzip", pinned)):
    for _ in range(2):
        res = wl.ours([src], threads, 0)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        res = wl.ours([src], threads, 0)
    dt = (time.perf_counter() - t0) / reps
    tm = rcpp.timing(0)
    print("%-10s end to end: %.1f M reads/s (%.1f ms; h2d %.1f MB; reader: %s)" % (name, n / dt / 1e6, dt * 1e3, tm["bytes_h2d"] / 1e6, tm["reader"]))
    if name == "raw text":
        ref = res
    else:
        assert wl.same(ref, res), "block-gzip result differs"
