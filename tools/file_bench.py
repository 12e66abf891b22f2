#!/usr/bin/env python
"""The counting call on FILES (what an R caller passes): raw FASTQ, block gzip and plain gzip of the same reads, page cache warm.
usage: file_bench.py [reads] [dir]"""
import gzip, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from screencounter_b200 import rcpp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
where = sys.argv[2] if len(sys.argv) > 2 else tempfile.mkdtemp(prefix="scg_files_")
wl = bench.make_workload(2)
text = wl.texts(0, n, pinned=False)[0]
paths = {"raw FASTQ": os.path.join(where, "reads.fastq"), "block gzip": os.path.join(where, "reads.bgzf.fastq.gz"),
         "plain gzip": os.path.join(where, "reads.plain.fastq.gz")}
open(paths["raw FASTQ"], "wb").write(text)
open(paths["block gzip"], "wb").write(rcpp.bgzf_compress(text, level=6).tobytes())
plain_n = min(n, 1_000_000)   # one zlib stream inflates on one thread: a smaller file keeps the run short
open(paths["plain gzip"], "wb").write(gzip.compress(text[: plain_n * 157], 6))
threads = len(os.sched_getaffinity(0)) or 1
ref = None
for name, path in paths.items():
    units = plain_n if name == "plain gzip" else n
    for _ in range(2):
        res = wl.ours([path], threads, 0)
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        res = wl.ours([path], threads, 0)
    dt = (time.perf_counter() - t0) / reps
    tm = rcpp.timing(0)
    print("%-10s file (%6.1f MB): %6.1f M reads/s (%.1f ms), h2d %.1f MB, reader: %s" % (
        name, os.path.getsize(path) / 1e6, units / dt / 1e6, dt * 1e3, tm["bytes_h2d"] / 1e6, tm["reader"]))
    if name == "raw FASTQ":
        ref = res
    elif name == "block gzip":
        assert wl.same(ref, res), "block-gzip file gives a different result"
