import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from screencounter_b200 import rcpp
from screencounter_b200.device import SynthSpec
from util import distinct_pool
rng = np.random.default_rng(4)
p1, p2 = distinct_pool(rng, 500, 20), distinct_pool(rng, 500, 20)
template = "CAGCTACG" + "-" * 20 + "GGTACCTT" + "-" * 20 + "CGATCGAG"
spec = SynthSpec(template, [p1, p2], seed=11, read_len=75, strand=2)
for n in (1000, 200000, 4000000):
    text = spec.fastq_pinned(0, n)
    for mm in (0, 1, 2):
        for strand in (2, 0):
            rcpp.count_combo_barcodes_single(text, template, strand, [p1, p2], mm, True, 16)
            t0 = time.perf_counter()
            for _ in range(3):
                rcpp.count_combo_barcodes_single(text, template, strand, [p1, p2], mm, True, 16)
            dt = (time.perf_counter() - t0) / 3
            print("n=%d mm=%d strand=%d: %.2f ms  %s" % (n, mm, strand, dt * 1e3, {k: v for k, v in rcpp.timing().items() if k.endswith("_s")}), flush=True)
