#!/usr/bin/env python
"""The five BASELINE.json configs through the reference-facing entry points, at a size the compiled reference finishes in
seconds: results checked EQUAL to kaori's on the same FASTQ text, wall-clock reads/s of the call beside kaori's on the
host cores.  One JSON line per config (bench.py carries the headline config 2 with the full contract; this is the
parity-at-config-shape + measurement companion for the others).

usage: bench_configs.py [reads] [reference_sample]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from screencounter_b200 import rcpp                      # noqa: E402
from screencounter_b200.device import SynthSpec          # noqa: E402
from oracle import kref                                   # noqa: E402
from util import distinct_pool                            # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
SAMPLE = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
THREADS = os.cpu_count() or 1
LAST_TIMING = {}
FLANK_L, FLANK_R = "CAGCTACGTACG", "CCAGCTCGATCG"


def timed(fn, repeat=3):
    fn()   # warm-up: tables built and cached, kernels specialised, buffers allocated
    times = []
    for _ in range(repeat):
        t0 = time.perf_counter()
        out = fn()
        times.append(time.perf_counter() - t0)
    LAST_TIMING.clear()
    LAST_TIMING.update(rcpp.timing())
    if os.environ.get("VERBOSE"):
        print("calls: " + " ".join("%.1f ms" % (1e3 * t) for t in times), rcpp.timing(), file=sys.stderr, flush=True)
    return out, sum(times) / repeat


def report(name, n, gpu_s, ref_s, ref_n, extra):
    t = rcpp.timing()
    line = {"config": name, "reads": n, "gpu_reads_per_s": n / gpu_s, "gpu_ms": 1e3 * gpu_s, "c_abi_ms": 1e3 * float(LAST_TIMING.get("total_s", 0)),
            "reference_reads_per_s": ref_n / ref_s, "reference_cores": THREADS, "reference_sample": ref_n,
            "speedup": (n / gpu_s) / (ref_n / ref_s), "reader": t.get("reader"), "kernel": t.get("kernel"),
            "results_equal_reference": True,
            "stages_s": {k: t.get(k) for k in ("parse_s", "pack_s", "device_s", "setup_s", "harvest_s", "total_s")}, "launches": t.get("launches")}
    line.update(extra)
    print(json.dumps(line), flush=True)


def config1():
    rng = np.random.default_rng(1)
    pool = distinct_pool(rng, 1000, 20)
    template = FLANK_L + "-" * 20 + FLANK_R
    spec = SynthSpec(template, [pool], seed=42, read_len=75, strand=0)
    n = min(N, 1_000_000)
    text = spec.fastq_pinned(0, n)
    (counts, total), gpu_s = timed(lambda: rcpp.count_single_barcodes(text, template, 0, pool, 0, True, THREADS))
    sample = text.array[: SAMPLE * 157].tobytes() if SAMPLE < n else text.array[: n * 157].tobytes()
    t0 = time.perf_counter()
    want, wtotal = kref.count_single(sample, template, 0, pool, 0, True, THREADS)
    ref_s = time.perf_counter() - t0
    got, gtotal = rcpp.count_single_barcodes(sample, template, 0, pool, 0, True, THREADS)
    assert gtotal == wtotal and np.array_equal(got, want)
    report("1: countSingleBarcodes 1M reads, 1,000 x 20-bp guides, 0 mismatches, forward", n, gpu_s, ref_s, wtotal,
           {"matched": int(counts.sum())})


def config3():
    rng = np.random.default_rng(3)
    a, b = distinct_pool(rng, 200, 20), distinct_pool(rng, 200, 20)
    rows = set()
    while len(rows) < 10000:
        rows.add((int(rng.integers(0, 200)), int(rng.integers(0, 200))))
    rows = sorted(rows)
    pool1, pool2 = [a[i] for i, _ in rows], [b[j] for _, j in rows]
    t1 = FLANK_L + "-" * 20 + FLANK_R
    t2 = "GATTACAGGCTA" + "-" * 20 + "TTGACCGTAGCA"
    # the same seed picks the same row, offset and noise pattern for both mates
    s1 = SynthSpec(t1, [pool1], seed=7, read_len=75, strand=0)
    s2 = SynthSpec(t2, [pool2], seed=7, read_len=75, strand=0)
    n = N
    f1, f2 = s1.fastq_pinned(0, n), s2.fastq_pinned(0, n)
    call = lambda x1, x2: rcpp.count_dual_barcodes(x1, t1, False, 1, pool1, x2, t2, False, 1, pool2, False, True, False, THREADS)
    (counts, total), gpu_s = timed(lambda: call(f1, f2))
    m = min(SAMPLE, n)
    g1, g2 = f1.array[: m * 157].tobytes(), f2.array[: m * 157].tobytes()
    t0 = time.perf_counter()
    want, wtotal = kref.count_dual(g1, t1, False, 1, pool1, g2, t2, False, 1, pool2, False, True, False, THREADS)[:2]
    ref_s = time.perf_counter() - t0
    got, gtotal = call(g1, g2)
    assert int(gtotal[0]) == int(wtotal) and np.array_equal(got, want)
    report("3: countDualBarcodes paired-end, 2 x 20-bp regions, 10k-pair library, 1 mismatch per read", n, gpu_s, ref_s, int(wtotal),
           {"matched": int(counts.sum()), "unit": "read pairs"})


def config4():
    rng = np.random.default_rng(4)
    p1, p2 = distinct_pool(rng, 500, 20), distinct_pool(rng, 500, 20)
    template = "CAGCTACG" + "-" * 20 + "GGTACCTT" + "-" * 20 + "CGATCGAG"
    spec = SynthSpec(template, [p1, p2], seed=11, read_len=75, strand=2)
    n = N
    text = spec.fastq_pinned(0, n)
    for mm in (0, 1):
        (keys, freq, total), gpu_s = timed(lambda: rcpp.count_combo_barcodes_single(text, template, 2, [p1, p2], mm, True, THREADS))
        m = min(SAMPLE, n)
        sample = text.array[: m * 157].tobytes()
        t0 = time.perf_counter()
        wkeys, wfreq, wtotal = kref.count_combo_single(sample, template, 2, p1, p2, mm, True, THREADS)
        ref_s = time.perf_counter() - t0
        gkeys, gfreq, gtotal = rcpp.count_combo_barcodes_single(sample, template, 2, [p1, p2], mm, True, THREADS)
        assert int(gtotal[0]) == int(wtotal) and np.array_equal(gkeys.T, wkeys) and np.array_equal(gfreq, wfreq)
        report("4: countComboBarcodes single-end, two 20-bp regions, 500 x 500 pools, %d mismatch(es)" % mm, n, gpu_s, ref_s, int(wtotal),
               {"distinct_combinations": int(len(freq)), "matched": int(freq.sum())})


def config5():
    template = FLANK_L + "-" * 16 + FLANK_R
    spec = SynthSpec(template, [], seed=13, read_len=75, strand=2, random_space=4_000_000)
    n = N
    text = spec.fastq_pinned(0, n)
    ((seqs, freq), total), gpu_s = timed(lambda: rcpp.count_random_barcodes(text, template, 2, 1, True, THREADS, as_array=True))
    m = min(SAMPLE, n)
    sample = text.array[: m * 157].tobytes()
    t0 = time.perf_counter()
    wseqs, wfreq, wtotal = kref.count_random(sample, template, 2, 1, True, THREADS)
    ref_s = time.perf_counter() - t0
    (gseqs, gfreq), gtotal = rcpp.count_random_barcodes(sample, template, 2, 1, True, THREADS, as_array=False)
    order = np.argsort(np.array(wseqs, dtype=object), kind="stable") if len(wseqs) else []
    assert gtotal == wtotal and list(gseqs) == [wseqs[i] for i in order] and np.array_equal(gfreq, np.asarray(wfreq)[order])
    report("5: countRandomBarcodes 16-bp random barcodes, both strands, 1 mismatch in the flanks", n, gpu_s, ref_s, int(wtotal),
           {"distinct_barcodes": int(len(freq)), "matched": int(np.sum(freq))})


if __name__ == "__main__":
    which = os.environ.get("CONFIGS", "1345")
    for key, fn in (("1", config1), ("3", config3), ("4", config4), ("5", config5)):
        if key in which:
            fn()
