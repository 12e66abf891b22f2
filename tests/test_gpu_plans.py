"""The run-time specialised handler kernels (spec_handlers.cuh: dual paired-end, combinatorial single-end, random
barcodes) and the resident plans built on them, against the compiled reference: adversarial uniform-length reads (dense
libraries with 1-substitution neighbours, N in flanks and regions, several constructs per read, constructs at the read
ends) with per-read outcomes, both through the file-level entry points and through scg_*_plan_run on resident reads; the
same runs with the specialised kernels switched off must agree too."""
import os

import numpy as np
import pytest

from engines import GpuEngine
from test_oracle_fuzz import _same, STRANDS
from util import fastq, random_seq, dense_pool, distinct_pool, adversarial_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    return GpuEngine()


def _uniform(rng, n, template, pools, strand, read_len, **kw):
    kw.setdefault("lower_rate", 0.0)
    kw.setdefault("short_frac", 0.0)
    reads = adversarial_reads(rng, n, template, pools, strand=strand, read_len=read_len, **kw)
    # junk reads come back at their own lengths: every read exactly read_len bases
    return [(r + random_seq(rng, read_len))[:read_len] for r in reads]


def _kernel():
    from screencounter_b200 import rcpp
    return rcpp.timing().get("kernel", "")


# ---- random barcodes ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("strand", ["original", "reverse", "both"])
@pytest.mark.parametrize("mm", [0, 1, 2])
@pytest.mark.parametrize("use_first", [True, False])
@pytest.mark.parametrize("template,read_len", [("AAAAACGT------ACGTGGGG", 40), ("AAAAACGTCC------ACGT", 33),   # asymmetric flanks: Quirk B
                                               ("CAGCTACGTACG" + "-" * 16 + "CCAGCTCGATCG", 75), ("ACGTACGG" + "-" * 21 + "GGTCATTA", 101)])
def test_random_specialised(gpu, kref, strand, mm, use_first, template, read_len):
    rng = np.random.default_rng(70 + mm + 100 * STRANDS[strand] + read_len)
    L = template.count("-")
    pool = [random_seq(rng, L) for _ in range(40)]
    reads = _uniform(rng, 3000, template, [pool], strand, read_len, double_frac=0.2)
    f = fastq(reads)
    _same(kref.count_random(f, template, STRANDS[strand], mm, use_first), gpu.count_random(f, template, STRANDS[strand], mm, use_first))
    assert "spec_random_kernel" in _kernel()


def test_random_specialised_raw_characters(gpu, kref):
    """Lower case and IUPAC letters inside uniform-length reads: the device decides where the barcode sits, the host
    renders it from the raw text (handlers/RandomBarcodeSingleEnd.hpp:93-120)."""
    template = "ACGTACGT----TTTTGGGG"
    body = ["ACGTACGTacgtTTTTGGGG", "ACGTACGTNNgtTTTTGGGG", "ACGTACGTAC.RTTTTGGGG", "ACGTACGTACGTTTTTGGGG"]
    reads = [("CC" + b + "AAAAAAAAAA")[:30] for b in body] * 40
    f = fastq(reads)
    _same(kref.count_random(f, template, 0, 0, True), gpu.count_random(f, template, 0, 0, True))
    _same(kref.count_random(f, template, 2, 1, False), gpu.count_random(f, template, 2, 1, False))


def test_random_plan_resident(kref):
    from screencounter_b200.device import SynthSpec, RandomPlan, DeviceArray
    template = "CAGCTACGTACG" + "-" * 16 + "CCAGCTCGATCG"
    spec = SynthSpec(template, [], seed=13, read_len=75, strand=2, random_space=50_000, n_per_10k=30)
    n = 400_003
    reads = spec.on_device(0, n)
    text = spec.fastq(0, n)
    for use_first in (True, False):
        want_seqs, want_freq, want_total = kref.count_random(text, template, 2, 1, use_first)
        order = np.argsort(np.array(want_seqs, dtype=object), kind="stable")
        for expected in (0, 200_000):   # a table that grows, and one sized up front
            plan = RandomPlan(template, 2, 1, use_first, expected_distinct=expected)
            index = DeviceArray(4 * n)
            plan.run(reads, index.ptr)
            seqs, freq = plan.harvest()
            assert "spec_random_kernel" in plan.kernel
            assert [s.decode() for s in seqs] == [want_seqs[i] for i in order]
            assert np.array_equal(freq, np.asarray(want_freq)[order])
            assert int((index.to_numpy(np.int32) >= 0).sum()) == int(freq.sum())
            # reset + run again gives the same table; two runs without a reset double it
            plan.reset()
            plan.run(reads)
            plan.run(reads)
            seqs2, freq2 = plan.harvest()
            assert np.array_equal(seqs2, seqs) and np.array_equal(freq2, 2 * freq)
            plan.free()


def test_random_plan_table_overflow_is_reported():
    from screencounter_b200.device import SynthSpec, RandomPlan
    from screencounter_b200 import ScreenCounterError
    template = "CAGCTACGTACG" + "-" * 16 + "CCAGCTCGATCG"
    spec = SynthSpec(template, [], seed=13, read_len=75, strand=0, random_space=1_000_000)
    reads = spec.on_device(0, 300_000)
    plan = RandomPlan(template, 0, 0, True, expected_distinct=500)   # 1024 slots for ~250k distinct barcodes
    plan.run(reads)
    with pytest.raises(ScreenCounterError, match="overflowed"):
        plan.harvest()


# ---- combinatorial, single-end ------------------------------------------------------------------------------------
@pytest.mark.parametrize("strand", ["original", "reverse", "both"])
@pytest.mark.parametrize("mm", [0, 1, 2])
@pytest.mark.parametrize("use_first", [True, False])
def test_combo_specialised(gpu, kref, strand, mm, use_first):
    rng = np.random.default_rng(31 + mm + 100 * STRANDS[strand])
    p1 = dense_pool(rng, 25, 5)
    p2 = dense_pool(rng, 30, 7)
    template = "ACGTCC" + "-" * 5 + "TGCAAG" + "-" * 7 + "GGATTC"
    reads = _uniform(rng, 4000, template, [p1, p2], strand, 50, double_frac=0.2)
    f = fastq(reads)
    _same(kref.count_combo_single(f, template, STRANDS[strand], p1, p2, mm, use_first),
          gpu.count_combo_single(f, template, STRANDS[strand], p1, p2, mm, use_first))
    assert "spec_combo_kernel" in _kernel()
    assert np.array_equal(kref.trace_combo_single(f, template, STRANDS[strand], p1, p2, mm, use_first),
                          gpu.trace_combo_single(f, template, STRANDS[strand], p1, p2, mm, use_first))


@pytest.mark.parametrize("npool", [500, 5000])
def test_combo_plan_resident(kref, npool):
    """BASELINE configs[3] shape; the 5000 x 5000 pools force the sparse (hash) tally."""
    from screencounter_b200.device import SynthSpec, ComboPlan, DeviceArray
    rng = np.random.default_rng(4)
    p1, p2 = distinct_pool(rng, npool, 20), distinct_pool(rng, npool, 20)
    template = "CAGCTACG" + "-" * 20 + "GGTACCTT" + "-" * 20 + "CGATCGAG"
    spec = SynthSpec(template, [p1, p2], seed=11, read_len=75, strand=2, n_per_10k=30)
    n = 300_001
    reads = spec.on_device(0, n)
    text = spec.fastq(0, n)
    for mm, use_first in ((0, True), (1, True), (1, False)):
        wkeys, wfreq, wtotal = kref.count_combo_single(text, template, 2, p1, p2, mm, use_first)
        wtrace = kref.trace_combo_single(text, template, 2, p1, p2, mm, use_first)
        plan = ComboPlan(template, 2, p1, p2, mm, use_first)
        pairs = DeviceArray(8 * n)
        plan.run(reads, pairs.ptr)
        keys, freq = plan.harvest()
        assert "spec_combo_kernel" in plan.kernel
        assert np.array_equal(keys, wkeys) and np.array_equal(freq, wfreq)
        assert np.array_equal(pairs.to_numpy(np.int32).reshape(n, 2), wtrace)
        plan.reset()
        plan.run(reads)
        keys2, freq2 = plan.harvest()
        assert np.array_equal(keys2, keys) and np.array_equal(freq2, freq)
        plan.free()


# ---- dual, paired-end ---------------------------------------------------------------------------------------------
def _dual_library(rng, n1, n2, len1, len2, npairs, neighbours=0.5):
    g1 = dense_pool(rng, n1, len1, neighbours)
    g2 = dense_pool(rng, n2, len2, neighbours)
    # last-base variants of the second segment so that the phantom of MismatchTrie.hpp:608-617 ("Quirk A") fires
    for k in range(3):
        s = g2[k]
        alt = s[:-1] + ("A" if s[-1] != "A" else "C")
        if alt not in g2:
            g2.append(alt)
    pairs = set()
    while len(pairs) < npairs:
        pairs.add((int(rng.integers(0, len(g1))), int(rng.integers(0, len(g2)))))
    pairs = sorted(pairs)
    return g1, g2, [g1[i] for i, _ in pairs], [g2[j] for _, j in pairs]


@pytest.mark.parametrize("rev", [(False, False), (True, False), (False, True)])
@pytest.mark.parametrize("mms", [(0, 0), (1, 1), (1, 0), (0, 1), (2, 1), (2, 0), (1, 2)])
@pytest.mark.parametrize("use_first", [True, False])
def test_dual_specialised(gpu, kref, rev, mms, use_first):
    rng = np.random.default_rng(91 + 7 * mms[0] + 3 * mms[1] + rev[0] + 2 * rev[1])
    g1, g2, p1, p2 = _dual_library(rng, 12, 12, 5, 6, 50)
    t1 = "ACGTACCA" + "-" * 5 + "TGCATGGT"
    t2 = "GGATCCTT" + "-" * 6 + "CCTAGGAA"
    r1 = _uniform(rng, 4000, t1, [g1], "reverse" if rev[0] else "original", 45, double_frac=0.15)
    r2 = _uniform(rng, 4000, t2, [g2], "reverse" if rev[1] else "original", 52, double_frac=0.15)
    f1, f2 = fastq(r1), fastq(r2)
    args = (f1, t1, rev[0], mms[0], p1, f2, t2, rev[1], mms[1], p2, False, use_first)
    want = kref.trace_dual(*args, fresh_state=2)
    got = gpu.trace_dual(*args)
    assert "spec_dual_pe_kernel" in _kernel()
    assert np.array_equal(got, want)
    # counts from the cache-free per-pair outcomes (the reference's own counts can depend on read order)
    counts, total = gpu.count_dual(*args)
    assert total == len(r1) and np.array_equal(counts, np.bincount(want[want >= 0], minlength=len(p1)))


def test_dual_specialised_long_keys(gpu, kref):
    """Two 24-base regions: the pair's key spans two words (48 bases, the longest the specialised kernel takes)."""
    rng = np.random.default_rng(5)
    g1, g2 = distinct_pool(rng, 40, 24), distinct_pool(rng, 40, 24)
    rows = sorted({(int(rng.integers(0, 40)), int(rng.integers(0, 40))) for _ in range(300)})
    p1, p2 = [g1[i] for i, _ in rows], [g2[j] for _, j in rows]
    t1 = "ACGTACCAGT" + "-" * 24 + "TGCATGGTCA"
    t2 = "GGATCCTTAG" + "-" * 24 + "CCTAGGAATC"
    r1 = _uniform(rng, 5000, t1, [g1], "original", 75, sub_rate=0.02, n_rate=0.003)
    r2 = _uniform(rng, 5000, t2, [g2], "original", 75, sub_rate=0.02, n_rate=0.003)
    f1, f2 = fastq(r1), fastq(r2)
    for mms in ((1, 1), (2, 1), (0, 0)):
        args = (f1, t1, False, mms[0], p1, f2, t2, False, mms[1], p2, False, True)
        assert np.array_equal(gpu.trace_dual(*args), kref.trace_dual(*args, fresh_state=2))
        assert "spec_dual_pe_kernel" in _kernel()


def test_dual_plan_resident(kref):
    """BASELINE configs[2] shape on resident reads: the plan's counts and per-pair rows against the reference."""
    from screencounter_b200.device import SynthSpec, DualPlan, DeviceArray
    rng = np.random.default_rng(3)
    a, b = distinct_pool(rng, 200, 20), distinct_pool(rng, 200, 20)
    rows = set()
    while len(rows) < 10000:
        rows.add((int(rng.integers(0, 200)), int(rng.integers(0, 200))))
    rows = sorted(rows)
    pool1, pool2 = [a[i] for i, _ in rows], [b[j] for _, j in rows]
    t1 = "CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG"
    t2 = "GATTACAGGCTA" + "-" * 20 + "TTGACCGTAGCA"
    # the same seed picks the same row, offset and noise pattern for both mates
    s1 = SynthSpec(t1, [pool1], seed=7, read_len=75, strand=0, n_per_10k=30)
    s2 = SynthSpec(t2, [pool2], seed=7, read_len=75, strand=0, n_per_10k=30)
    n = 500_003
    d1, d2 = s1.on_device(0, n), s2.on_device(0, n)
    f1, f2 = s1.fastq(0, n), s2.fastq(0, n)
    for mms, use_first in (((1, 1), True), ((1, 1), False), ((0, 1), True)):
        want = kref.trace_dual(f1, t1, False, mms[0], pool1, f2, t2, False, mms[1], pool2, False, use_first, fresh_state=2)
        plan = DualPlan(t1, False, mms[0], pool1, t2, False, mms[1], pool2, False, use_first)
        counts = DeviceArray(4 * len(pool1))
        index = DeviceArray(4 * n)
        plan.run(d1, d2, counts.ptr, index.ptr)
        got = index.to_numpy(np.int32)
        assert "spec_dual_pe_kernel" in plan.kernel
        assert np.array_equal(got, want)
        assert np.array_equal(counts.to_numpy(np.int32), np.bincount(want[want >= 0], minlength=len(pool1)))
        plan.free()


# ---- the same designs with the specialised kernels switched off ---------------------------------------------------
def test_generic_kernels_still_agree(gpu, kref, monkeypatch):
    monkeypatch.setenv("SCG_NO_SPEC_HANDLERS", "1")
    rng = np.random.default_rng(17)
    template = "CAGCTACGTACG" + "-" * 16 + "CCAGCTCGATCG"
    reads = _uniform(rng, 2000, template, [[random_seq(rng, 16) for _ in range(30)]], "both", 75)
    f = fastq(reads)
    _same(kref.count_random(f, template, 2, 1, False), gpu.count_random(f, template, 2, 1, False))
    assert "generic random_kernel" in _kernel()
    p1, p2 = dense_pool(rng, 25, 5), dense_pool(rng, 30, 7)
    t = "ACGTCC" + "-" * 5 + "TGCAAG" + "-" * 7 + "GGATTC"
    f = fastq(_uniform(rng, 2000, t, [p1, p2], "both", 50))
    _same(kref.count_combo_single(f, t, 2, p1, p2, 1, True), gpu.count_combo_single(f, t, 2, p1, p2, 1, True))
    assert "generic combo_kernel" in _kernel()


# ---- small pools: counters privatised in shared memory (single-barcode kernel) ------------------------------------
@pytest.mark.parametrize("npool", [7, 1000, 1024, 1025])
def test_single_shared_memory_histogram(kref, npool):
    from screencounter_b200.device import SynthSpec, SinglePlan, DeviceArray
    rng = np.random.default_rng(npool)
    template = "CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG"
    pool = distinct_pool(rng, npool, 20)
    spec = SynthSpec(template, [pool], seed=5, read_len=75, strand=0)
    n = 300_017
    reads = spec.on_device(0, n)
    plan = SinglePlan(template, 0, pool, 0, True)
    counts = DeviceArray(4 * npool)
    index = DeviceArray(4 * n)
    plan.run(reads, counts.ptr, index.ptr)
    want, total = kref.count_single(spec.fastq(0, n), template, 0, pool, 0, True)
    assert np.array_equal(counts.to_numpy(np.int32), want)
    got_index = index.to_numpy(np.int32)
    assert np.array_equal(np.bincount(got_index[got_index >= 0], minlength=npool), want)
