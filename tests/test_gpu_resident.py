"""Resident objects (reads + plan in HBM) and the synthetic generator: the device-generated packed
reads must be the reads of the host-generated FASTQ text, and the resident path must agree with the
file-level call and with the reference."""
import numpy as np
import pytest

from engines import GpuEngine
from util import distinct_pool

pytestmark = pytest.mark.gpu

TEMPLATE = "CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG"


@pytest.mark.parametrize("strand", [0, 1, 2])
@pytest.mark.parametrize("read_len", [75, 44, 100, 150])
def test_device_synth_matches_host_text(kref, strand, read_len):
    from screencounter_b200.device import SynthSpec, SinglePlan, DeviceArray
    rng = np.random.default_rng(1)
    pool = distinct_pool(rng, 300, 20)
    spec = SynthSpec(TEMPLATE, [pool], seed=7, read_len=read_len, strand=strand, n_per_10k=50)
    n, first = 20011, 12345
    reads = spec.on_device(first, n)
    assert len(reads) == n
    plan = SinglePlan(TEMPLATE, strand, pool, 1, True)
    counts = DeviceArray(4 * len(pool))
    index = DeviceArray(4 * n)
    plan.run(reads, counts.ptr, index.ptr)
    got_index = index.to_numpy(np.int32)
    got_counts = counts.to_numpy(np.int32)
    text = spec.fastq(first, n)
    want_index, _ = kref.trace_single(text, TEMPLATE, strand, pool, 1, True)
    assert np.array_equal(got_index, want_index)
    assert np.array_equal(got_counts, np.bincount(want_index[want_index >= 0], minlength=len(pool)))
    # the file-level call on the text agrees too
    c2, total = GpuEngine().count_single(text, TEMPLATE, strand, pool, 1, True, 4)
    assert total == n and np.array_equal(c2, got_counts)
    # accumulation: a second pass doubles the counts
    plan.run(reads, counts.ptr)
    assert np.array_equal(counts.to_numpy(np.int32), 2 * got_counts)


def test_reads_from_fastq_resident(kref):
    from screencounter_b200.device import SynthSpec, SinglePlan, DeviceArray, Reads
    rng = np.random.default_rng(2)
    pool = distinct_pool(rng, 100, 20)
    spec = SynthSpec(TEMPLATE, [pool], seed=3)
    text = spec.fastq(0, 5000)
    reads = Reads.from_fastq(text, nthreads=2)
    assert len(reads) == 5000
    plan = SinglePlan(TEMPLATE, 2, pool, 1, False)
    counts = DeviceArray(4 * len(pool))
    plan.run(reads, counts.ptr)
    want, _ = kref.count_single(text, TEMPLATE, 2, pool, 1, False)
    assert np.array_equal(counts.to_numpy(np.int32), want)


@pytest.mark.parametrize("use_first", [True, False])
def test_many_tiles_per_warp(kref, use_first):
    """Enough reads that every persistent warp walks many tiles: exercises the TMA ring, the probe
    carried from one tile to the next and the deferred-read queue (filled, drained mid-way and
    flushed at the end).  Every read must come back, exactly as the reference decides it."""
    from screencounter_b200.device import SynthSpec, SinglePlan, DeviceArray
    rng = np.random.default_rng(5)
    pool = distinct_pool(rng, 2000, 20)
    # noisy constructs: many reads need the mismatch-tolerant search (the deferred path)
    spec = SynthSpec(TEMPLATE, [pool], seed=11, read_len=75, strand=2, sub_per_10k=300, n_per_10k=30)
    n = 1_500_003
    reads = spec.on_device(0, n)
    plan = SinglePlan(TEMPLATE, 2, pool, 1, use_first)
    counts = DeviceArray(4 * len(pool))
    index = DeviceArray(4 * n)
    plan.run(reads, counts.ptr, index.ptr)
    got_index = index.to_numpy(np.int32)
    got_counts = counts.to_numpy(np.int32)
    want_index, _ = kref.trace_single(spec.fastq(0, n), TEMPLATE, 2, pool, 1, use_first)
    assert np.array_equal(got_index, want_index)
    assert np.array_equal(got_counts, np.bincount(want_index[want_index >= 0], minlength=len(pool)))


@pytest.mark.parametrize("mm", [0, 1, 2])
@pytest.mark.parametrize("joint", [True, False])
def test_exact_table_variants(kref, monkeypatch, mm, joint):
    """The uniform-length kernel with the joint exact table of both strands (default) and with the per-strand tables
    (SCG_SPEC_NO_JOINT=1), all three budgets."""
    from screencounter_b200.device import SynthSpec, SinglePlan, DeviceArray
    if not joint:
        monkeypatch.setenv("SCG_SPEC_NO_JOINT", "1")
    rng = np.random.default_rng(6 + mm)
    pool = distinct_pool(rng, 3000, 20)
    spec = SynthSpec(TEMPLATE, [pool], seed=12, read_len=75, strand=2, sub_per_10k=200, n_per_10k=20)
    n = 300_001
    reads = spec.on_device(0, n)
    plan = SinglePlan(TEMPLATE, 2, pool, mm, True)
    counts = DeviceArray(4 * len(pool))
    index = DeviceArray(4 * n)
    plan.run(reads, counts.ptr, index.ptr)
    want_index, _ = kref.trace_single(spec.fastq(0, n), TEMPLATE, 2, pool, mm, True)
    assert np.array_equal(index.to_numpy(np.int32), want_index)
    assert np.array_equal(counts.to_numpy(np.int32), np.bincount(want_index[want_index >= 0], minlength=len(pool)))


@pytest.mark.parametrize("mm", [0, 1, 2])
@pytest.mark.parametrize("use_first", [True, False])
def test_trimmed_reads_take_the_filter_verify_kernel(kref, mm, use_first):
    """Reads of different lengths (adapter-trimmed data), here at most 32 windows each: the filter + verify kernel with per-lane
    window masks; reads shorter than the template have no window at all (ScanTemplate.hpp:153,168-170)."""
    from screencounter_b200 import rcpp
    from util import adversarial_reads, fastq
    rng = np.random.default_rng(40 + mm)
    pool = distinct_pool(rng, 400, 20)
    T = len(TEMPLATE)
    reads = adversarial_reads(rng, 30000, TEMPLATE, [pool], strand="both", sub_rate=0.02, n_rate=0.005, lower_rate=0.01,
                              double_frac=0.0, short_frac=0.05, edge_frac=0.3)
    reads = [r[: T + 31] for r in reads]            # at most 32 windows each
    reads[7] = ""                                     # and the odd empty or tiny read
    reads[8] = "ACGT"
    assert len({len(r) for r in reads}) > 20
    data = fastq(reads)
    want_index, want_info = kref.trace_single(data, TEMPLATE, 2, pool, mm, use_first)
    counts, total, (index, info) = rcpp.count_single_barcodes(data, TEMPLATE, 2, pool, mm, use_first, 4, trace=True)
    assert "trimmed reads, filter+verify" in rcpp.timing()["kernel"], rcpp.timing()
    assert total == len(reads)
    assert np.array_equal(index[:, 0] if index.ndim > 1 else index, want_index)
    assert np.array_equal(info, want_info)
    assert (want_index >= 0).mean() > 0.2


@pytest.mark.parametrize("read_len", [76, 100, 108, 150, 192, 250, 320])
@pytest.mark.parametrize("mm,use_first", [(0, True), (1, True), (1, False), (2, False)])
def test_long_reads_scan_several_window_blocks(kref, read_len, mm, use_first):
    """Reads with more than 32 windows: the filter + verify kernel goes through the window blocks in turn.  Constructs at
    every offset (first window, last window, block boundaries), two constructs in one read (first-versus-best, ties across
    blocks), per-read outcomes (index, strand, mismatches, position) equal to the reference's."""
    from screencounter_b200 import rcpp
    from util import adversarial_reads, fastq, random_seq, fill_template
    rng = np.random.default_rng(1000 + read_len + mm)
    pool = distinct_pool(rng, 500, 20)
    T = len(TEMPLATE)
    reads = adversarial_reads(rng, 20000, TEMPLATE, [pool], strand="both", read_len=read_len, sub_rate=0.02, n_rate=0.004,
                              lower_rate=0.01, double_frac=0.3, short_frac=0.0, edge_frac=0.3)
    reads = [(r + random_seq(rng, read_len))[:read_len] for r in reads]     # the junk reads come back shorter
    # a clean construct at every possible offset
    for off in range(read_len - T + 1):
        c = fill_template(TEMPLATE, pool[off % len(pool)])
        reads.append((random_seq(rng, off) + c + random_seq(rng, read_len))[:read_len])
    assert {len(r) for r in reads} == {read_len}
    data = fastq(reads)
    want_index, want_info = kref.trace_single(data, TEMPLATE, 2, pool, mm, use_first)
    counts, total, (index, info) = rcpp.count_single_barcodes(data, TEMPLATE, 2, pool, mm, use_first, 4, trace=True)
    assert "uniform-length filter+verify" in rcpp.timing()["kernel"], rcpp.timing()
    assert total == len(reads)
    assert np.array_equal(index[:, 0] if index.ndim > 1 else index, want_index)
    assert np.array_equal(info, want_info)
    assert np.array_equal(counts, np.bincount(want_index[want_index >= 0], minlength=len(pool)))
    assert (want_index >= 0).mean() > 0.2


@pytest.mark.parametrize("maxlen", [192, 320])
@pytest.mark.parametrize("mm,use_first", [(1, True), (1, False)])
def test_long_trimmed_reads(kref, mm, use_first, maxlen):
    """Ragged batch whose longest read has 149 (277) windows: every lane masks, block by block, the windows its own read lacks."""
    from screencounter_b200 import rcpp
    from util import adversarial_reads, fastq
    rng = np.random.default_rng(77 + mm)
    pool = distinct_pool(rng, 500, 20)
    reads = adversarial_reads(rng, 30000, TEMPLATE, [pool], strand="both", sub_rate=0.02, n_rate=0.004, lower_rate=0.01,
                              double_frac=0.3, short_frac=0.05, edge_frac=0.3)
    reads = [(r + "ACGT" * 80)[: int(rng.integers(0, maxlen + 1))] if k % 3 == 0 else r for k, r in enumerate(reads)]
    reads[5] = ""
    assert max(len(r) for r in reads) > 150
    data = fastq(reads)
    want_index, want_info = kref.trace_single(data, TEMPLATE, 2, pool, mm, use_first)
    counts, total, (index, info) = rcpp.count_single_barcodes(data, TEMPLATE, 2, pool, mm, use_first, 4, trace=True)
    assert "trimmed reads, filter+verify" in rcpp.timing()["kernel"], rcpp.timing()
    assert total == len(reads)
    assert np.array_equal(index[:, 0] if index.ndim > 1 else index, want_index)
    assert np.array_equal(info, want_info)
    assert (want_index >= 0).mean() > 0.2
