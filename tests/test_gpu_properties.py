"""Size-independent properties of the counting paths at sizes the reference does not finish quickly (millions of reads):
results must not depend on how the reads are sharded (the multi-GPU partition), on which reader fed them
(device-side or host), on chunk sizes, or on the number of host threads; totals must add up."""
import numpy as np
import pytest

from screencounter_b200 import rcpp
from screencounter_b200.device import SynthSpec
from util import distinct_pool

pytestmark = pytest.mark.gpu

FLANK_L, FLANK_R = "CAGCTACGTACG", "CCAGCTCGATCG"
REC = 2 * 75 + 7   # bytes per synthetic record


def _slices(text, n, parts):
    cuts = [n * k // parts for k in range(parts + 1)]
    return [text[cuts[k] * REC:cuts[k + 1] * REC] for k in range(parts)]


def test_single_counts_are_shard_and_reader_invariant(monkeypatch):
    rng = np.random.default_rng(31)
    pool = distinct_pool(rng, 20000, 20)
    template = FLANK_L + "-" * 20 + FLANK_R
    n = 3_000_000
    text = SynthSpec(template, [pool], seed=5, read_len=75, strand=2).fastq(0, n)
    whole, total = rcpp.count_single_barcodes(text, template, 2, pool, 1, True, 8)
    assert total == n and 0.7 * n < whole.sum() < 0.95 * n
    for parts in (2, 7):   # contiguous shards, as the ranks of a multi-GPU run take them
        acc = np.zeros_like(whole)
        seen = 0
        for piece in _slices(text, n, parts):
            c, t = rcpp.count_single_barcodes(piece, template, 2, pool, 1, True, 8)
            acc += c
            seen += t
        assert seen == n and np.array_equal(acc, whole)
    monkeypatch.setenv("SCG_HOST_PARSE", "1")
    host, _ = rcpp.count_single_barcodes(text, template, 2, pool, 1, True, 3)
    assert np.array_equal(host, whole)
    monkeypatch.delenv("SCG_HOST_PARSE")
    monkeypatch.setenv("SCG_INGEST_CHUNK", str(5_000_001))
    odd, _ = rcpp.count_single_barcodes(text, template, 2, pool, 1, True, 1)
    assert np.array_equal(odd, whole)
    # best mode can only lose matches to ties, never find different ones at these noise levels; it sees the same reads
    best, tb = rcpp.count_single_barcodes(text, template, 2, pool, 1, False, 8)
    assert tb == n and best.sum() <= whole.sum() and best.sum() > 0.99 * whole.sum()


def test_combo_tables_merge_like_counts(monkeypatch):
    rng = np.random.default_rng(32)
    p1, p2 = distinct_pool(rng, 300, 20), distinct_pool(rng, 300, 20)
    template = "CAGCTACG" + "-" * 20 + "GGTACCTT" + "-" * 20 + "CGATCGAG"
    n = 2_000_000
    text = SynthSpec(template, [p1, p2], seed=6, read_len=75, strand=2).fastq(0, n)

    def table(data):
        keys, freq, total = rcpp.count_combo_barcodes_single(data, template, 2, [p1, p2], 1, True, 8)
        dense = np.zeros((len(p1), len(p2)), dtype=np.int64)
        dense[keys[0], keys[1]] = freq
        assert np.all(np.diff(keys[0] * len(p2) + keys[1]) > 0)   # sorted, distinct (reference utils.hpp:173-198)
        return dense, int(total[0])

    whole, total = table(text)
    assert total == n and whole.sum() > 0.6 * n
    acc = np.zeros_like(whole)
    for piece in _slices(text, n, 3):
        d, _ = table(piece)
        acc += d
    assert np.array_equal(acc, whole)
    monkeypatch.setenv("SCG_HOST_PARSE", "1")
    assert np.array_equal(table(text)[0], whole)


def test_random_barcode_tables_merge_like_counts(monkeypatch):
    template = FLANK_L + "-" * 16 + FLANK_R
    n = 2_000_000
    text = SynthSpec(template, [], seed=7, read_len=75, strand=2, random_space=300_000).fastq(0, n)

    def table(data):
        (seqs, freq), total = rcpp.count_random_barcodes(data, template, 2, 1, True, 8, as_array=True)
        assert np.all(seqs[:-1] < seqs[1:])   # sorted as text, distinct
        return dict(zip(seqs.tolist(), freq.tolist())), total

    whole, total = table(text)
    assert total == n and sum(whole.values()) > 0.7 * n
    acc = {}
    for piece in _slices(text, n, 4):
        d, _ = table(piece)
        for k, v in d.items():
            acc[k] = acc.get(k, 0) + v
    assert acc == whole
    monkeypatch.setenv("SCG_HOST_PARSE", "1")
    assert table(text)[0] == whole
