"""Many files behind one call, and several devices behind one context (csrc/runners_multi.cu), against the compiled
reference run file by file: the count matrix is the column-bind of the per-file results (R/countSingleBarcodes.R:112-126),
the sparse designs give the sorted union of the files' keys with one column per file (R/combineComboCounts.R:31-57,
R/countRandomBarcodes.R:84-105).  A context may list one device several times: the parts of a cut-up file then run on
several streams of that device, which exercises the cutting and the merge on a one-GPU box; with two GPUs visible the
same tests also run over devices (0, 1)."""
import os

import numpy as np
import pytest

from util import fastq, distinct_pool, adversarial_reads

pytestmark = pytest.mark.gpu

TEMPLATE = "CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG"
COMBO_TEMPLATE = "CAGCTACG" + "-" * 12 + "GGTACCTT" + "-" * 12 + "CGATCGAG"
RANDOM_TEMPLATE = "CAGCTACGTACG" + "-" * 10 + "CCAGCTCGATCG"


def _oracle():
    from oracle import port, kref
    return kref if kref.available() else port


def _device_sets():
    from screencounter_b200 import rcpp  # noqa: F401
    import ctypes as C
    sets = [0, (0, 0), (0, 0, 0)]
    try:
        import torch
        if torch.cuda.device_count() >= 2:
            sets.append((0, 1))
    except Exception:
        pass
    return sets


def _files(rng, template, pools, sizes, **kw):
    return [fastq(adversarial_reads(rng, n, template, pools, strand="both", **kw)) if n else b"" for n in sizes]


@pytest.mark.parametrize("devices", _device_sets(), ids=str)
def test_matrix_of_single_barcodes(devices):
    from screencounter_b200 import rcpp
    rng = np.random.default_rng(5)
    pool = distinct_pool(rng, 300, 20)
    files = _files(rng, TEMPLATE, [pool], [3000, 1, 0, 7000, 2500])
    matrix, totals = rcpp.matrix_of_single_barcodes(files, TEMPLATE, 2, pool, 1, True, 2, device=devices)
    assert matrix.shape == (len(pool), len(files))
    for f, text in enumerate(files):
        counts, total = _oracle().count_single(text, TEMPLATE, 2, pool, 1, True)
        assert total == totals[f]
        assert np.array_equal(matrix[:, f], counts), "column %d differs from the reference's counts of that file" % f


@pytest.mark.parametrize("devices", _device_sets()[1:], ids=str)
def test_one_file_cut_over_the_devices(devices, monkeypatch):
    """scg_count_single on a context of several devices: the text is cut at record boundaries, counts and the per-read trace
    are those of the whole file."""
    from screencounter_b200 import rcpp
    monkeypatch.setenv("SCG_MULTI_MIN_BYTES", "1000")
    rng = np.random.default_rng(6)
    pool = distinct_pool(rng, 200, 20)
    reads = adversarial_reads(rng, 20000, TEMPLATE, [pool], strand="both")
    text = fastq(reads)
    counts, total, (index, info) = rcpp.count_single_barcodes(text, TEMPLATE, 2, pool, 1, False, 2, trace=True, device=devices)
    want_index, want_info = _oracle().trace_single(text, TEMPLATE, 2, pool, 1, False)
    assert total == len(reads)
    assert np.array_equal(index, want_index) and np.array_equal(info, want_info)
    assert np.array_equal(counts, np.bincount(want_index[want_index >= 0], minlength=len(pool)))
    assert "devices" in rcpp.timing(devices)["kernel"]
    # a text that cannot be cut cleanly (a wrapped record in the middle) is read by one device, with the same answer
    half = len(reads) // 2
    wrapped = ("@wrapped\n" + reads[half][:10] + "\n" + reads[half][10:] + "\n+\n" + "I" * len(reads[half]) + "\n").encode()
    odd = fastq(reads[:half]) + wrapped + fastq(reads[half + 1:])
    c2, t2 = rcpp.count_single_barcodes(odd, TEMPLATE, 2, pool, 1, False, 2, device=devices)
    assert t2 == len(reads) and np.array_equal(c2, counts)
    # a malformed record raises the reference's error, line number included, whichever part it sits in
    bad = fastq(reads[:half]) + b"@broken\nACGT\n+\nII\n" + fastq(reads[half:])
    with pytest.raises(rcpp.ScreenCounterError) as e:
        rcpp.count_single_barcodes(bad, TEMPLATE, 2, pool, 1, False, 2, device=devices)
    try:
        _oracle().count_single(bad, TEMPLATE, 2, pool, 1, False)
        raise AssertionError("the reference accepted the malformed file")
    except Exception as ref_error:
        assert str(ref_error) in str(e.value) or str(e.value) in str(ref_error)


@pytest.mark.parametrize("devices", _device_sets(), ids=str)
@pytest.mark.parametrize("npool", [40, 4500])   # dense matrix / device hash
def test_matrix_of_combo_barcodes(devices, npool):
    from screencounter_b200 import rcpp
    rng = np.random.default_rng(7)
    p1, p2 = distinct_pool(rng, npool, 12), distinct_pool(rng, npool, 12)
    files = _files(rng, COMBO_TEMPLATE, [p1[:60], p2[:60]], [2500, 0, 4000, 1500])
    keys, counts, totals = rcpp.matrix_of_combo_barcodes(files, COMBO_TEMPLATE, 2, [p1, p2], 1, True, 2, device=devices)
    want = {}
    for f, text in enumerate(files):
        k, freq, total = _oracle().count_combo_single(text, COMBO_TEMPLATE, 2, p1, p2, 1, True)
        assert total == totals[f]
        for row, n in zip(np.asarray(k).reshape(-1, 2), freq):
            want.setdefault((int(row[0]), int(row[1])), [0] * len(files))[f] = int(n)
    rows = sorted(want)
    assert keys.shape == (2, len(rows))
    assert [tuple(int(v) for v in keys[:, i]) for i in range(len(rows))] == rows
    assert np.array_equal(counts, np.array([want[r] for r in rows], dtype=np.int32).reshape(len(rows), len(files)))


@pytest.mark.parametrize("devices", _device_sets(), ids=str)
@pytest.mark.parametrize("lower_rate", [0.0, 0.02])   # device union / host union (keys rendered from raw read text)
def test_matrix_of_random_barcodes(devices, lower_rate):
    from screencounter_b200 import rcpp
    rng = np.random.default_rng(9)
    barcodes = distinct_pool(rng, 150, 10)
    files = _files(rng, RANDOM_TEMPLATE, [barcodes], [3000, 2000, 0, 1], lower_rate=lower_rate)
    seqs, counts, totals = rcpp.matrix_of_random_barcodes(files, RANDOM_TEMPLATE, 2, 1, True, 2, device=devices, as_array=False)
    want = {}
    for f, text in enumerate(files):
        s, freq, total = _oracle().count_random(text, RANDOM_TEMPLATE, 2, 1, True)
        assert total == totals[f]
        for key, n in zip(s, freq):
            want.setdefault(key, [0] * len(files))[f] = int(n)
    rows = sorted(want)
    assert list(seqs) == rows
    assert np.array_equal(counts, np.array([want[r] for r in rows], dtype=np.int32).reshape(len(rows), len(files)))


def test_files_of_every_format_in_one_call(tmp_path):
    """A list that mixes a raw file, a block-gzip file, a plain gzip file, text in memory and a block-gzip image in memory:
    every column equals the reference's counts of that input's text (the reference sniffs gzip like we do,
    inst/include/byteme/SomeFileReader.hpp:25-66)."""
    import gzip
    from screencounter_b200 import rcpp
    from util import bgzf
    rng = np.random.default_rng(12)
    pool = distinct_pool(rng, 120, 20)
    texts = _files(rng, TEMPLATE, [pool], [4000, 3000, 2000, 1500, 2500])
    raw, blk, gz = tmp_path / "a.fastq", tmp_path / "b.fastq.gz", tmp_path / "c.fastq.gz"
    raw.write_bytes(texts[0])
    blk.write_bytes(bgzf(texts[1], 9000))
    gz.write_bytes(gzip.compress(texts[2], 3))
    inputs = [str(raw), str(blk), str(gz), texts[3], bgzf(texts[4], 30000)]
    for devices in _device_sets():
        matrix, totals = rcpp.matrix_of_single_barcodes(inputs, TEMPLATE, 2, pool, 1, False, 3, device=devices)
        for f, text in enumerate(texts):
            counts, total = _oracle().count_single(text, TEMPLATE, 2, pool, 1, False)
            assert total == totals[f] and np.array_equal(matrix[:, f], counts), (devices, f)


@pytest.mark.parametrize("devices", _device_sets()[1:], ids=str)
@pytest.mark.parametrize("block", [1500, 65280])
def test_one_block_gzip_file_cut_over_the_devices(devices, block, monkeypatch, tmp_path):
    """The same for a block-gzip input: the cuts are made in the coordinates of the TEXT (a part begins inside some member, at a
    record start found on that member's text), every device inflates and reads the members that hold its part."""
    from screencounter_b200 import rcpp
    from util import bgzf
    monkeypatch.setenv("SCG_MULTI_MIN_BYTES", "1000")
    rng = np.random.default_rng(14)
    pool = distinct_pool(rng, 200, 20)
    reads = adversarial_reads(rng, 15000, TEMPLATE, [pool], strand="both")
    text = fastq(reads)
    want_index, want_info = _oracle().trace_single(text, TEMPLATE, 2, pool, 1, False)
    image = bgzf(text, block)
    path = tmp_path / "reads.fastq.gz"
    path.write_bytes(image)
    for src in (image, str(path)):
        counts, total, (index, info) = rcpp.count_single_barcodes(src, TEMPLATE, 2, pool, 1, False, 2, trace=True, device=devices)
        t = rcpp.timing(devices)
        assert "devices" in t["kernel"] and "inflated on the device" in t["reader"], t
        assert total == len(reads)
        assert np.array_equal(index, want_index) and np.array_equal(info, want_info)
        assert np.array_equal(counts, np.bincount(want_index[want_index >= 0], minlength=len(pool)))
    # a wrapped record in the middle: some part does not parse cleanly, one device reads the whole file, same answer
    half = len(reads) // 2
    wrapped = ("@wrapped\n" + reads[half][:10] + "\n" + reads[half][10:] + "\n+\n" + "I" * len(reads[half]) + "\n").encode()
    odd = fastq(reads[:half]) + wrapped + fastq(reads[half + 1:])
    c2, t2 = rcpp.count_single_barcodes(bgzf(odd, block), TEMPLATE, 2, pool, 1, False, 2, device=devices)
    assert t2 == len(reads) and np.array_equal(c2, np.bincount(want_index[want_index >= 0], minlength=len(pool)))
