"""The device-side FASTQ reader (screencounter_b200/csrc/ingest.cu) against the reference's parser:
four-line records split and packed by kernels, everything else handed to the host reader at the exact byte, with the
reference's results, error texts and line numbers (FastqReader.hpp:42-110).  Small chunk sizes (SCG_INGEST_CHUNK /
SCG_INGEST_CARRY) make texts of a few kilobytes cross many chunk boundaries."""
import numpy as np
import pytest

from engines import GpuEngine
from fastq_cases import GOOD, BAD
from util import fastq, random_seq, dense_pool, distinct_pool, adversarial_reads
from screencounter_b200 import rcpp

pytestmark = pytest.mark.gpu

TEMPLATE = "ACGTA" + "-" * 6 + "TGCAT"


@pytest.fixture(scope="module")
def gpu():
    return GpuEngine()


def _set(monkeypatch, chunk=None, carry=None, host=False):
    for name, value in (("SCG_INGEST_CHUNK", chunk), ("SCG_INGEST_CARRY", carry), ("SCG_HOST_PARSE", "1" if host else None)):
        if value is None:
            monkeypatch.delenv(name, raising=False)
        else:
            monkeypatch.setenv(name, str(value))


def _case(seed, n=4000, **kw):
    rng = np.random.default_rng(seed)
    pool = dense_pool(rng, 50, 6)
    reads = adversarial_reads(rng, n, TEMPLATE, [pool], strand="both", **kw)
    return pool, reads


@pytest.mark.parametrize("chunk,carry", [(64, 1024), (256, 512), (4096, 4096), (100000, 4096), (None, None)])
def test_many_chunks(gpu, kref, monkeypatch, chunk, carry):
    """Ragged reads (0 .. ~130 bases) across many chunk boundaries: per-read outcomes equal the reference's."""
    pool, reads = _case(7)
    data = fastq(reads)
    want = kref.trace_single(data, TEMPLATE, 2, pool, 1, False)
    _set(monkeypatch, chunk, carry)
    got = gpu.trace_single(data, TEMPLATE, 2, pool, 1, False)
    t = rcpp.timing()
    assert t["reader"].startswith("device"), t
    assert "then host" not in t["reader"], t
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert t["reads"] == len(reads)


@pytest.mark.parametrize("final_newline", [True, False])
@pytest.mark.parametrize("crlf", [False, True])
def test_line_endings(gpu, kref, monkeypatch, final_newline, crlf):
    pool, reads = _case(8, n=700)
    data = fastq(reads, crlf=crlf)
    if not final_newline:
        data = data[:-1]   # CRLF files keep the '\r' (it is a base, and a quality character) and lose only the '\n'
    want = kref.trace_single(data, TEMPLATE, 2, pool, 1, True)
    for chunk in (128, 3000, None):
        _set(monkeypatch, chunk, 2048)
        got = gpu.trace_single(data, TEMPLATE, 2, pool, 1, True)
        assert "then host" not in rcpp.timing()["reader"]
        assert np.array_equal(got[0], want[0])
        assert np.array_equal(got[1], want[1])


def test_uniform_and_ragged_lengths(gpu, kref, monkeypatch):
    """Uniform 75-base reads take the no-length-array layout, one odd read switches the batch to per-read lengths."""
    rng = np.random.default_rng(9)
    pool = distinct_pool(rng, 300, 20)
    template = "CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG"
    reads = adversarial_reads(rng, 5000, template, [pool], strand="both", read_len=75, sub_rate=0.01, n_rate=0.001, lower_rate=0.01,
                              double_frac=0.0, short_frac=0.0)
    for extra in ([], ["ACGT"], ["A" * 300]):
        data = fastq(reads[:2500] + extra + reads[2500:])
        want = kref.trace_single(data, template, 2, pool, 1, True)
        for chunk in (50000, None):
            _set(monkeypatch, chunk, 4096)
            got = gpu.trace_single(data, template, 2, pool, 1, True)
            assert np.array_equal(got[0], want[0])
            assert np.array_equal(got[1], want[1])


def _irregular(kind, seq):
    q = "I" * len(seq)
    if kind == "wrapped_seq":
        return "@w\n%s\n%s\n+\n%s\n" % (seq[:3], seq[3:], q)
    if kind == "wrapped_qual":
        return "@w\n%s\n+\n%s\n%s\n" % (seq, q[:2], q[2:])
    if kind == "plus_in_seq":   # the sequence ends at the first '+' wherever it is (FastqReader.hpp:72)
        return "@w\n%s+junk\n%s\n" % (seq, q)
    if kind == "at_in_qual":
        return "@w\n%s\n+\n@%s\n" % (seq, q[1:])
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["wrapped_seq", "wrapped_qual", "plus_in_seq", "at_in_qual"])
@pytest.mark.parametrize("where", [0, 1, 777, -1])
def test_handover_to_the_host_reader(gpu, kref, monkeypatch, kind, where):
    """One record that is not a plain four-line record, anywhere in the text: the device reader stops in front of it,
    the host reader resumes at that byte, and nothing is lost or read twice."""
    pool, reads = _case(10, n=1500, short_frac=0.0)
    recs = [fastq([r]).decode() for r in reads]
    at = where if where >= 0 else len(recs)
    recs.insert(at, _irregular(kind, reads[5]))
    data = "".join(recs).encode()
    want = kref.trace_single(data, TEMPLATE, 2, pool, 1, False)
    assert len(want[0]) == len(reads) + 1
    for chunk in (200, 5000, None):
        _set(monkeypatch, chunk, 1024)
        got = gpu.trace_single(data, TEMPLATE, 2, pool, 1, False)
        t = rcpp.timing()
        if kind != "at_in_qual":   # '@' as a quality character is still a four-line record
            assert "then host from byte %d" % len("".join(recs[:at])) in t["reader"], t
        else:
            assert "then host" not in t["reader"], t
        assert np.array_equal(got[0], want[0])
        assert np.array_equal(got[1], want[1])


@pytest.mark.parametrize("name", sorted(GOOD))
@pytest.mark.parametrize("chunk", [64, 80])
def test_grammar_cases_small_chunks(gpu, kref, monkeypatch, name, chunk):
    data = GOOD[name]
    want = kref.trace_single(data, "AC--", 2, ["GT", "AC", "TT", "GG"], 1, False)
    _set(monkeypatch, chunk, 32)   # a 32-byte carry area overflows on most of these: another way to hand over
    got = gpu.trace_single(data, "AC--", 2, ["GT", "AC", "TT", "GG"], 1, False)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])


@pytest.mark.parametrize("name", sorted(BAD))
@pytest.mark.parametrize("lead", [0, 3, 500])
def test_errors_keep_their_line_numbers(gpu, monkeypatch, name, lead):
    """A malformed record after `lead` good ones: the message names the line the reference would name."""
    data, msg = BAD[name]
    good = fastq(["ACGTAGG"] * lead)
    import re
    shifted = re.sub(r"line (\d+)", lambda m: "line %d" % (int(m.group(1)) + 4 * lead), msg)
    for chunk in (96, None):
        _set(monkeypatch, chunk, 64)
        with pytest.raises(Exception) as err:
            gpu.count_single(good + data, "AC--", 2, ["GT"], 0, True)
        assert str(err.value) == shifted


@pytest.mark.parametrize("junk", [b"\n" * 5000, b"\n\n\n\n" * 300 + b"@r\nACGTAGGGTTTTGCAT\n+\nIIIIIIIIIIIIIIII\n", b"+\n" * 999, b"@\n" * 4001])
def test_texts_that_are_mostly_newlines(gpu, kref, monkeypatch, junk):
    """Every group of four lines is validated, however short the lines: blank-line floods are the host reader's to reject,
    with the reference's message."""
    good = fastq(["ACGTAGGGTTTTGCAT"] * 50)
    for data in (junk, good + junk):
        try:
            want = kref.count_single(data, TEMPLATE, 2, ["GGGTTT"], 0, True)
            err = None
        except Exception as e:
            want, err = None, str(e)
        for chunk in (4096, None):
            _set(monkeypatch, chunk, 1024)
            if err is None:
                got = gpu.count_single(data, TEMPLATE, 2, ["GGGTTT"], 0, True)
                assert got[1] == want[1] and np.array_equal(got[0], want[0])
            else:
                with pytest.raises(Exception) as raised:
                    gpu.count_single(data, TEMPLATE, 2, ["GGGTTT"], 0, True)
                assert str(raised.value) == err


def test_record_longer_than_the_carry_area(gpu, kref, monkeypatch):
    pool, reads = _case(11, n=300)
    reads.insert(100, random_seq(np.random.default_rng(1), 5000))
    data = fastq(reads)
    want = kref.trace_single(data, TEMPLATE, 2, pool, 1, True)
    _set(monkeypatch, 2048, 512)
    got = gpu.trace_single(data, TEMPLATE, 2, pool, 1, True)
    assert "then host" in rcpp.timing()["reader"]
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])


def test_pinned_text_and_both_readers_agree(gpu, kref, monkeypatch):
    """Text in page-locked memory (scg_host_alloc) is copied by DMA straight from the caller's buffer; the host reader
    (SCG_HOST_PARSE=1) gives the same answer as both device routes."""
    pool, reads = _case(12, n=20000)
    data = fastq(reads)
    want = kref.trace_single(data, TEMPLATE, 2, pool, 1, False)
    _set(monkeypatch, 1 << 20, 1 << 16)
    pinned = rcpp.PinnedText.from_bytes(data)
    got = gpu.trace_single(pinned, TEMPLATE, 2, pool, 1, False)
    assert "page-locked" in rcpp.timing()["reader"]
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    got = gpu.trace_single(data, TEMPLATE, 2, pool, 1, False)
    assert "bounce" in rcpp.timing()["reader"]
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    _set(monkeypatch, host=True)
    got = gpu.trace_single(data, TEMPLATE, 2, pool, 1, False)
    assert rcpp.timing()["reader"] == "host"
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    pinned.free()


def test_files_take_the_device_reader(gpu, kref, monkeypatch, tmp_path):
    pool, reads = _case(13, n=30000)
    data = fastq(reads)
    path = tmp_path / "reads.fastq"
    path.write_bytes(data)
    want = kref.count_single(data, TEMPLATE, 2, pool, 1, True)
    _set(monkeypatch, 1 << 18, 1 << 12)
    got = gpu.count_single(str(path), TEMPLATE, 2, pool, 1, True, nthreads=4)
    assert rcpp.timing()["reader"].startswith("device")
    assert got[1] == want[1] and np.array_equal(got[0], want[0])


def test_block_gzip_file(gpu, kref, monkeypatch, tmp_path):
    """A BGZF file crosses PCIe compressed and is inflated on the device (tests/test_gpu_bgzf.py has the details); a plain gzip
    file is inflated by zlib on the host, member by member, and parsed there."""
    import gzip
    from util import bgzf
    pool, reads = _case(16, n=30000)
    data = fastq(reads)
    want = kref.count_single(data, TEMPLATE, 2, pool, 1, True)
    _set(monkeypatch)
    for name, blob in (("b.fastq.gz", bgzf(data, 20000)), ("p.fastq.gz", gzip.compress(data, 1))):
        path = tmp_path / name
        path.write_bytes(blob)
        got = gpu.count_single(str(path), TEMPLATE, 2, pool, 1, True, nthreads=6)
        reader = rcpp.timing()["reader"]
        assert reader == "host" if name.startswith("p.") else ("block-gzip" in reader and "then host" not in reader), reader
        assert got[1] == want[1] and np.array_equal(got[0], want[0])


def test_other_single_end_entry_points(gpu, kref, monkeypatch):
    """Combinatorial and dual single-end counting read through the same pipeline."""
    rng = np.random.default_rng(14)
    p1, p2 = distinct_pool(rng, 30, 5), distinct_pool(rng, 30, 5)
    template = "ACGT" + "-" * 5 + "GG" + "-" * 5 + "TGCA"
    reads = adversarial_reads(rng, 5000, template, [p1, p2], strand="both")
    data = fastq(reads)
    _set(monkeypatch, 3000, 1024)
    want = kref.count_combo_single(data, template, 2, p1, p2, 1, True)
    got = gpu.count_combo_single(data, template, 2, p1, p2, 1, True)
    assert rcpp.timing()["reader"].startswith("device")
    assert got[2] == want[2] and np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    want = kref.count_dual_single_end(data, template, [p1, p2], 2, 1, True)
    got = gpu.count_dual_single_end(data, template, [p1, p2], 2, 1, True)
    assert got[1] == want[1] and np.array_equal(got[0], want[0])


def test_two_full_size_chunks(gpu, kref, monkeypatch):
    """Default chunk size (32 MiB): 330k reads of 75 bases cross one real chunk boundary."""
    rng = np.random.default_rng(15)
    pool = distinct_pool(rng, 1000, 20)
    template = "CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG"
    base = adversarial_reads(rng, 3000, template, [pool], strand="both", read_len=75, sub_rate=0.01, n_rate=0.001, lower_rate=0.0,
                             double_frac=0.0, short_frac=0.0)
    reads = base * 110
    data = fastq(reads)
    assert len(data) > (32 << 20)
    _set(monkeypatch)
    want = kref.count_single(data, template, 2, pool, 1, True)
    got = gpu.count_single(data, template, 2, pool, 1, True, nthreads=8)
    t = rcpp.timing()
    assert t["reader"].startswith("device") and "then host" not in t["reader"]
    assert got[1] == want[1] == len(reads)
    assert np.array_equal(got[0], want[0])


# ---- paired input: two device readers that must cut their chunks at the same record ----------------------------------
def _paired_case(seed, n, len2_extra=0, dense=True):
    rng = np.random.default_rng(seed)
    p1, p2 = (dense_pool(rng, 40, 6), dense_pool(rng, 40, 6)) if dense else (distinct_pool(rng, 40, 6), distinct_pool(rng, 40, 6))
    t1, t2 = "ACGTA" + "-" * 6 + "TGCAT", "GGATC" + "-" * 6 + "CCTAG"
    r1 = adversarial_reads(rng, n, t1, [p1], strand="original", short_frac=0.01)
    r2 = adversarial_reads(rng, n, t2, [p2], strand="original", short_frac=0.01)
    if len2_extra:
        r2 = [r + "A" * len2_extra for r in r2]
    return (t1, p1, r1), (t2, p2, r2)


@pytest.mark.parametrize("chunk,carry", [(300, 2048), (5000, 4096), (None, None)])
@pytest.mark.parametrize("use_first", [True, False])
def test_paired_chunks_stay_aligned(gpu, kref, monkeypatch, chunk, carry, use_first):
    (t1, p1, r1), (t2, p2, r2) = _paired_case(21, 3000)
    f1, f2 = fastq(r1), fastq(r2, names=["pair%d/2" % i for i in range(len(r2))])   # records of different sizes
    want = kref.trace_dual(f1, t1, False, 1, p1, f2, t2, False, 1, p2, True, use_first)
    _set(monkeypatch, chunk, carry)
    got = gpu.trace_dual(f1, t1, False, 1, p1, f2, t2, False, 1, p2, True, use_first)
    assert rcpp.timing()["reader"].startswith("device")
    assert np.array_equal(np.asarray(got).ravel(), np.asarray(want).ravel())
    # counts against the per-pair trace (the reference's own counts can depend on read order through its result cache,
    # SURVEY 8.1 T20; the trace above is its fresh-state run)
    gc = gpu.count_dual(f1, t1, False, 1, p1, f2, t2, False, 1, p2, True, use_first)
    w = np.asarray(want).ravel()
    assert gc[1] == len(r1) and np.array_equal(gc[0], np.bincount(w[w >= 0], minlength=len(p1)))


def test_paired_mates_of_very_different_size(gpu, kref, monkeypatch):
    """Mate 2 records are much longer: its reader falls behind, the carry area runs over, the host readers finish."""
    (t1, p1, r1), (t2, p2, r2) = _paired_case(22, 2500, len2_extra=150)
    f1, f2 = fastq(r1), fastq(r2)
    want = kref.count_combo_paired(f1, t1, False, 1, p1, f2, t2, False, 1, p2, False, True)
    for chunk, carry in ((2000, 512), (None, None)):
        _set(monkeypatch, chunk, carry)
        got = gpu.count_combo_paired(f1, t1, False, 1, p1, f2, t2, False, 1, p2, False, True)
        assert got[2] == want[2] == len(r1)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        assert got[3] == want[3] and got[4] == want[4]


@pytest.mark.parametrize("which", [1, 2])
def test_paired_files_of_different_length(gpu, monkeypatch, which):
    (t1, p1, r1), (t2, p2, r2) = _paired_case(23, 1200)
    if which == 1:
        r1 = r1 + ["ACGTAGGGTTTTGCAT"]
    else:
        r2 = r2 + ["GGATCAAAAAACCTAG", "GGATCAAAAAACCTAG"]
    f1, f2 = fastq(r1), fastq(r2)
    for chunk in (700, None):
        _set(monkeypatch, chunk, 1024)
        with pytest.raises(Exception, match="different number of reads in paired FASTQ files"):
            gpu.count_dual(f1, t1, False, 1, p1, f2, t2, False, 1, p2, False, True)


def test_paired_irregular_record_in_one_mate(gpu, kref, monkeypatch):
    (t1, p1, r1), (t2, p2, r2) = _paired_case(24, 2000, dense=False)
    recs2 = [fastq([r]).decode() for r in r2]
    recs2[1234] = _irregular("wrapped_seq", r2[1234] if len(r2[1234]) > 4 else "ACGTACGT")
    f1, f2 = fastq(r1), "".join(recs2).encode()
    want = kref.count_dual(f1, t1, False, 0, p1, f2, t2, False, 0, p2, False, False)
    for chunk in (900, None):
        _set(monkeypatch, chunk, 1024)
        got = gpu.count_dual(f1, t1, False, 0, p1, f2, t2, False, 0, p2, False, False)
        assert "then host" in rcpp.timing()["reader"]
        assert got[1] == want[1] and np.array_equal(got[0], want[0])


# ---- random barcodes: the device reader flags reads whose raw characters the host has to render -------------------------
@pytest.mark.parametrize("chunk", [200, 4096, None])
def test_random_barcodes_through_the_device_reader(gpu, kref, monkeypatch, chunk):
    rng = np.random.default_rng(25)
    template = "ACGTACGT" + "-" * 8 + "TTGCAGCA"
    reads = []
    for _ in range(3000):
        core = "ACGTACGT" + random_seq(rng, 8) + "TTGCAGCA"
        u = rng.random()
        if u < 0.1:
            core = core.lower()
        elif u < 0.2:
            core = core[:10] + "n" + core[11:]
        elif u < 0.3:
            core = core[:9] + "N" + core[10:]
        reads.append(random_seq(rng, int(rng.integers(0, 9))) + core + random_seq(rng, int(rng.integers(0, 9))))
    data = fastq(reads)
    want = kref.count_random(data, template, 0, 1, True)
    _set(monkeypatch, chunk, 1024)
    got = gpu.count_random(data, template, 0, 1, True)
    assert rcpp.timing()["reader"].startswith("device")
    order = sorted(range(len(want[0])), key=lambda i: want[0][i])
    assert list(got[0]) == [want[0][i] for i in order]
    assert np.array_equal(got[1], np.asarray(want[1])[order])
    assert got[2] == want[2]
