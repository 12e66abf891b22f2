"""countSingleBarcodes / matchBarcodes on the GPU (through the C ABI) against the reference:
golden vectors, adversarial fuzz with per-read outcomes, FASTQ grammar, edge cases."""
import numpy as np
import pytest

import golden_cases
from engines import GpuEngine
from fastq_cases import GOOD, BAD
from util import fastq, random_seq, dense_pool, distinct_pool, adversarial_reads, mutate

pytestmark = pytest.mark.gpu

STRANDS = {"original": 0, "reverse": 1, "both": 2}
SINGLE_GOLDEN = [golden_cases.single_substitutions, golden_cases.single_conflicts, golden_cases.single_iupac,
                 golden_cases.match_simple, golden_cases.match_iupac]


@pytest.fixture(scope="module")
def gpu():
    return GpuEngine()


@pytest.mark.parametrize("case", SINGLE_GOLDEN, ids=lambda f: f.__name__)
def test_golden(gpu, case):
    case(gpu)


@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("strand", ["original", "reverse", "both"])
@pytest.mark.parametrize("mm", [0, 1, 2, 3])
@pytest.mark.parametrize("use_first", [True, False])
def test_fuzz_against_reference(gpu, kref, seed, strand, mm, use_first):
    rng = np.random.default_rng(1000 * seed + 10 * mm + STRANDS[strand])
    L = int(rng.integers(5, 9))
    pool = dense_pool(rng, 60, L)
    template = "ACGTA" + "-" * L + ("TGCAT" if seed == 1 else "GG")
    reads = adversarial_reads(rng, 3000, template, [pool], strand=strand)
    f = fastq(reads)
    want_counts, want_total = kref.count_single(f, template, STRANDS[strand], pool, mm, use_first)
    got_counts, got_total = gpu.count_single(f, template, STRANDS[strand], pool, mm, use_first)
    assert got_total == want_total
    assert np.array_equal(got_counts, want_counts)
    want_index, want_info = kref.trace_single(f, template, STRANDS[strand], pool, mm, use_first)
    got_index, got_info = gpu.trace_single(f, template, STRANDS[strand], pool, mm, use_first)
    assert np.array_equal(got_index, want_index)
    assert np.array_equal(got_info, want_info)


@pytest.mark.parametrize("mm", [0, 1, 2])
@pytest.mark.parametrize("use_first", [True, False])
def test_config_shapes(gpu, kref, mm, use_first):
    """20-bp guides in a 12+20+12 template inside 75-bp reads (the BASELINE configs' shape), 1 % noise."""
    rng = np.random.default_rng(42 + mm)
    pool = distinct_pool(rng, 2000, 20)
    template = "CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG"
    reads = adversarial_reads(rng, 20000, template, [pool], strand="both", read_len=75, sub_rate=0.01, n_rate=0.001,
                              lower_rate=0.0, double_frac=0.0, short_frac=0.0)
    f = fastq(reads)
    want_index, want_info = kref.trace_single(f, template, 2, pool, mm, use_first)
    got_index, got_info = gpu.trace_single(f, template, 2, pool, mm, use_first)
    assert np.array_equal(got_index, want_index)
    assert np.array_equal(got_info, want_info)
    assert (want_index >= 0).mean() > 0.5


@pytest.mark.parametrize("tlen", [31, 32, 33, 63, 64, 65, 100, 128, 200, 256])
def test_template_lengths(gpu, kref, tlen):
    """Templates across the reference's 32/64/128/256 dispatch boundaries, long variable regions and long reads."""
    rng = np.random.default_rng(tlen)
    L = max(4, tlen // 3)
    left = (tlen - L) // 2
    template = random_seq(rng, left) + "-" * L + random_seq(rng, tlen - L - left)
    pool = distinct_pool(rng, 50, L)
    reads = adversarial_reads(rng, 1500, template, [pool], strand="both", sub_rate=0.01)
    f = fastq(reads)
    for mm in (0, 2):
        want_index, want_info = kref.trace_single(f, template, 2, pool, mm, False)
        got_index, got_info = gpu.trace_single(f, template, 2, pool, mm, False)
        assert np.array_equal(got_index, want_index)
        assert np.array_equal(got_info, want_info)


def test_template_too_long(gpu):
    with pytest.raises(Exception, match="lacking compile-time support for constant regions longer than 256 bp"):
        gpu.count_single(fastq(["ACGT"]), "A" * 250 + "-" * 7, 0, ["ACGTACG"], 0, True)


def test_large_budget(gpu, kref):
    """A mismatch budget above the number of constant bases (everything matches the flanks)."""
    rng = np.random.default_rng(9)
    pool = distinct_pool(rng, 30, 6)
    template = "AC" + "-" * 6 + "GT"
    reads = adversarial_reads(rng, 800, template, [pool], strand="both")
    f = fastq(reads)
    for mm in (4, 6, 12):
        for use_first in (True, False):
            want = kref.trace_single(f, template, 2, pool, mm, use_first)
            got = gpu.trace_single(f, template, 2, pool, mm, use_first)
            assert np.array_equal(got[0], want[0])


@pytest.mark.parametrize("name", sorted(GOOD))
def test_fastq_grammar(gpu, kref, name):
    data = GOOD[name]
    template = "AC--"
    pool = ["GT", "AC", "TT", "GG"]
    want = kref.trace_single(data, template, 2, pool, 1, False)
    got = gpu.trace_single(data, template, 2, pool, 1, False)
    assert np.array_equal(got[0], want[0])
    assert gpu.count_single(data, template, 2, pool, 1, False)[1] == kref.count_single(data, template, 2, pool, 1, False)[1]


@pytest.mark.parametrize("name", sorted(BAD))
def test_fastq_errors(gpu, name):
    data, msg = BAD[name]
    with pytest.raises(Exception) as err:
        gpu.count_single(data, "AC--", 2, ["GT"], 0, True)
    assert str(err.value) == msg


def test_files_and_threads(gpu, kref, tmp_path):
    import gzip
    rng = np.random.default_rng(3)
    pool = distinct_pool(rng, 100, 10)
    template = "ACGT" + "-" * 10 + "TGCA"
    reads = adversarial_reads(rng, 50000, template, [pool], strand="both")
    data = fastq(reads)
    raw = tmp_path / "x.fastq"
    raw.write_bytes(data)
    gz = tmp_path / "x.fastq.gz"
    with gzip.open(gz, "wb") as f:
        f.write(data)
    want = kref.count_single(data, template, 2, pool, 1, True)
    for src in (data, str(raw), str(gz)):
        for nthreads in (1, 4):
            got = gpu.count_single(src, template, 2, pool, 1, True, nthreads)
            assert got[1] == want[1]
            assert np.array_equal(got[0], want[0])


def test_match_barcodes_fuzz(gpu, kref):
    rng = np.random.default_rng(5)
    IUPAC = "ACGTRYSWKMBDHVN"
    for trial in range(10):
        L = int(rng.integers(4, 40))
        lib = []
        while len(lib) < 60:
            s = random_seq(rng, L)
            if rng.random() < 0.2:
                pos = int(rng.integers(0, L))
                s = s[:pos] + IUPAC[int(rng.integers(4, len(IUPAC)))] + s[pos + 1:]
            lib.append(s)
        queries = [mutate(rng, random_seq(rng, L) if rng.random() < 0.2 else
                          lib[int(rng.integers(0, len(lib)))].translate(str.maketrans("RYSWKMBDHVN", "ACCAGACAAAA")), 0.1, 0.03, 0.03)
                   for _ in range(500)]
        for subs in (0, 1, 2, 3):
            for reverse in (False, True):
                try:
                    want = kref.match_barcodes(queries, lib, subs, reverse)
                except Exception as err:
                    with pytest.raises(Exception) as got:
                        gpu.match_barcodes(queries, lib, subs, reverse)
                    assert str(got.value) == str(err)
                    continue
                got = gpu.match_barcodes(queries, lib, subs, reverse)
                assert np.array_equal(got[0], want[0])
                assert np.array_equal(got[1], want[1])
