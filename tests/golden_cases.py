"""Known-answer vectors transcribed from the reference's own testthat suite
(/root/reference/tests/testthat/*.R) plus the quirk vectors
established in the survey (8.1 T8 "Quirk A", T9 "Quirk B").

Each case is a function taking an *engine*: any object exposing the reference's
Rcpp-level entry points (src/RcppExports.cpp:136-145) as

    count_single(fastq, template, strand, pool, mismatches, use_first)
    count_random(fastq, template, strand, mismatches, use_first)
    count_combo_single(fastq, template, strand, pool1, pool2, mismatches, use_first)
    count_dual_single_end(fastq, template, pools, strand, mismatches, use_first, diagnostics=False)
    count_dual(fastq1, template1, reverse1, mismatches1, pool1, fastq2, ..., randomized, use_first, diagnostics=False)
    count_combo_paired(fastq1, template1, reverse1, mismatches1, pool1, fastq2, ..., randomized, use_first)
    match_barcodes(seqs, choices, substitutions, reverse)    (0-based index, -1 = NA)

so the same vectors pin the C restatement, the compiled reference and the CUDA path.
The R wrappers' argument munging (template N -> '-', strand names -> 0/1/2,
use_first = !find.best; R/countSingleBarcodes.R:93-100) is applied here.
"""
import numpy as np

from util import fastq, BASES

STRAND = {"original": 0, "reverse": 1, "both": 2}


def _tmpl(t):
    return t.replace("N", "-")


def _eq(a, b):
    assert np.array_equal(np.asarray(a), np.asarray(b)), "%r != %r" % (np.asarray(a).tolist(), np.asarray(b).tolist())


# -- test-single.R:55-70 ----------------------------------------------------
def single_substitutions(e):
    template = _tmpl("ACGT" + "N" * 10 + "TGCA")
    reads = ["ACGTGGGGGGGGGGTGCA", "ACGTGGGGCGGGGGTGCA", "ACGTGGGGCCGGGGTGCA", "ACGTGGGGGGGGGGTTCA", "CCGTGGGGGGGGGGTGCA"]
    choices = [b * 10 for b in BASES]
    for use_first in (True, False):
        counts, total = e.count_single(fastq(reads), template, STRAND["both"], choices, 1, use_first)
        _eq(counts, [0, 0, 4, 0])
        assert total == 5


# -- test-single.R:72-88 ----------------------------------------------------
def single_conflicts(e):
    template = _tmpl("ACGT" + "N" * 10 + "TGCA")
    reads = ["ACGTCCCCCCCCCCTGCA", "ACGTCCCCCCCCCATGCA", "ACGTCCCCCCCCCGTGCA", "ACGTCCCCCCCCCTTGCA"]
    choices = ["CCCCCCCCCC", "CCCCCCCCCA"]
    counts, total = e.count_single(fastq(reads), template, STRAND["both"], choices, 1, True)
    _eq(counts, [1, 1])
    assert total == 4


# -- test-single.R:126-147 --------------------------------------------------
def single_iupac(e):
    template = _tmpl("ACGT" + "N" * 10 + "TGCA")
    choices = ["AAAAABAAAA", "CCCCCDCCCC", "GGGGGHGGGG", "TTTTTVTTTT"]
    pool = ["AAAAACAAAA", "AAAAAGAAAA", "AAAAAAAAAA", "CCCCCACCCC", "CCCCAACCCC", "GGGGGTGGGG", "TTTTTGTTTT"]
    reads = ["ACGT" + p + "TGCA" for p in pool]
    counts, _ = e.count_single(fastq(reads), template, STRAND["original"], choices, 0, True)
    _eq(counts, [2, 1, 1, 1])
    counts, _ = e.count_single(fastq(reads), template, STRAND["original"], choices, 1, True)
    _eq(counts, [3, 2, 1, 1])


# -- test-matchBarcodes.R:4-22 ----------------------------------------------
def match_simple(e):
    choices = ["AAAAAA", "CCCCCC", "GGGGGG", "TTTTTT"]
    q = ["AAAAAA", "AAATAA"]
    idx, mm = e.match_barcodes(q, choices, 0, False)
    _eq(idx, [0, -1]); _eq(mm, [0, -1])
    idx, mm = e.match_barcodes(q, choices, 1, False)
    _eq(idx, [0, 0]); _eq(mm, [0, 1])
    idx, mm = e.match_barcodes(q, choices, 0, True)
    _eq(idx, [3, -1]); _eq(mm, [0, -1])
    idx, mm = e.match_barcodes(q, choices, 1, True)
    _eq(idx, [3, 3]); _eq(mm, [0, 1])


# -- test-matchBarcodes.R:24-38 ---------------------------------------------
def match_iupac(e):
    choices = ["AAARAA", "CCCYCC", "GGGMGG", "TTTSTT"]
    idx, mm = e.match_barcodes(["AAAAAA", "AAAGAA"], choices, 0, False)
    _eq(idx, [0, 0]); _eq(mm, [0, 0])
    idx, mm = e.match_barcodes(["AAAAAA", "AAAGAA", "AAGAAA"], choices, 0, True)
    _eq(idx, [-1, -1, 3]); _eq(mm, [-1, -1, 0])
    idx, mm = e.match_barcodes(["AAAAAA", "AAAGAA", "AAGAAA"], choices, 2, True)
    _eq(idx, [3, 3, 3]); _eq(mm, [1, 2, 0])


# -- test-random.R:51-70 ----------------------------------------------------
def random_substitutions(e):
    template = _tmpl("AAAAACGT" + "N" * 6 + "ACGTGGGG")
    reads = ["AAAAACGTGGGGGGACGTGGGG", "AAATACGTGGGGGCACGTGGGG", "AAATACGTGGGGCCACGTGGGC"]
    seqs, freq, total = e.count_random(fastq(reads), template, STRAND["both"], 1, True)
    assert list(seqs) == ["GGGGGC", "GGGGGG"]
    _eq(freq, [1, 1])
    assert total == 3
    seqs, freq, total = e.count_random(fastq(reads), template, STRAND["both"], 2, True)
    assert list(seqs) == ["GGGGCC", "GGGGGC", "GGGGGG"]
    _eq(freq, [1, 1, 1])


# -- test-dual.R:46-93 ------------------------------------------------------
def dual_edits(e):
    template1 = _tmpl("ACGT" + "N" * 10 + "TGCA")
    b1 = ["ACGTGGGGGGGGGGTGCA", "ACGTGGGGCGGGGGTGCA", "ACGTGGGGGGGGGTGCA", "ACGTGGGGGGGGGGGTGCA"]
    choices = [b * 10 for b in BASES]
    f1 = fastq(b1)

    def run(fa, fb, s1, s2):
        counts, total = e.count_dual(fa, template1, False, s1, choices, fb, template1, False, s2, choices, False, True)
        return counts, total

    counts, total = run(f1, f1, 0, 0)
    assert int(np.sum(counts)) == 1 and total == 4
    ref, _ = e.count_single(f1, template1, STRAND["original"], choices, 0, True)
    _eq(ref, counts)

    counts, _ = run(f1, f1, 1, 1)
    assert int(np.sum(counts)) == 2
    ref, _ = e.count_single(f1, template1, STRAND["original"], choices, 1, True)
    _eq(ref, counts)

    b2 = ["ACGTGGGGCGGGGGTGCA", "ACGTGGGGGGGGGGTGCA", "ACGTGGGGGGGGGGGTGCA", "ACGTGGGGGGGGGTGCA"]
    f2 = fastq(b2)
    assert int(np.sum(run(f1, f2, 0, 0)[0])) == 0
    assert int(np.sum(run(f1, f2, 0, 1)[0])) == 1
    assert int(np.sum(run(f1, f2, 1, 0)[0])) == 1
    assert int(np.sum(run(f1, f2, 1, 1)[0])) == 2


# -- test-dual.R:173-218 ----------------------------------------------------
def dual_randomization_edges(e):
    b1 = ["AAAAAAAAA", "AAAAAAAAA", "AAAAAACAA", "AAAAAACAA"]
    b2 = ["AAAAAAAAA", "AAAAAACAA", "AAAAAAAAA", "AAAAAACAA"]
    f1, f2 = fastq(b1), fastq(b2)
    template = "---------"
    one = ["AAAAAAAAA"]

    def run(p1, p2, s1, s2, randomized, use_first=True):
        counts, _ = e.count_dual(f1, template, False, s1, p1, f2, template, False, s2, p2, randomized, use_first)
        return counts

    _eq(run(one, one, 0, 0, False), [1])
    _eq(run(one, one, 0, 0, True), [1])
    _eq(run(one, one, 1, 0, False), [2])
    _eq(run(one, one, 0, 1, False), [2])
    _eq(run(one, one, 0, 1, True), [3])
    _eq(run(one, one, 1, 1, False), run(one, one, 1, 1, True))
    out7 = run(["AAAAAAAAA", "AAAAAACAA"], ["AAAAAAAAA", "AAAAAAAAA"], 1, 1, True)
    _eq(out7, [2, 2])


# -- test-countDualBarcodesSingleEnd.R:38-59 ----------------------------------
def dual_single_end_edits(e):
    template = _tmpl("ACGT" + "N" * 10 + "TGCAAGGA" + "N" * 15 + "AGGA")
    reads = [
        "ACGTGGGGGGGGGGTGCAAGGAAAAAAAAAAAAAAAAAGGA",
        "ACGTGGGGGGGGGGTGCAAGGAAAAAAAAAAATAAAAAGGA",
        "ACGTGGGGCGGGGGTGCAAGGAAAAAAAAAAATAAAAAGGA",
    ]
    choices1 = [b * 10 for b in BASES]
    choices2 = ["A" * 15] * 4
    f = fastq(reads)
    # R default strand for countDualBarcodesSingleEnd is "both" (R/countDualBarcodesSingleEnd.R)
    for subs, expect in ((0, [0, 0, 1, 0]), (1, [0, 0, 2, 0]), (2, [0, 0, 3, 0])):
        counts, total = e.count_dual_single_end(f, template, [choices1, choices2], STRAND["both"], subs, True)
        _eq(counts, expect)
        assert total == 3


# -- segmented-search phantom ("Quirk A"), worked example ---------
def dual_quirk_a(e):
    # library {AAAA|CCCC, AAAT|CCCG}, query AAAA|CCCG, caps [1,0]: kaori says ambiguous -> counts 0 0
    t = "----"
    f1 = fastq(["AAAA"])
    f2 = fastq(["CCCG"])
    counts, total = e.count_dual(f1, t, False, 1, ["AAAA", "AAAT"], f2, t, False, 0, ["CCCC", "CCCG"], False, True)
    _eq(counts, [0, 0])
    assert total == 1
    # control library {AAAA|GGGG, AAAT|CCCG}: no library prefix AAAA|CCC -> genuine hit survives
    counts, _ = e.count_dual(f1, t, False, 1, ["AAAA", "AAAT"], f2, t, False, 0, ["GGGG", "CCCG"], False, True)
    _eq(counts, [0, 1])


# -- random-barcode reverse coordinates ("Quirk B") ----------------
def random_quirk_b(e):
    # template AAAAACGT------ACGT (8/4 flanks).  A reverse-strand read is the reverse complement
    # of AAAAACGT CATTGA ACGT = ACGT TCAATG ACGTTTTT; kaori extracts with the FORWARD coordinates.
    template = "AAAAACGT------ACGT"
    read = "ACGTTCAATGACGTTTTT"
    seqs, freq, total = e.count_random(fastq([read]), template, STRAND["reverse"], 0, True)
    assert list(seqs) == ["ACGTCA"], seqs
    _eq(freq, [1])
    assert total == 1


ALL = [
    single_substitutions,
    single_conflicts,
    single_iupac,
    match_simple,
    match_iupac,
    random_substitutions,
    dual_edits,
    dual_randomization_edges,
    dual_single_end_edits,
    dual_quirk_a,
    random_quirk_b,
]
