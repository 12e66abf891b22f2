"""The two-segment search behind countDualBarcodes (SegmentedBarcodeSearch<2>, MismatchTrie.hpp:513-663) on the device
against the compiled reference, query by query: every cap pair up to [3, 3], including the pairs [>= 2, 0] that only the
reference's own walk over its trie reproduces ("Quirk A"), and countDualBarcodes with two substitutions."""
import numpy as np
import pytest

from engines import GpuEngine
from util import fastq, random_seq, mutate, dense_pool, adversarial_reads
from screencounter_b200 import rcpp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("caps", [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (3, 0), (2, 1), (2, 2), (3, 1), (1, 2), (3, 3)])
@pytest.mark.parametrize("lens", [(4, 3), (6, 5), (20, 20)])
def test_segmented_search_fuzz(kref, caps, lens):
    rng = np.random.default_rng(100 * caps[0] + 10 * caps[1] + lens[0])
    L1, L2 = lens
    for trial in range(6):
        n = int(rng.integers(2, 80)) if L1 < 10 else int(rng.integers(50, 400))
        lib = list({random_seq(rng, L1 + L2) for _ in range(n)})
        # neighbours: one-base variants (also of the LAST base, where the phantom lives) of existing rows
        for _ in range(len(lib) // 2):
            base = lib[int(rng.integers(0, len(lib)))]
            pos = int(rng.integers(0, L1 + L2)) if rng.random() < 0.6 else L1 + L2 - 1
            alt = base[:pos] + "ACGT"[int(rng.integers(0, 4))] + base[pos + 1:]
            if alt not in lib:
                lib.append(alt)
        queries = [mutate(rng, lib[int(rng.integers(0, len(lib)))], 0.15 if L1 < 10 else 0.06, 0.02, 0.03) for _ in range(1500)]
        qcaps = np.stack([rng.integers(0, caps[0] + 1, size=len(queries)), rng.integers(0, caps[1] + 1, size=len(queries))], axis=1).astype(np.int32)
        want = kref.search_segmented2(queries, qcaps, lib, L1, L2, caps[0], caps[1])
        got = rcpp.search_segmented(queries, qcaps, lib, L1, L2, caps[0], caps[1])
        assert np.array_equal(np.maximum(got[0], -1), np.maximum(want[0], -1))   # -1 missing and -2 ambiguous are both "no match"
        hit = want[0] >= 0
        assert np.array_equal(got[1][hit], want[1][hit])


@pytest.mark.parametrize("subs", [(2, 0), (2, 1), (3, 0), (2, 2)])
@pytest.mark.parametrize("use_first", [True, False])
def test_dual_paired_with_two_substitutions(kref, subs, use_first):
    """countDualBarcodes with 2+ substitutions on read 1 (was refused): per-pair outcomes equal the reference's fresh-state run."""
    gpu = GpuEngine()
    rng = np.random.default_rng(7 + subs[0] + 3 * subs[1])
    p1, p2 = dense_pool(rng, 60, 7, 0.4), dense_pool(rng, 60, 7, 0.4)
    t1, t2 = "ACGTA" + "-" * 7 + "TGCAT", "GGATC" + "-" * 7 + "CCTAG"
    r1, r2 = [], []
    for _ in range(3000):   # the same row on both mates, noisy, at a random offset
        i = int(rng.integers(0, len(p1)))
        j = i if rng.random() < 0.9 else int(rng.integers(0, len(p2)))
        r1.append(random_seq(rng, int(rng.integers(0, 6))) + mutate(rng, t1.replace("-" * 7, p1[i]), 0.06, 0.01, 0.02) + random_seq(rng, 3))
        r2.append(random_seq(rng, int(rng.integers(0, 6))) + mutate(rng, t2.replace("-" * 7, p2[j]), 0.06, 0.01, 0.02) + random_seq(rng, 3))
    f1, f2 = fastq(r1), fastq(r2)
    want = kref.trace_dual(f1, t1, False, subs[0], p1, f2, t2, False, subs[1], p2, False, use_first)
    got = gpu.trace_dual(f1, t1, False, subs[0], p1, f2, t2, False, subs[1], p2, False, use_first)
    assert np.array_equal(np.asarray(got).ravel(), np.asarray(want).ravel())
    assert (np.asarray(want) >= 0).mean() > 0.2
