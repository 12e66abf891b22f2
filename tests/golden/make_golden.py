#!/usr/bin/env python
"""Generates tests/golden/reference_vectors.json.gz: seeded adversarial inputs for every entry point
together with the outputs of the UNMODIFIED reference (kaori compiled from /root/reference into
oracle/_ref/libkaori_ref.so -- `make -C oracle ref`).  Run in the build container, where the
reference tree exists; the fixture then travels with the repo, so the C restatement (CPU tests) and
the CUDA path (GPU tests) are checked against real reference outputs on machines that have neither
/root/reference nor oracle/_ref.

    python tests/golden/make_golden.py
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import kref  # noqa: E402
from util import adversarial_reads, dense_pool, distinct_pool, fastq, random_seq, revcomp  # noqa: E402

STRANDS = {"original": 0, "reverse": 1, "both": 2}


def L(a):
    return np.asarray(a).tolist()


def single_cases(rng):
    out = []
    for name, template, npool, strand, mm, use_first, read_len in [
        ("single_c1_shape", "CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG", 300, "original", 0, True, 75),
        ("single_c2_shape", "CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG", 400, "both", 1, True, 75),
        ("single_dense_best", "ACGT" + "-" * 8 + "TGCA", 120, "both", 2, False, None),
        ("single_dense_first", "ACGT" + "-" * 8 + "TGCA", 120, "both", 1, True, None),
        ("single_reverse_asym", "AAAAACGT" + "-" * 10 + "ACGT", 80, "reverse", 1, False, None),
        ("single_long_barcode", "ACGTAC" + "-" * 40 + "GGTCA", 60, "both", 2, True, None),
    ]:
        Lv = template.count("-")
        pool = dense_pool(rng, npool, Lv) if Lv <= 10 else distinct_pool(rng, npool, Lv)
        reads = adversarial_reads(rng, 1500, template, [pool], strand=strand, read_len=read_len)
        f = fastq(reads)
        counts, total = kref.count_single(f, template, STRANDS[strand], pool, mm, use_first)
        index, info = kref.trace_single(f, template, STRANDS[strand], pool, mm, use_first)
        out.append({"name": name, "kind": "single", "template": template, "strand": STRANDS[strand], "pool": pool, "mismatches": mm,
                    "use_first": use_first, "reads": reads, "counts": L(counts), "total": int(total), "index": L(index), "info": L(info)})
    return out


def random_cases(rng):
    out = []
    for name, template, strand, mm, use_first in [
        ("random_sym", "ACGTACGTACGT" + "-" * 16 + "TGCATGCATGCA", "both", 1, True),
        ("random_asym_quirkB", "AAAAACGT------ACGT", "both", 1, False),
        ("random_fwd_best", "CAG--------T", "original", 2, False),
    ]:
        Lv = template.count("-")
        truth = [random_seq(rng, Lv) for _ in range(50)]
        reads = adversarial_reads(rng, 1500, template, [truth], strand=strand, lower_rate=0.0)
        f = fastq(reads)
        seqs, freq, total = kref.count_random(f, template, STRANDS[strand], mm, use_first)
        order = np.argsort(np.array(list(seqs)))
        out.append({"name": name, "kind": "random", "template": template, "strand": STRANDS[strand], "mismatches": mm, "use_first": use_first,
                    "reads": reads, "seqs": [list(seqs)[i] for i in order], "freq": L(np.asarray(freq)[order]), "total": int(total)})
    return out


def combo_cases(rng):
    out = []
    for name, template, strand, mm, use_first in [
        ("combo_c4_shape", "ACGTACGT" + "-" * 20 + "TTGCAACG" + "-" * 20 + "GGATCCAA", "original", 1, True),
        ("combo_dense_best", "AAAA" + "-" * 6 + "CC" + "-" * 6 + "GGGG", "both", 2, False),
        ("combo_asym_reverse", "AAAAC" + "-" * 5 + "CGC" + "-" * 7 + "GG", "both", 1, True),
    ]:
        lens = [len(run) for run in template.replace("A", " ").replace("C", " ").replace("G", " ").replace("T", " ").split()]
        p1 = dense_pool(rng, 40, lens[0]) if lens[0] <= 10 else distinct_pool(rng, 40, lens[0])
        p2 = dense_pool(rng, 30, lens[1]) if lens[1] <= 10 else distinct_pool(rng, 30, lens[1])
        reads = adversarial_reads(rng, 1500, template, [p1, p2], strand=strand)
        f = fastq(reads)
        keys, freq, total = kref.count_combo_single(f, template, STRANDS[strand], p1, p2, mm, use_first)
        out.append({"name": name, "kind": "combo_single", "template": template, "strand": STRANDS[strand], "pool1": p1, "pool2": p2,
                    "mismatches": mm, "use_first": use_first, "reads": reads, "keys": L(np.asarray(keys).reshape(len(freq), -1)),
                    "freq": L(freq), "total": int(total)})
    return out


def dual_se_cases(rng):
    out = []
    for name, template, strand, mm, use_first, diagnostics in [
        ("dualse_first", "AAAA" + "-" * 6 + "CC" + "-" * 6 + "GGGG", "both", 1, True, False),
        ("dualse_best_diag", "AAAA" + "-" * 6 + "CC" + "-" * 6 + "GGGG", "both", 2, False, True),
    ]:
        a, b = dense_pool(rng, 12, 6), dense_pool(rng, 12, 6)
        rows = sorted({(int(rng.integers(0, 12)), int(rng.integers(0, 12))) for _ in range(40)})
        pools = [[a[i] for i, _ in rows], [b[j] for _, j in rows]]
        reads = []
        for _ in range(1500):
            if rng.random() < 0.7:
                i, j = rows[int(rng.integers(0, len(rows)))]
            else:
                i, j = int(rng.integers(0, 12)), int(rng.integers(0, 12))
            reads.extend(adversarial_reads(rng, 1, template, [[a[i]], [b[j]]], strand=strand))
        f = fastq(reads)
        res = kref.count_dual_single_end(f, template, pools, STRANDS[strand], mm, use_first, diagnostics)
        case = {"name": name, "kind": "dual_single_end", "template": template, "strand": STRANDS[strand], "pools": pools, "mismatches": mm,
                "use_first": use_first, "diagnostics": diagnostics, "reads": reads, "counts": L(res[0]), "total": int(res[1])}
        if diagnostics:
            case["keys"] = L(np.asarray(res[2]).reshape(len(res[3]), -1))
            case["freq"] = L(res[3])
        out.append(case)
    return out


def paired_inputs(rng, t1, t2, p1, p2, rows, n, rev1, rev2, randomized):
    r1, r2 = [], []
    for _ in range(n):
        if rows is not None and rng.random() < 0.75:
            i, j = rows[int(rng.integers(0, len(rows)))]
        else:
            i, j = int(rng.integers(0, len(p1))), int(rng.integers(0, len(p2)))
        a = adversarial_reads(rng, 1, t1, [[p1[i]]], strand="reverse" if rev1 else "original")[0]
        b = adversarial_reads(rng, 1, t2, [[p2[j]]], strand="reverse" if rev2 else "original")[0]
        if randomized and rng.random() < 0.5:
            a, b = b, a
        r1.append(a)
        r2.append(b)
    return r1, r2


def dual_pe_cases(rng):
    out = []
    for name, mm1, mm2, rev1, rev2, randomized, use_first, diagnostics in [
        ("dualpe_first", 1, 1, False, False, False, True, False),
        ("dualpe_best_random", 1, 1, False, True, True, False, False),
        ("dualpe_quirkA_caps", 1, 0, False, False, False, True, False),
        ("dualpe_diag", 1, 1, False, False, True, True, True),
    ]:
        t1, t2 = "ACGT" + "-" * 7 + "TGCA", "GGAC" + "-" * 6 + "CCTT"
        a, b = dense_pool(rng, 14, 7), dense_pool(rng, 14, 6, frac_neighbours=0.6)
        rows = sorted({(int(rng.integers(0, 14)), int(rng.integers(0, 14))) for _ in range(60)})
        pool1, pool2 = [a[i] for i, _ in rows], [b[j] for _, j in rows]
        r1, r2 = paired_inputs(rng, t1, t2, a, b, [(i, j) for i, j in rows], 1500, rev1, rev2, randomized)
        f1, f2 = fastq(r1), fastq(r2)
        res = kref.count_dual(f1, t1, rev1, mm1, pool1, f2, t2, rev2, mm2, pool2, randomized, use_first, diagnostics)
        # the reference's own answer can depend on read order through its result cache;
        # the per-pair trace with a fresh state per pair is the cache-free definition the GPU implements
        fresh = kref.trace_dual(f1, t1, rev1, mm1, pool1, f2, t2, rev2, mm2, pool2, randomized, use_first, fresh_state=1)
        normal = kref.trace_dual(f1, t1, rev1, mm1, pool1, f2, t2, rev2, mm2, pool2, randomized, use_first, fresh_state=0)
        case = {"name": name, "kind": "dual", "template1": t1, "template2": t2, "reverse1": rev1, "reverse2": rev2, "mismatches1": mm1,
                "mismatches2": mm2, "pool1": pool1, "pool2": pool2, "randomized": randomized, "use_first": use_first, "diagnostics": diagnostics,
                "reads1": r1, "reads2": r2, "counts": L(res[0]), "total": int(res[1]), "index_fresh": L(fresh),
                "order_dependent_pairs": int((np.asarray(fresh) != np.asarray(normal)).sum())}
        if diagnostics:
            case["keys"] = L(np.asarray(res[2]).reshape(len(res[3]), -1))
            case["freq"] = L(res[3])
            case["barcode1_only"] = int(res[4])
            case["barcode2_only"] = int(res[5])
        out.append(case)
    return out


def combo_pe_cases(rng):
    out = []
    for name, mm1, mm2, rev1, rev2, randomized, use_first in [
        ("combope_first", 1, 1, False, False, False, True),
        ("combope_best_random", 1, 2, False, True, True, False),
    ]:
        t1, t2 = "ACGT" + "-" * 7 + "TGCA", "GGAC" + "-" * 6 + "CCTT"
        a, b = dense_pool(rng, 20, 7), dense_pool(rng, 16, 6)
        r1, r2 = paired_inputs(rng, t1, t2, a, b, None, 1500, rev1, rev2, randomized)
        f1, f2 = fastq(r1), fastq(r2)
        keys, freq, total, b1, b2 = kref.count_combo_paired(f1, t1, rev1, mm1, a, f2, t2, rev2, mm2, b, randomized, use_first)
        out.append({"name": name, "kind": "combo_paired", "template1": t1, "template2": t2, "reverse1": rev1, "reverse2": rev2,
                    "mismatches1": mm1, "mismatches2": mm2, "pool1": a, "pool2": b, "randomized": randomized, "use_first": use_first,
                    "reads1": r1, "reads2": r2, "keys": L(np.asarray(keys).reshape(len(freq), -1)), "freq": L(freq), "total": int(total),
                    "barcode1_only": int(b1), "barcode2_only": int(b2)})
    return out


def match_cases(rng):
    out = []
    for name, L_, subs, reverse, iupac in [("match_plain", 10, 1, False, False), ("match_reverse_2mm", 8, 2, True, False),
                                           ("match_iupac", 10, 1, False, True)]:
        choices = dense_pool(rng, 50, L_)
        if iupac:
            choices = ["AAAAABAAAA", "CCCCCDCCCC", "GGGGGHGGGG", "TTTTTVTTTT", "ACGTNACGTA"] + distinct_pool(rng, 20, L_)
        seqs = []
        for _ in range(600):
            base = choices[int(rng.integers(0, len(choices)))]
            base = "".join(c if c in "ACGT" else "ACGT"[int(rng.integers(0, 4))] for c in base)
            if reverse:
                base = revcomp(base)
            s = list(base)
            for _k in range(int(rng.integers(0, 4))):
                s[int(rng.integers(0, L_))] = "ACGTN"[int(rng.integers(0, 5))]
            seqs.append("".join(s))
        try:
            index, mm = kref.match_barcodes(seqs, choices, subs, reverse)
        except Exception as e:  # overlapping IUPAC expansions are a construction error in the reference
            out.append({"name": name, "kind": "match", "choices": choices, "seqs": seqs, "substitutions": subs, "reverse": reverse, "error": str(e)})
            continue
        out.append({"name": name, "kind": "match", "choices": choices, "seqs": seqs, "substitutions": subs, "reverse": reverse,
                    "index": L(index), "mm": L(mm)})
    return out


def main():
    if not kref.available():
        raise SystemExit("oracle/_ref/libkaori_ref.so is missing: run `make -C oracle ref` where /root/reference exists")
    rng = np.random.default_rng(20261018)
    cases = (single_cases(rng) + random_cases(rng) + combo_cases(rng) + dual_se_cases(rng) + dual_pe_cases(rng) + combo_pe_cases(rng) +
             match_cases(rng))
    doc = {"generator": "tests/golden/make_golden.py", "reference": "screenCounter 1.5.1 / kaori v1.1.1 compiled from /root/reference",
           "seed": 20261018, "cases": cases}
    path = os.path.join(HERE, "reference_vectors.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(doc, separators=(",", ":")).encode())
    print("wrote %s: %d cases, %d bytes" % (path, len(cases), os.path.getsize(path)))
    for c in cases:
        print("  %-24s %s" % (c["name"], c["kind"]))


if __name__ == "__main__":
    main()
