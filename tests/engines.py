"""Adapter giving the CUDA path (screencounter_b200.rcpp) the call shapes of the oracles
(oracle/_binding.py), so one test body drives the oracle and the product alike."""
import numpy as np

from screencounter_b200 import rcpp


class GpuEngine:
    name = "gpu"

    def count_single(self, fastq, template, strand, pool, mismatches, use_first, nthreads=1):
        counts, total = rcpp.count_single_barcodes(fastq, template, strand, pool, mismatches, use_first, nthreads)
        return counts, total

    def trace_single(self, fastq, template, strand, pool, mismatches, use_first):
        counts, total, (index, info) = rcpp.count_single_barcodes(fastq, template, strand, pool, mismatches, use_first, 1, trace=True)
        return index, info

    def match_barcodes(self, seqs, choices, substitutions, reverse):
        idx, mm = rcpp.match_barcodes(seqs, choices, substitutions, reverse)
        index = np.array([-1 if i is None else i - 1 for i in idx], dtype=np.int32)
        mms = np.array([-1 if m is None else m for m in mm], dtype=np.int32)
        return index, mms

    def count_random(self, fastq, template, strand, mismatches, use_first, nthreads=1):
        (seqs, freq), total = rcpp.count_random_barcodes(fastq, template, strand, mismatches, use_first, nthreads, as_array=False)
        return seqs, freq, total

    def count_combo_single(self, fastq, template, strand, pool1, pool2, mismatches, use_first, nthreads=1):
        keys, freq, total = rcpp.count_combo_barcodes_single(fastq, template, strand, [pool1, pool2], mismatches, use_first, nthreads)
        return keys.T.copy(), freq, total[0]

    def trace_combo_single(self, fastq, template, strand, pool1, pool2, mismatches, use_first):
        out = rcpp.count_combo_barcodes_single(fastq, template, strand, [pool1, pool2], mismatches, use_first, 1, trace=True)
        return out[3]

    def count_dual_single_end(self, fastq, template, pools, strand, mismatches, use_first, diagnostics=False, nthreads=1):
        out = rcpp.count_dual_barcodes_single_end(fastq, template, pools, strand, mismatches, use_first, diagnostics, nthreads)
        if diagnostics:
            counts, (keys, freq), total = out
            return counts, total[0], keys.T.copy(), freq
        counts, total = out
        return counts, total[0]

    def trace_dual_single_end(self, fastq, template, pools, strand, mismatches, use_first):
        out = rcpp.count_dual_barcodes_single_end(fastq, template, pools, strand, mismatches, use_first, False, 1, trace=True)
        return out[2]

    def count_dual(self, fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
                   randomized, use_first, diagnostics=False, nthreads=1):
        out = rcpp.count_dual_barcodes(fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
                                       randomized, use_first, diagnostics, nthreads)
        if diagnostics:
            counts, (keys, freq), total, b1, b2 = out
            return counts, total[0], keys.T.copy(), freq, b1[0], b2[0]
        counts, total = out
        return counts, total[0]

    def trace_dual(self, fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
                   randomized, use_first, fresh_state=2):
        out = rcpp.count_dual_barcodes(fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
                                       randomized, use_first, False, 1, trace=True)
        return out[2]

    def count_combo_paired(self, fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
                           randomized, use_first, nthreads=1):
        keys, freq, total, b1, b2 = rcpp.count_combo_barcodes_paired(fastq1, template1, reverse1, mismatches1, pool1,
                                                                     fastq2, template2, reverse2, mismatches2, pool2,
                                                                     randomized, use_first, nthreads)[:5]
        return keys.T.copy(), freq, total[0], b1[0], b2[0]

    def trace_combo_paired(self, fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
                           randomized, use_first):
        out = rcpp.count_combo_barcodes_paired(fastq1, template1, reverse1, mismatches1, pool1,
                                               fastq2, template2, reverse2, mismatches2, pool2,
                                               randomized, use_first, 1, trace=True)
        return out[5]
