"""TEST INFRASTRUCTURE ONLY: one ctypes binding for both oracles.

oracle/ref_harness.cpp (the unmodified reference, prefix ``kref_``) and
oracle/kaori_port.c (the plain-C restatement, prefix ``kport_``) export the
same signatures, so one class drives either.  The product package
(screencounter_b200/) never imports this module.
"""
import ctypes as C
import os

import numpy as np

DUP_FIRST, DUP_LAST, DUP_NONE, DUP_ERROR = 0, 1, 2, 3


class KaoriError(RuntimeError):
    pass


def _src(fastq):
    """fastq: bytes (in-memory FASTQ text) or str (file path) -> (path, data, size)."""
    if isinstance(fastq, (bytes, bytearray, memoryview)):
        b = bytes(fastq)
        return None, b, len(b)
    return os.fsencode(fastq), None, 0


def _strs(seqs):
    enc = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
    arr = (C.c_char_p * max(len(enc), 1))(*enc)
    return arr, enc


def _ip(a):
    return a.ctypes.data_as(C.c_void_p)


def _flat_pools(pools):
    nchoices = len(pools[0]) if pools else 0
    flat = []
    for p in pools:
        if len(p) != nchoices:
            raise ValueError("all pools must have the same number of choices for the flattened call")
        flat.extend(p)
    arr, keep = _strs(flat)
    return arr, keep, nchoices


class OracleBinding:
    def __init__(self, path, prefix, hint):
        self.path = path
        self.prefix = prefix
        self.hint = hint
        self._lib = None

    def available(self):
        return os.path.exists(self.path)

    def lib(self):
        if self._lib is None:
            if not self.available():
                raise RuntimeError("%s missing: %s" % (self.path, self.hint))
            L = C.CDLL(self.path)
            self._f("last_error", L).restype = C.c_char_p
            self._f("table_size", L).restype = C.c_size_t
            self._f("table_size", L).argtypes = [C.c_void_p]
            self._f("table_width", L).argtypes = [C.c_void_p]
            self._f("table_copy", L).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            self._f("table_copy", L).restype = None
            self._f("table_free", L).argtypes = [C.c_void_p]
            self._f("table_free", L).restype = None
            self._lib = L
        return self._lib

    def _f(self, name, L=None):
        return getattr(L if L is not None else self.lib(), self.prefix + name)

    def _check(self, status):
        if status != 0:
            raise KaoriError(self._f("last_error")().decode("latin-1"))

    def _table(self, handle, kind):
        n = self._f("table_size")(handle)
        w = self._f("table_width")(handle)
        freq = np.zeros(n, dtype=np.int32)
        if kind == "combo":
            keys = np.zeros((n, w), dtype=np.int32)
            self._f("table_copy")(handle, _ip(keys), None, _ip(freq))
            self._f("table_free")(handle)
            return keys, freq
        buf = C.create_string_buffer(max(n * w, 1))
        self._f("table_copy")(handle, None, buf, _ip(freq))
        self._f("table_free")(handle)
        raw = buf.raw[: n * w]
        seqs = [raw[i * w:(i + 1) * w].decode("latin-1") for i in range(n)]
        return seqs, freq

    def count_reads(self, fastq):
        p, d, s = _src(fastq)
        n = C.c_longlong()
        b = C.c_longlong()
        self._check(self._f("count_reads")(p, d, C.c_size_t(s), C.byref(n), C.byref(b)))
        return n.value, b.value


    def parse(self, fastq):
        """The reference parser's view of a FASTQ: list of sequences (latin-1 str)."""
        n, b = self.count_reads(fastq)
        p, d, s = _src(fastq)
        bases = C.create_string_buffer(max(b, 1))
        off = np.zeros(n + 1, dtype=np.int64)
        self._check(self._f("parse")(p, d, C.c_size_t(s), bases, _ip(off)))
        raw = bases.raw
        return [raw[off[i]:off[i + 1]].decode("latin-1") for i in range(n)]


    def count_single(self, fastq, template, strand, pool, mismatches, use_first, nthreads=1):
        p, d, s = _src(fastq)
        arr, keep = _strs(pool)
        counts = np.zeros(len(pool), dtype=np.int32)
        total = C.c_int()
        self._check(self._f("count_single")(p, d, C.c_size_t(s), template.encode("latin-1"), int(strand), arr, len(pool),
                                       int(mismatches), int(bool(use_first)), int(nthreads), _ip(counts), C.byref(total)))
        return counts, total.value


    def count_single_paired(self, fastq1, fastq2, template, strand, pool, mismatches, use_first, nthreads=1):
        """SingleBarcodePairedEnd (compiled reference only)."""
        p1, d1, s1 = _src(fastq1)
        p2, d2, s2 = _src(fastq2)
        arr, keep = _strs(pool)
        counts = np.zeros(len(pool), dtype=np.int32)
        total = C.c_int()
        self._check(self._f("count_single_paired")(p1, d1, C.c_size_t(s1), p2, d2, C.c_size_t(s2), template.encode("latin-1"), int(strand),
                                                   arr, len(pool), int(mismatches), int(bool(use_first)), int(nthreads), _ip(counts),
                                                   C.byref(total)))
        return counts, total.value

    def trace_single(self, fastq, template, strand, pool, mismatches, use_first):
        n, _ = self.count_reads(fastq)
        p, d, s = _src(fastq)
        arr, keep = _strs(pool)
        index = np.zeros(n, dtype=np.int32)
        info = np.zeros((n, 4), dtype=np.int32)
        nreads = C.c_longlong()
        self._check(self._f("trace_single")(p, d, C.c_size_t(s), template.encode("latin-1"), int(strand), arr, len(pool),
                                       int(mismatches), int(bool(use_first)), _ip(index), _ip(info),
                                       C.c_longlong(n), C.byref(nreads)))
        return index, info


    def count_random(self, fastq, template, strand, mismatches, use_first, nthreads=1):
        p, d, s = _src(fastq)
        handle = C.c_void_p()
        total = C.c_int()
        self._check(self._f("count_random")(p, d, C.c_size_t(s), template.encode("latin-1"), int(strand), int(mismatches),
                                       int(bool(use_first)), int(nthreads), C.byref(handle), C.byref(total)))
        seqs, freq = self._table(handle, "random")
        return seqs, freq, total.value


    def count_combo_single(self, fastq, template, strand, pool1, pool2, mismatches, use_first, nthreads=1):
        p, d, s = _src(fastq)
        a1, k1 = _strs(pool1)
        a2, k2 = _strs(pool2)
        handle = C.c_void_p()
        total = C.c_int()
        self._check(self._f("count_combo_single")(p, d, C.c_size_t(s), template.encode("latin-1"), int(strand), a1, len(pool1), a2, len(pool2),
                                             int(mismatches), int(bool(use_first)), int(nthreads), C.byref(handle), C.byref(total)))
        keys, freq = self._table(handle, "combo")
        return keys, freq, total.value


    def trace_combo_single(self, fastq, template, strand, pool1, pool2, mismatches, use_first):
        n, _ = self.count_reads(fastq)
        p, d, s = _src(fastq)
        a1, k1 = _strs(pool1)
        a2, k2 = _strs(pool2)
        combo = np.zeros((n, 2), dtype=np.int32)
        nreads = C.c_longlong()
        self._check(self._f("trace_combo_single")(p, d, C.c_size_t(s), template.encode("latin-1"), int(strand), a1, len(pool1), a2, len(pool2),
                                             int(mismatches), int(bool(use_first)), _ip(combo), C.c_longlong(n), C.byref(nreads)))
        return combo


    def count_dual_single_end(self, fastq, template, pools, strand, mismatches, use_first, diagnostics=False, nthreads=1):
        p, d, s = _src(fastq)
        arr, keep, nchoices = _flat_pools(pools)
        counts = np.zeros(nchoices, dtype=np.int32)
        total = C.c_int()
        handle = C.c_void_p()
        self._check(self._f("count_dual_single_end")(p, d, C.c_size_t(s), template.encode("latin-1"), arr, len(pools), nchoices, int(strand),
                                                int(mismatches), int(bool(use_first)), int(bool(diagnostics)), int(nthreads),
                                                _ip(counts), C.byref(total), C.byref(handle)))
        if diagnostics:
            keys, freq = self._table(handle, "combo")
            return counts, total.value, keys, freq
        return counts, total.value


    def trace_dual_single_end(self, fastq, template, pools, strand, mismatches, use_first):
        n, _ = self.count_reads(fastq)
        p, d, s = _src(fastq)
        arr, keep, nchoices = _flat_pools(pools)
        index = np.zeros(n, dtype=np.int32)
        nreads = C.c_longlong()
        self._check(self._f("trace_dual_single_end")(p, d, C.c_size_t(s), template.encode("latin-1"), arr, len(pools), nchoices, int(strand),
                                                int(mismatches), int(bool(use_first)), _ip(index), C.c_longlong(n), C.byref(nreads)))
        return index


    def count_dual(self, fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
                   randomized, use_first, diagnostics=False, nthreads=1):
        p1, d1, s1 = _src(fastq1)
        p2, d2, s2 = _src(fastq2)
        a1, k1 = _strs(pool1)
        a2, k2 = _strs(pool2)
        counts = np.zeros(len(pool1), dtype=np.int32)
        total = C.c_int()
        b1 = C.c_int()
        b2 = C.c_int()
        handle = C.c_void_p()
        self._check(self._f("count_dual")(p1, d1, C.c_size_t(s1), template1.encode("latin-1"), int(bool(reverse1)), int(mismatches1), a1, len(pool1),
                                     p2, d2, C.c_size_t(s2), template2.encode("latin-1"), int(bool(reverse2)), int(mismatches2), a2, len(pool2),
                                     int(bool(randomized)), int(bool(use_first)), int(bool(diagnostics)), int(nthreads),
                                     _ip(counts), C.byref(total), C.byref(handle), C.byref(b1), C.byref(b2)))
        if diagnostics:
            keys, freq = self._table(handle, "combo")
            return counts, total.value, keys, freq, b1.value, b2.value
        return counts, total.value


    def trace_dual(self, fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
                   randomized, use_first, fresh_state=True):
        n, _ = self.count_reads(fastq1)
        p1, d1, s1 = _src(fastq1)
        p2, d2, s2 = _src(fastq2)
        a1, k1 = _strs(pool1)
        a2, k2 = _strs(pool2)
        index = np.zeros(n, dtype=np.int32)
        npairs = C.c_longlong()
        self._check(self._f("trace_dual")(p1, d1, C.c_size_t(s1), template1.encode("latin-1"), int(bool(reverse1)), int(mismatches1), a1, len(pool1),
                                     p2, d2, C.c_size_t(s2), template2.encode("latin-1"), int(bool(reverse2)), int(mismatches2), a2, len(pool2),
                                     int(bool(randomized)), int(bool(use_first)), int(fresh_state),
                                     _ip(index), C.c_longlong(n), C.byref(npairs)))
        return index


    def count_combo_paired(self, fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
                           randomized, use_first, nthreads=1):
        p1, d1, s1 = _src(fastq1)
        p2, d2, s2 = _src(fastq2)
        a1, k1 = _strs(pool1)
        a2, k2 = _strs(pool2)
        total = C.c_int()
        b1 = C.c_int()
        b2 = C.c_int()
        handle = C.c_void_p()
        self._check(self._f("count_combo_paired")(p1, d1, C.c_size_t(s1), template1.encode("latin-1"), int(bool(reverse1)), int(mismatches1), a1, len(pool1),
                                             p2, d2, C.c_size_t(s2), template2.encode("latin-1"), int(bool(reverse2)), int(mismatches2), a2, len(pool2),
                                             int(bool(randomized)), int(bool(use_first)), int(nthreads),
                                             C.byref(handle), C.byref(total), C.byref(b1), C.byref(b2)))
        keys, freq = self._table(handle, "combo")
        return keys, freq, total.value, b1.value, b2.value


    def trace_combo_paired(self, fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
                           randomized, use_first):
        n, _ = self.count_reads(fastq1)
        p1, d1, s1 = _src(fastq1)
        p2, d2, s2 = _src(fastq2)
        a1, k1 = _strs(pool1)
        a2, k2 = _strs(pool2)
        combo = np.zeros((n, 2), dtype=np.int32)
        code = np.zeros(n, dtype=np.int32)
        npairs = C.c_longlong()
        self._check(self._f("trace_combo_paired")(p1, d1, C.c_size_t(s1), template1.encode("latin-1"), int(bool(reverse1)), int(mismatches1), a1, len(pool1),
                                             p2, d2, C.c_size_t(s2), template2.encode("latin-1"), int(bool(reverse2)), int(mismatches2), a2, len(pool2),
                                             int(bool(randomized)), int(bool(use_first)),
                                             _ip(combo), _ip(code), C.c_longlong(n), C.byref(npairs)))
        return combo, code


    def match_barcodes(self, seqs, choices, substitutions, reverse, duplicates=DUP_ERROR):
        """0-based index, -1 where R would give NA (src/match_barcodes.cpp:24-30)."""
        a, k = _strs(seqs)
        c, kc = _strs(choices)
        index = np.zeros(len(seqs), dtype=np.int32)
        mm = np.zeros(len(seqs), dtype=np.int32)
        self._check(self._f("match_barcodes")(a, len(seqs), c, len(choices), int(substitutions), int(bool(reverse)), int(duplicates),
                                         _ip(index), _ip(mm)))
        return index, mm


    def search_any(self, seqs, caps, choices, max_mismatches, reverse=False, duplicates=DUP_ERROR):
        a, k = _strs(seqs)
        c, kc = _strs(choices)
        caps = np.ascontiguousarray(caps, dtype=np.int32)
        index = np.zeros(len(seqs), dtype=np.int32)
        mm = np.zeros(len(seqs), dtype=np.int32)
        self._check(self._f("search_any")(a, len(seqs), _ip(caps), c, len(choices), int(max_mismatches), int(bool(reverse)),
                                     int(duplicates), _ip(index), _ip(mm)))
        return index, mm


    def search_segmented2(self, seqs, caps, choices, len1, len2, max1, max2, duplicates=DUP_ERROR):
        a, k = _strs(seqs)
        c, kc = _strs(choices)
        caps = np.ascontiguousarray(caps, dtype=np.int32).reshape(-1, 2)
        index = np.zeros(len(seqs), dtype=np.int32)
        mm = np.zeros(len(seqs), dtype=np.int32)
        self._check(self._f("search_segmented2")(a, len(seqs), _ip(caps), c, len(choices), int(len1), int(len2), int(max1), int(max2),
                                            int(duplicates), _ip(index), _ip(mm)))
        return index, mm
