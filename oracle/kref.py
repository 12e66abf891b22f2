"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libkaori_ref.so.

That library is the unmodified reference (kaori v1.1.1 as vendored in
screenCounter 1.5.1) compiled behind oracle/ref_harness.cpp.  It is the ground
truth the parity tests and bench.py's cpu_baseline leg compare against.  The
product package (screencounter_b200/) never imports this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libkaori_ref.so")

_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libkaori_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        _lib = C.CDLL(_PATH)
        _lib.kref_last_error.restype = C.c_char_p
        _lib.kref_table_size.restype = C.c_size_t
        _lib.kref_table_size.argtypes = [C.c_void_p]
        _lib.kref_table_width.argtypes = [C.c_void_p]
        _lib.kref_table_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.kref_table_copy.restype = None
        _lib.kref_table_free.argtypes = [C.c_void_p]
        _lib.kref_table_free.restype = None
    return _lib


class KaoriError(RuntimeError):
    pass


def _check(status):
    if status != 0:
        raise KaoriError(lib().kref_last_error().decode())


def _src(fastq):
    """fastq: bytes (in-memory FASTQ text) or str (file path) -> (path, data, size, keepalive)."""
    if isinstance(fastq, (bytes, bytearray, memoryview)):
        b = bytes(fastq)
        return None, b, len(b)
    return os.fsencode(fastq), None, 0


def _strs(seqs):
    enc = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    arr = (C.c_char_p * max(len(enc), 1))(*enc)
    return arr, enc


def _ip(a):
    return a.ctypes.data_as(C.c_void_p)


def _table(handle, kind):
    L = lib()
    n = L.kref_table_size(handle)
    w = L.kref_table_width(handle)
    freq = np.zeros(n, dtype=np.int32)
    if kind == "combo":
        keys = np.zeros((n, w), dtype=np.int32)
        L.kref_table_copy(handle, _ip(keys), None, _ip(freq))
        L.kref_table_free(handle)
        return keys, freq
    buf = C.create_string_buffer(max(n * w, 1))
    L.kref_table_copy(handle, None, buf, _ip(freq))
    L.kref_table_free(handle)
    raw = buf.raw[: n * w]
    seqs = [raw[i * w:(i + 1) * w].decode("latin-1") for i in range(n)]
    return seqs, freq


def count_reads(fastq):
    p, d, s = _src(fastq)
    n = C.c_longlong()
    b = C.c_longlong()
    _check(lib().kref_count_reads(p, d, C.c_size_t(s), C.byref(n), C.byref(b)))
    return n.value, b.value


def parse(fastq):
    """The reference parser's view of a FASTQ: list of sequences (latin-1 str)."""
    n, b = count_reads(fastq)
    p, d, s = _src(fastq)
    bases = C.create_string_buffer(max(b, 1))
    off = np.zeros(n + 1, dtype=np.int64)
    _check(lib().kref_parse(p, d, C.c_size_t(s), bases, _ip(off)))
    raw = bases.raw
    return [raw[off[i]:off[i + 1]].decode("latin-1") for i in range(n)]


def count_single(fastq, template, strand, pool, mismatches, use_first, nthreads=1):
    p, d, s = _src(fastq)
    arr, keep = _strs(pool)
    counts = np.zeros(len(pool), dtype=np.int32)
    total = C.c_int()
    _check(lib().kref_count_single(p, d, C.c_size_t(s), template.encode(), int(strand), arr, len(pool),
                                   int(mismatches), int(bool(use_first)), int(nthreads), _ip(counts), C.byref(total)))
    return counts, total.value


def trace_single(fastq, template, strand, pool, mismatches, use_first):
    n, _ = count_reads(fastq)
    p, d, s = _src(fastq)
    arr, keep = _strs(pool)
    index = np.zeros(n, dtype=np.int32)
    info = np.zeros((n, 4), dtype=np.int32)
    nreads = C.c_longlong()
    _check(lib().kref_trace_single(p, d, C.c_size_t(s), template.encode(), int(strand), arr, len(pool),
                                   int(mismatches), int(bool(use_first)), _ip(index), _ip(info),
                                   C.c_longlong(n), C.byref(nreads)))
    return index, info


def count_random(fastq, template, strand, mismatches, use_first, nthreads=1):
    p, d, s = _src(fastq)
    handle = C.c_void_p()
    total = C.c_int()
    _check(lib().kref_count_random(p, d, C.c_size_t(s), template.encode(), int(strand), int(mismatches),
                                   int(bool(use_first)), int(nthreads), C.byref(handle), C.byref(total)))
    seqs, freq = _table(handle, "random")
    return seqs, freq, total.value


def count_combo_single(fastq, template, strand, pool1, pool2, mismatches, use_first, nthreads=1):
    p, d, s = _src(fastq)
    a1, k1 = _strs(pool1)
    a2, k2 = _strs(pool2)
    handle = C.c_void_p()
    total = C.c_int()
    _check(lib().kref_count_combo_single(p, d, C.c_size_t(s), template.encode(), int(strand), a1, len(pool1), a2, len(pool2),
                                         int(mismatches), int(bool(use_first)), int(nthreads), C.byref(handle), C.byref(total)))
    keys, freq = _table(handle, "combo")
    return keys, freq, total.value


def trace_combo_single(fastq, template, strand, pool1, pool2, mismatches, use_first):
    n, _ = count_reads(fastq)
    p, d, s = _src(fastq)
    a1, k1 = _strs(pool1)
    a2, k2 = _strs(pool2)
    combo = np.zeros((n, 2), dtype=np.int32)
    nreads = C.c_longlong()
    _check(lib().kref_trace_combo_single(p, d, C.c_size_t(s), template.encode(), int(strand), a1, len(pool1), a2, len(pool2),
                                         int(mismatches), int(bool(use_first)), _ip(combo), C.c_longlong(n), C.byref(nreads)))
    return combo


def _flat_pools(pools):
    nchoices = len(pools[0]) if pools else 0
    flat = []
    for p in pools:
        if len(p) != nchoices:
            raise ValueError("all pools must have the same number of choices for the flattened call")
        flat.extend(p)
    arr, keep = _strs(flat)
    return arr, keep, nchoices


def count_dual_single_end(fastq, template, pools, strand, mismatches, use_first, diagnostics=False, nthreads=1):
    p, d, s = _src(fastq)
    arr, keep, nchoices = _flat_pools(pools)
    counts = np.zeros(nchoices, dtype=np.int32)
    total = C.c_int()
    handle = C.c_void_p()
    _check(lib().kref_count_dual_single_end(p, d, C.c_size_t(s), template.encode(), arr, len(pools), nchoices, int(strand),
                                            int(mismatches), int(bool(use_first)), int(bool(diagnostics)), int(nthreads),
                                            _ip(counts), C.byref(total), C.byref(handle)))
    if diagnostics:
        keys, freq = _table(handle, "combo")
        return counts, total.value, keys, freq
    return counts, total.value


def trace_dual_single_end(fastq, template, pools, strand, mismatches, use_first):
    n, _ = count_reads(fastq)
    p, d, s = _src(fastq)
    arr, keep, nchoices = _flat_pools(pools)
    index = np.zeros(n, dtype=np.int32)
    nreads = C.c_longlong()
    _check(lib().kref_trace_dual_single_end(p, d, C.c_size_t(s), template.encode(), arr, len(pools), nchoices, int(strand),
                                            int(mismatches), int(bool(use_first)), _ip(index), C.c_longlong(n), C.byref(nreads)))
    return index


def count_dual(fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
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
    _check(lib().kref_count_dual(p1, d1, C.c_size_t(s1), template1.encode(), int(bool(reverse1)), int(mismatches1), a1, len(pool1),
                                 p2, d2, C.c_size_t(s2), template2.encode(), int(bool(reverse2)), int(mismatches2), a2, len(pool2),
                                 int(bool(randomized)), int(bool(use_first)), int(bool(diagnostics)), int(nthreads),
                                 _ip(counts), C.byref(total), C.byref(handle), C.byref(b1), C.byref(b2)))
    if diagnostics:
        keys, freq = _table(handle, "combo")
        return counts, total.value, keys, freq, b1.value, b2.value
    return counts, total.value


def trace_dual(fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
               randomized, use_first, fresh_state=True):
    n, _ = count_reads(fastq1)
    p1, d1, s1 = _src(fastq1)
    p2, d2, s2 = _src(fastq2)
    a1, k1 = _strs(pool1)
    a2, k2 = _strs(pool2)
    index = np.zeros(n, dtype=np.int32)
    npairs = C.c_longlong()
    _check(lib().kref_trace_dual(p1, d1, C.c_size_t(s1), template1.encode(), int(bool(reverse1)), int(mismatches1), a1, len(pool1),
                                 p2, d2, C.c_size_t(s2), template2.encode(), int(bool(reverse2)), int(mismatches2), a2, len(pool2),
                                 int(bool(randomized)), int(bool(use_first)), int(bool(fresh_state)),
                                 _ip(index), C.c_longlong(n), C.byref(npairs)))
    return index


def count_combo_paired(fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
                       randomized, use_first, nthreads=1):
    p1, d1, s1 = _src(fastq1)
    p2, d2, s2 = _src(fastq2)
    a1, k1 = _strs(pool1)
    a2, k2 = _strs(pool2)
    total = C.c_int()
    b1 = C.c_int()
    b2 = C.c_int()
    handle = C.c_void_p()
    _check(lib().kref_count_combo_paired(p1, d1, C.c_size_t(s1), template1.encode(), int(bool(reverse1)), int(mismatches1), a1, len(pool1),
                                         p2, d2, C.c_size_t(s2), template2.encode(), int(bool(reverse2)), int(mismatches2), a2, len(pool2),
                                         int(bool(randomized)), int(bool(use_first)), int(nthreads),
                                         C.byref(handle), C.byref(total), C.byref(b1), C.byref(b2)))
    keys, freq = _table(handle, "combo")
    return keys, freq, total.value, b1.value, b2.value


def trace_combo_paired(fastq1, template1, reverse1, mismatches1, pool1, fastq2, template2, reverse2, mismatches2, pool2,
                       randomized, use_first):
    n, _ = count_reads(fastq1)
    p1, d1, s1 = _src(fastq1)
    p2, d2, s2 = _src(fastq2)
    a1, k1 = _strs(pool1)
    a2, k2 = _strs(pool2)
    combo = np.zeros((n, 2), dtype=np.int32)
    code = np.zeros(n, dtype=np.int32)
    npairs = C.c_longlong()
    _check(lib().kref_trace_combo_paired(p1, d1, C.c_size_t(s1), template1.encode(), int(bool(reverse1)), int(mismatches1), a1, len(pool1),
                                         p2, d2, C.c_size_t(s2), template2.encode(), int(bool(reverse2)), int(mismatches2), a2, len(pool2),
                                         int(bool(randomized)), int(bool(use_first)),
                                         _ip(combo), _ip(code), C.c_longlong(n), C.byref(npairs)))
    return combo, code


DUP_FIRST, DUP_LAST, DUP_NONE, DUP_ERROR = 0, 1, 2, 3


def match_barcodes(seqs, choices, substitutions, reverse, duplicates=DUP_ERROR):
    """0-based index, -1 where R would give NA (src/match_barcodes.cpp:24-30)."""
    a, k = _strs(seqs)
    c, kc = _strs(choices)
    index = np.zeros(len(seqs), dtype=np.int32)
    mm = np.zeros(len(seqs), dtype=np.int32)
    _check(lib().kref_match_barcodes(a, len(seqs), c, len(choices), int(substitutions), int(bool(reverse)), int(duplicates),
                                     _ip(index), _ip(mm)))
    return index, mm


def search_any(seqs, caps, choices, max_mismatches, reverse=False, duplicates=DUP_ERROR):
    a, k = _strs(seqs)
    c, kc = _strs(choices)
    caps = np.ascontiguousarray(caps, dtype=np.int32)
    index = np.zeros(len(seqs), dtype=np.int32)
    mm = np.zeros(len(seqs), dtype=np.int32)
    _check(lib().kref_search_any(a, len(seqs), _ip(caps), c, len(choices), int(max_mismatches), int(bool(reverse)),
                                 int(duplicates), _ip(index), _ip(mm)))
    return index, mm


def search_segmented2(seqs, caps, choices, len1, len2, max1, max2, duplicates=DUP_ERROR):
    a, k = _strs(seqs)
    c, kc = _strs(choices)
    caps = np.ascontiguousarray(caps, dtype=np.int32).reshape(-1, 2)
    index = np.zeros(len(seqs), dtype=np.int32)
    mm = np.zeros(len(seqs), dtype=np.int32)
    _check(lib().kref_search_segmented2(a, len(seqs), _ip(caps), c, len(choices), int(len1), int(len2), int(max1), int(max2),
                                        int(duplicates), _ip(index), _ip(mm)))
    return index, mm
