"""The seven Rcpp-exported entry points of screenCounter, same names and argument order
(reference R/RcppExports.R:4-30, src/RcppExports.cpp:13-150), on top of the C ABI.

`path` arguments accept a file path (str) or in-memory FASTQ text (bytes).  Return values are
Python lists shaped like the R lists the reference returns (SURVEY.md 8.2 row b), with numpy
arrays for the integer vectors.  Indices are 0-based exactly where the C++ glue is 0-based
(the R wrappers add 1); `match_barcodes` returns 1-based indices and None for NA like the glue.
"""
import ctypes as C
import os
import threading

import numpy as np

from ._lib import lib, ScgSource


class ScreenCounterError(RuntimeError):
    """An error raised by the engine; the text is kaori's own wherever the reference has one."""


_local = threading.local()


def context(device=None):
    """Per-thread, per-device context (CUDA is initialised lazily by the first counting call)."""
    if device is None:
        device = int(os.environ.get("SCG_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    cache = getattr(_local, "ctx", None)
    if cache is None:
        cache = _local.ctx = {}
    if device not in cache:
        h = C.c_void_p()
        if lib().scg_ctx_create(C.byref(h), int(device)) != 0:
            raise ScreenCounterError(lib().scg_last_error(None).decode("latin-1"))
        cache[device] = h
    return cache[device]


def _check(ctx, status):
    if status != 0:
        raise ScreenCounterError(lib().scg_last_error(ctx).decode("latin-1"))


class _Src:
    """Keeps the bytes alive for the duration of the call."""

    def __init__(self, fastq):
        if isinstance(fastq, (bytes, bytearray, memoryview)):
            self.keep = fastq if isinstance(fastq, bytes) else bytes(fastq)
            # a pointer INTO the bytes object: no copy of what can be gigabytes of text
            self.struct = ScgSource(None, C.cast(C.c_char_p(self.keep), C.c_void_p), len(self.keep))
        else:
            self.keep = os.fsencode(fastq)
            self.struct = ScgSource(self.keep, None, 0)

    def ref(self):
        return C.byref(self.struct)


def _strs(seqs):
    enc = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
    arr = (C.c_char_p * max(len(enc), 1))(*enc)
    return arr, enc


def _ip(a):
    return a.ctypes.data_as(C.c_void_p)


def _table(handle, kind):
    L = lib()
    n = L.scg_result_rows(handle)
    w = L.scg_result_width(handle)
    freq = np.zeros(n, dtype=np.int32)
    if kind == "combo":
        keys = np.zeros((n, w), dtype=np.int32)
        L.scg_result_copy_table(handle, _ip(keys), None, _ip(freq))
        return keys, freq
    buf = C.create_string_buffer(max(n * w, 1))
    L.scg_result_copy_table(handle, None, buf, _ip(freq))
    raw = buf.raw[: n * w]
    return [raw[i * w:(i + 1) * w].decode("latin-1") for i in range(n)], freq


def _trace(handle):
    L = lib()
    n = L.scg_result_reads(handle)
    w = max(L.scg_result_trace_width(handle), 1)
    index = np.zeros((n, w), dtype=np.int32)
    info = np.zeros(n, dtype=np.uint32)
    L.scg_result_copy_trace(handle, _ip(index), _ip(info))
    return index, info


def decode_info(info):
    """(position, reverse, mismatches, variable_mismatches) columns of a single-barcode trace, -1/0 where not found."""
    info = np.asarray(info, dtype=np.uint32)
    found = (info >> 31) & 1
    out = np.zeros((len(info), 4), dtype=np.int32)
    out[:, 0] = np.where(found, info & 0xFFFFF, -1)
    out[:, 1] = np.where(found, (info >> 30) & 1, 0)
    out[:, 2] = np.where(found, (info >> 20) & 31, -1)
    out[:, 3] = np.where(found, (info >> 25) & 31, -1)
    return out


# ---------------------------------------------------------------------------------------------
def count_single_barcodes(path, constant, strand, pool, mismatches, use_first, nthreads, trace=False, device=None):
    """reference: src/count_single_barcodes.cpp:29-50 -> list(counts, total)."""
    ctx = context(device)
    src = _Src(path)
    arr, keep = _strs(pool)
    counts = np.zeros(len(pool), dtype=np.int32)
    total = C.c_int32()
    handle = C.c_void_p()
    _check(ctx, lib().scg_count_single(ctx, src.ref(), constant.encode("latin-1"), int(strand), arr, len(pool), int(mismatches),
                                       int(bool(use_first)), int(nthreads), _ip(counts), C.byref(total),
                                       C.byref(handle) if trace else None))
    out = [counts, total.value]
    if trace:
        index, info = _trace(handle)
        lib().scg_result_free(handle)
        out.append((index[:, 0], decode_info(info)))
    return out


def match_barcodes(sequences, choices, substitutions, reverse, device=None):
    """reference: src/match_barcodes.cpp:7-37 -> list(index (1-based, None = NA), mismatches (None = NA))."""
    ctx = context(device)
    a, k1 = _strs(sequences)
    c, k2 = _strs(choices)
    index = np.zeros(len(sequences), dtype=np.int32)
    mm = np.zeros(len(sequences), dtype=np.int32)
    _check(ctx, lib().scg_match_barcodes(ctx, a, len(sequences), c, len(choices), int(substitutions), int(bool(reverse)),
                                         _ip(index), _ip(mm)))
    return [[int(i) + 1 if i >= 0 else None for i in index], [int(m) if i >= 0 else None for i, m in zip(index, mm)]]


def timing(device=None):
    import json
    return json.loads(lib().scg_timing_json(context(device)).decode())


def kernel_launches(device=None):
    return int(lib().scg_kernel_launches(context(device)))
