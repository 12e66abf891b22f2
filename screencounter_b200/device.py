"""Resident objects: packed reads and compiled handlers that stay in HBM across calls.

Used by bench.py (and by any caller that streams many files against one library).  Device
buffers may be owned by the caller (e.g. torch tensors: pass `tensor.data_ptr()`), the library
only ever sees plain pointers."""
import ctypes as C

import numpy as np

from ._lib import lib, ScgSynthSpec
from .rcpp import context, _check, _Src, _strs, _ip, ScreenCounterError  # noqa: F401


def _synth_spec(seed, first_read, n_reads, read_len, constant, pools, paired_rows, strand, construct_permille,
                sub_per_10k, n_per_10k, random_space):
    spec = ScgSynthSpec()
    keep = []
    spec.seed = int(seed)
    spec.first_read = int(first_read)
    spec.n_reads = int(n_reads)
    spec.read_len = int(read_len)
    spec.constant = constant.encode("latin-1")
    spec.n_pools = len(pools)
    for v, p in enumerate(pools):
        arr, enc = _strs(p)
        keep.append((arr, enc))
        spec.pools[v] = C.cast(arr, C.POINTER(C.c_char_p))
        spec.n_choices[v] = len(p)
    spec.paired_rows = int(bool(paired_rows))
    spec.strand = int(strand)
    spec.construct_permille = int(construct_permille)
    spec.sub_per_10k = int(sub_per_10k)
    spec.n_per_10k = int(n_per_10k)
    spec.random_space = int(random_space)
    return spec, keep


def _stream(stream):
    # SCG_STREAM_OWN = (void*)-1 selects the context's stream; anything else is a cudaStream_t used as given
    return C.c_void_p(-1 & (2 ** (8 * C.sizeof(C.c_void_p)) - 1)) if stream is None else C.c_void_p(int(stream))


class SynthSpec:
    """The synthetic workload of SURVEY.md 8(d): identical reads on the host (FASTQ text) and on the device."""

    def __init__(self, constant, pools, seed=42, read_len=75, strand=2, paired_rows=False, construct_permille=900,
                 sub_per_10k=100, n_per_10k=10, random_space=0):
        self.constant = constant
        self.pools = [list(p) for p in pools]
        self.seed = seed
        self.read_len = read_len
        self.strand = strand
        self.paired_rows = paired_rows
        self.construct_permille = construct_permille
        self.sub_per_10k = sub_per_10k
        self.n_per_10k = n_per_10k
        self.random_space = random_space

    def _c(self, first_read, n_reads):
        return _synth_spec(self.seed, first_read, n_reads, self.read_len, self.constant, self.pools, self.paired_rows,
                           self.strand, self.construct_permille, self.sub_per_10k, self.n_per_10k, self.random_space)

    def fastq(self, first_read, n_reads):
        """FASTQ text (bytes) of reads [first_read, first_read + n_reads)."""
        spec, keep = self._c(first_read, n_reads)
        used = C.c_size_t()
        if lib().scg_synth_fastq(C.byref(spec), None, C.c_size_t(0), C.byref(used)) != 0:
            raise ScreenCounterError(lib().scg_last_error(None).decode("latin-1"))
        buf = C.create_string_buffer(max(used.value, 1))
        if lib().scg_synth_fastq(C.byref(spec), buf, C.c_size_t(used.value), C.byref(used)) != 0:
            raise ScreenCounterError(lib().scg_last_error(None).decode("latin-1"))
        return buf.raw[: used.value]

    def fastq_pinned(self, first_read, n_reads, device=None):
        """The same text written straight into page-locked host memory (rcpp.PinnedText)."""
        from .rcpp import PinnedText
        spec, keep = self._c(first_read, n_reads)
        used = C.c_size_t()
        if lib().scg_synth_fastq(C.byref(spec), None, C.c_size_t(0), C.byref(used)) != 0:
            raise ScreenCounterError(lib().scg_last_error(None).decode("latin-1"))
        text = PinnedText(used.value, device)
        if lib().scg_synth_fastq(C.byref(spec), text.ptr, C.c_size_t(used.value), C.byref(used)) != 0:
            raise ScreenCounterError(lib().scg_last_error(None).decode("latin-1"))
        text.size = used.value
        return text

    def on_device(self, first_read, n_reads, device=None):
        ctx = context(device)
        spec, keep = self._c(first_read, n_reads)
        handle = C.c_void_p()
        _check(ctx, lib().scg_reads_synthesize(ctx, C.byref(spec), C.byref(handle)))
        return Reads(ctx, handle)


class Reads:
    """Packed reads resident in device memory (tile-planar 2-bit bases + N mask)."""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.handle = handle

    @classmethod
    def from_fastq(cls, fastq, nthreads=1, device=None):
        ctx = context(device)
        src = _Src(fastq)
        handle = C.c_void_p()
        _check(ctx, lib().scg_reads_from_source(ctx, src.ref(), int(nthreads), C.byref(handle)))
        return cls(ctx, handle)

    def __len__(self):
        return int(lib().scg_reads_count(self.handle))

    @property
    def device_bytes(self):
        return int(lib().scg_reads_device_bytes(self.handle))

    def free(self):
        if self.handle:
            lib().scg_reads_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SinglePlan:
    """countSingleBarcodes with the template and library compiled onto the device once."""

    def __init__(self, constant, strand, pool, mismatches, use_first, device=None):
        self.ctx = context(device)
        self.npool = len(pool)
        arr, keep = _strs(pool)
        self.handle = C.c_void_p()
        _check(self.ctx, lib().scg_single_plan_create(self.ctx, constant.encode("latin-1"), int(strand), arr, len(pool),
                                                      int(mismatches), int(bool(use_first)), C.byref(self.handle)))

    def run(self, reads, counts_ptr, index_ptr=None, stream=None):
        """One pass over `reads`; counts (device int32[npool]) are accumulated into.  Asynchronous.
        stream: None = the context's own stream; an int = that cudaStream_t (0 = CUDA's default stream)."""
        _check(self.ctx, lib().scg_single_plan_run(self.handle, reads.handle, C.c_void_p(counts_ptr),
                                                   C.c_void_p(index_ptr) if index_ptr else None, _stream(stream)))

    @property
    def kernel(self):
        return lib().scg_plan_kernel(self.handle).decode()

    def free(self):
        if self.handle:
            lib().scg_plan_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _Plan:
    """Common part of the compiled handlers: the C handle, which kernel ran, release."""

    handle = None

    @property
    def kernel(self):
        return lib().scg_plan_kernel(self.handle).decode()

    def reset(self, stream=None):
        """Empties the plan's own tally (combinations, random barcodes); asynchronous."""
        _check(self.ctx, lib().scg_plan_reset(self.handle, _stream(stream)))

    def harvest(self, as_array=True):
        """The plan's table exactly as the file-level call returns it (synchronises the device)."""
        from .rcpp import _table
        h = C.c_void_p()
        _check(self.ctx, lib().scg_plan_harvest(self.handle, C.byref(h)))
        try:
            return _table(h, self._table_kind if as_array or self._table_kind == "combo" else "random")
        finally:
            lib().scg_result_free(h)

    def free(self):
        if self.handle:
            lib().scg_plan_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DualPlan(_Plan):
    """countDualBarcodes (paired-end) with both templates and the library of pairs compiled onto the device once."""

    def __init__(self, constant1, reverse1, mismatches1, pool1, constant2, reverse2, mismatches2, pool2, randomized, use_first, device=None):
        self.ctx = context(device)
        self.npool = len(pool1)
        a1, k1 = _strs(pool1)
        a2, k2 = _strs(pool2)
        self.handle = C.c_void_p()
        _check(self.ctx, lib().scg_dual_plan_create(self.ctx, constant1.encode("latin-1"), int(bool(reverse1)), int(mismatches1), a1, len(pool1),
                                                    constant2.encode("latin-1"), int(bool(reverse2)), int(mismatches2), a2, len(pool2),
                                                    int(bool(randomized)), int(bool(use_first)), C.byref(self.handle)))

    def run(self, reads1, reads2, counts_ptr, index_ptr=None, stream=None):
        """One pass over the pairs; counts (device int32[npool]) are accumulated into.  Asynchronous."""
        _check(self.ctx, lib().scg_dual_plan_run(self.handle, reads1.handle, reads2.handle, C.c_void_p(counts_ptr),
                                                 C.c_void_p(index_ptr) if index_ptr else None, _stream(stream)))


class ComboPlan(_Plan):
    """countComboBarcodes (single-end, two variable regions); the plan owns the tally of combinations."""
    _table_kind = "combo"

    def __init__(self, constant, strand, pool1, pool2, mismatches, use_first, device=None):
        self.ctx = context(device)
        a1, k1 = _strs(pool1)
        a2, k2 = _strs(pool2)
        self.handle = C.c_void_p()
        _check(self.ctx, lib().scg_combo_plan_create(self.ctx, constant.encode("latin-1"), int(strand), a1, len(pool1), a2, len(pool2),
                                                     int(mismatches), int(bool(use_first)), C.byref(self.handle)))

    def run(self, reads, pairs_ptr=None, stream=None):
        _check(self.ctx, lib().scg_combo_plan_run(self.handle, reads.handle, C.c_void_p(pairs_ptr) if pairs_ptr else None, _stream(stream)))


class RandomPlan(_Plan):
    """countRandomBarcodes; the plan owns the device count table (sized once when expected_distinct > 0)."""
    _table_kind = "random_array"

    def __init__(self, constant, strand, mismatches, use_first, expected_distinct=0, device=None):
        self.ctx = context(device)
        self.handle = C.c_void_p()
        _check(self.ctx, lib().scg_random_plan_create(self.ctx, constant.encode("latin-1"), int(strand), int(mismatches), int(bool(use_first)),
                                                      C.c_longlong(int(expected_distinct)), C.byref(self.handle)))

    def run(self, reads, index_ptr=None, stream=None):
        _check(self.ctx, lib().scg_random_plan_run(self.handle, reads.handle, C.c_void_p(index_ptr) if index_ptr else None, _stream(stream)))


class DeviceArray:
    """A zero-initialised device buffer owned through the C ABI (for callers without torch)."""

    def __init__(self, nbytes, device=None):
        self.ctx = context(device)
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        _check(self.ctx, lib().scg_device_alloc(self.ctx, C.c_size_t(self.nbytes), C.byref(p)))
        self.ptr = p.value

    def zero(self):
        _check(self.ctx, lib().scg_device_zero(self.ctx, C.c_void_p(self.ptr), C.c_size_t(self.nbytes), _stream(None)))

    def to_numpy(self, dtype):
        out = np.zeros(self.nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        _check(self.ctx, lib().scg_device_to_host(self.ctx, _ip(out), C.c_void_p(self.ptr), C.c_size_t(self.nbytes)))
        return out

    def free(self):
        if self.ptr:
            lib().scg_device_free(self.ctx, C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def synchronize(device=None):
    ctx = context(device)
    _check(ctx, lib().scg_synchronize(ctx))
