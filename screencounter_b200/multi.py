"""Multi-GPU plumbing: one process per GPU, reads sharded by contiguous range, results combined
with ONE collective (SURVEY.md 8.2 row e).

The reference merges per-thread results in `reduce()` -- a vector add for the dense counters
(handlers/SingleBarcodeSingleEnd.hpp:119-125, handlers/DualBarcodesPairedEnd.hpp:202-208), an
append + sort for combinations (handlers/CombinatorialBarcodesSingleEnd.hpp:268-305), a map merge
for random barcodes (handlers/RandomBarcodeSingleEnd.hpp:197-207).  Across GPUs the same merges
are: one all-reduce (sum) of [counts..., total, extras...] for the dense outputs, and an
all-gather of the per-rank reduced (key, freq) tables followed by a merge by key for the sparse
ones.  Results do not depend on how reads are split.

Everything here works on whatever backend the process group has: NCCL over NVLink on the GPUs
(tensors on the rank's device), gloo in the CPU tests.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous range [first, first + count) of `n` reads (pairs) owned by `rank` of `world`."""
    first = (n * rank) // world
    last = (n * (rank + 1)) // world
    return first, last - first


def _dist():
    import torch.distributed as dist
    return dist


def _device_for(group=None):
    import torch
    dist = _dist()
    backend = dist.get_backend(group)
    if "nccl" in str(backend):
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def combine_dense(counts, scalars=(), group=None):
    """Sum `counts` (int32 vector) and the integer `scalars` (total, barcode1_only, ...) over all
    ranks with a single all-reduce of one packed buffer.  Returns (counts, [scalars...]).

    `counts` may be a numpy array or a torch tensor (a CUDA tensor stays on the device and the
    all-reduce runs over NCCL); the scalars ride in the same buffer so there is exactly one
    collective per call."""
    import torch
    dist = _dist()
    scalars = [int(s) for s in scalars]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return counts, scalars
    dev = _device_for(group)
    is_tensor = isinstance(counts, torch.Tensor)
    body = counts.to(device=dev, dtype=torch.int64) if is_tensor else torch.from_numpy(np.asarray(counts, dtype=np.int64)).to(dev)
    # int64 on the wire: per-rank int32 counters cannot overflow the sum of up to 8 ranks
    packed = torch.cat([body.reshape(-1), torch.tensor(scalars, dtype=torch.int64, device=dev)])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    n = body.numel()
    if int(packed.max().item()) > np.iinfo(np.int32).max:
        raise OverflowError("combined count exceeds the reference's 32-bit counters (SURVEY.md 8.1 T17)")
    out_scalars = [int(v) for v in packed[n:].tolist()]
    if is_tensor:
        return packed[:n].to(dtype=torch.int32).reshape(counts.shape), out_scalars
    return packed[:n].cpu().numpy().astype(np.int32).reshape(np.asarray(counts).shape), out_scalars


def _merge_int_keys(keys, freq):
    """Sum frequencies of identical key rows; rows come back sorted ascending (first column major):
    the order of sort_combinations + count_combinations (utils.hpp:173-198, src/utils.h:14-45)."""
    keys = np.asarray(keys, dtype=np.int32).reshape(len(freq), -1)
    if len(freq) == 0:
        return keys, np.zeros(0, dtype=np.int32)
    uniq, inverse = np.unique(keys, axis=0, return_inverse=True)
    total = np.zeros(len(uniq), dtype=np.int64)
    np.add.at(total, inverse.reshape(-1), np.asarray(freq, dtype=np.int64))
    return uniq.astype(np.int32), total.astype(np.int32)


def _merge_string_keys(seqs, freq):
    """Sum frequencies of identical sequences; sorted like R's order() on upper-case ACGTN strings
    (A < C < G < N < T, R/countRandomBarcodes.R:73), other strings in plain byte order after them."""
    acc = {}
    for s, f in zip(seqs, freq):
        acc[s] = acc.get(s, 0) + int(f)
    out = sorted(acc, key=lambda s: s.encode("latin-1"))
    return out, np.array([acc[s] for s in out], dtype=np.int32)


def combine_table(keys, freq, scalars=(), group=None):
    """Merge per-rank (key, freq) tables by key over all ranks.  `keys` is an int32 matrix with one
    combination per ROW, or a list of strings (random barcodes).  One all-gather carries every rank's
    table; the integer `scalars` are summed.  Returns (keys, freq, [scalars...]) identical on all ranks."""
    dist = _dist()
    strings = isinstance(keys, (list, tuple))
    scalars = [int(s) for s in scalars]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if strings:
            k, f = _merge_string_keys(keys, freq)
        else:
            k, f = _merge_int_keys(keys, freq)
        return k, f, scalars
    world = dist.get_world_size(group)
    payload = (list(keys) if strings else np.asarray(keys, dtype=np.int32), np.asarray(freq, dtype=np.int32), scalars)
    gathered = [None] * world
    dist.all_gather_object(gathered, payload, group=group)
    all_freq = np.concatenate([g[1] for g in gathered]) if gathered else np.zeros(0, dtype=np.int32)
    summed = [sum(g[2][i] for g in gathered) for i in range(len(scalars))]
    if strings:
        all_keys = [s for g in gathered for s in g[0]]
        k, f = _merge_string_keys(all_keys, all_freq)
    else:
        width = max([np.asarray(g[0]).reshape(len(g[1]), -1).shape[1] for g in gathered if len(g[1])] or [2])
        all_keys = np.concatenate([np.asarray(g[0], dtype=np.int32).reshape(len(g[1]), width) for g in gathered])
        k, f = _merge_int_keys(all_keys, all_freq)
    return k, f, summed


def count_single_barcodes_sharded(fastq_shard, constant, strand, pool, mismatches, use_first, nthreads=1, group=None, engine=None):
    """countSingleBarcodes over a file split across ranks: each rank passes ITS contiguous part of
    the reads; every rank gets the whole-file (counts, total).  `engine` defaults to the CUDA path."""
    if engine is None:
        from . import rcpp
        counts, total = rcpp.count_single_barcodes(fastq_shard, constant, strand, pool, mismatches, use_first, nthreads)
    else:
        counts, total = engine.count_single(fastq_shard, constant, strand, pool, mismatches, use_first)
    counts, (total,) = combine_dense(counts, [total], group=group)
    return counts, total
