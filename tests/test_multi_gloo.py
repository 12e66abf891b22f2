"""N > 1 host path on CPU: processes over gloo, reads sharded by contiguous range, results combined with
screencounter_b200.multi (the same code runs over NCCL on the GPUs).  The per-rank counting engine here is the oracle and
the table operations are a numpy stand-in (both test infrastructure), so what is being tested is the sharding and the
exchange logic: dense all-reduce; sparse tables cut at splitters, all-to-all, merge tree, gather (SURVEY.md 8.2 row e,
8.1 T24).  tests/test_gpu_multi.py runs the product's own table kernels."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TEMPLATE = "ACGT" + "-" * 10 + "TGCA"
COMBO_TEMPLATE = "AAAA" + "-" * 6 + "CC" + "-" * 6 + "GGGG"
RANDOM_TEMPLATE = "ACGTACGT" + "-" * 8 + "TGCATGCA"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _workload():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import adversarial_reads, dense_pool, random_seq
    rng = np.random.default_rng(7)
    pool = dense_pool(rng, 60, 10)
    reads = adversarial_reads(rng, 3001, TEMPLATE, [pool], strand="both", read_len=40, sub_rate=0.03, n_rate=0.01,
                              lower_rate=0.0, double_frac=0.05, short_frac=0.02)
    p1, p2 = dense_pool(rng, 12, 6), dense_pool(rng, 9, 6)
    combo_reads = adversarial_reads(rng, 2000, COMBO_TEMPLATE, [p1, p2], strand="both", read_len=40, sub_rate=0.02, n_rate=0.005,
                                    lower_rate=0.0, double_frac=0.0, short_frac=0.02)
    truth = [random_seq(rng, 8) for _ in range(40)]
    rand_reads = adversarial_reads(rng, 2000, RANDOM_TEMPLATE, [truth], strand="both", read_len=40, sub_rate=0.02, n_rate=0.01,
                                   lower_rate=0.0, double_frac=0.0, short_frac=0.02)
    return pool, reads, (p1, p2), combo_reads, rand_reads


class NumpyTableOps:
    """TEST INFRASTRUCTURE: the table operations of screencounter_b200.multi.merge_tables_across_ranks on host tensors, so that
    the exchange logic (splitters, cuts, all-to-all, merge tree, gather) runs over gloo without a GPU.  The product's own
    implementation is DeviceTableOps (kernels behind the C ABI); tests/test_gpu_multi.py runs that one."""

    def views(self, table):
        return table["keys"], table["counts"]

    def key_len(self, table):
        return table["key_len"]

    def from_tensors(self, keys, counts, key_len):
        return {"keys": keys.clone(), "counts": counts.clone(), "key_len": key_len}

    def merge(self, a, b):
        import torch
        keys = torch.cat([a["keys"], b["keys"]])
        counts = torch.cat([a["counts"], b["counts"]]).to(torch.int64)
        uniq, inverse = torch.unique(keys, sorted=True, return_inverse=True)
        total = torch.zeros(uniq.numel(), dtype=torch.int64).index_add_(0, inverse, counts)
        return {"keys": uniq, "counts": total.to(torch.int32), "key_len": a["key_len"]}

    def free(self, table):
        pass


def _combo_table(keys, freq):
    import torch
    keys = np.asarray(keys, dtype=np.int64).reshape(len(freq), 2)
    packed = (keys[:, 0] << 32) | keys[:, 1]
    order = np.argsort(packed, kind="stable")
    return {"keys": torch.from_numpy(packed[order]), "counts": torch.from_numpy(np.asarray(freq, dtype=np.int32)[order]), "key_len": 0}


_RANK = {"A": 0, "C": 1, "G": 2, "N": 3, "T": 4}


def _random_table(seqs, freq, key_len):
    import torch
    codes = np.array([sum(_RANK[ch] << (3 * (key_len - 1 - i)) for i, ch in enumerate(s)) for s in seqs], dtype=np.int64)
    order = np.argsort(codes, kind="stable")
    return {"keys": torch.from_numpy(codes[order]), "counts": torch.from_numpy(np.asarray(freq, dtype=np.int32)[order]), "key_len": key_len}


def _decode_random(keys, key_len):
    return ["".join("ACGNT"[(int(k) >> (3 * (key_len - 1 - i))) & 7] for i in range(key_len)) for k in keys]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import port as engine
    from screencounter_b200 import multi
    from util import fastq
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pool, reads, (p1, p2), combo_reads, rand_reads = _workload()
        ops = NumpyTableOps()
        # dense: single barcodes, one all-reduce of [counts..., total]
        first, count = multi.shard_range(len(reads), rank, world)
        counts, total = engine.count_single(fastq(reads[first:first + count]), TEMPLATE, 2, pool, 1, False)
        counts, (total,) = multi.combine_dense(counts, [total])
        # sparse: combinations -- sorted tables cut at splitters, one all-to-all, merge tree, gather on rank 0
        first, count = multi.shard_range(len(combo_reads), rank, world)
        keys, freq, ctotal = engine.count_combo_single(fastq(combo_reads[first:first + count]), COMBO_TEMPLATE, 2, p1, p2, 1, True)
        merged, info = multi.merge_tables_across_ranks(ops, _combo_table(keys, freq))
        _, (ctotal,) = multi.combine_dense(np.zeros(1, dtype=np.int32), [ctotal])
        # sparse: random barcodes
        first, count = multi.shard_range(len(rand_reads), rank, world)
        seqs, rfreq, rtotal = engine.count_random(fastq(rand_reads[first:first + count]), RANDOM_TEMPLATE, 2, 1, True)
        rmerged, rinfo = multi.merge_tables_across_ranks(ops, _random_table(list(seqs), rfreq, 8))
        _, (rtotal,) = multi.combine_dense(np.zeros(1, dtype=np.int32), [rtotal])
        assert (merged is None) == (rank != 0) and (rmerged is None) == (rank != 0)
        out = dict(counts=counts, total=total, ctotal=ctotal, rtotal=rtotal)
        if rank == 0:
            k = merged["keys"].numpy()
            out.update(keys=np.stack([k >> 32, k & 0xFFFFFFFF], axis=1).astype(np.int32), freq=merged["counts"].numpy(),
                       seqs=np.array(_decode_random(rmerged["keys"].numpy(), 8)), rfreq=rmerged["counts"].numpy(),
                       rows_total=info["rows_total"])
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **out)
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    from screencounter_b200 import multi
    for n in (0, 1, 7, 100, 12345):
        for world in (1, 2, 3, 8):
            spans = [multi.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0
            assert sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1


def test_single_rank_is_identity():
    from screencounter_b200 import multi
    counts = np.arange(5, dtype=np.int32)
    out, scalars = multi.combine_dense(counts, [3])
    assert np.array_equal(out, counts) and scalars == [3]


def test_numpy_table_ops_merge():
    ops = NumpyTableOps()
    a = _combo_table([[0, 5], [1, 2]], [2, 1])
    b = _combo_table([[1, 2], [3, 0]], [3, 4])
    m = ops.merge(a, b)
    assert m["keys"].tolist() == [5, (1 << 32) | 2, 3 << 32] and m["counts"].tolist() == [2, 4, 4]
    t = _random_table(["T", "N", "A"], [1, 2, 3], 1)
    assert _decode_random(t["keys"].numpy(), 1) == ["A", "N", "T"] and t["counts"].tolist() == [3, 2, 1]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_ranks_gloo_match_unsharded(port, tmp_path, world):
    import torch.multiprocessing as mp
    from util import fastq
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    pool, reads, (p1, p2), combo_reads, rand_reads = _workload()
    want_counts, want_total = port.count_single(fastq(reads), TEMPLATE, 2, pool, 1, False)
    want_keys, want_freq, want_ctotal = port.count_combo_single(fastq(combo_reads), COMBO_TEMPLATE, 2, p1, p2, 1, True)
    want_seqs, want_rfreq, want_rtotal = port.count_random(fastq(rand_reads), RANDOM_TEMPLATE, 2, 1, True)
    order = np.argsort(np.array(list(want_seqs)))
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert np.array_equal(got["counts"], want_counts)
        assert int(got["total"]) == want_total == len(reads)
        assert int(got["ctotal"]) == want_ctotal
        assert int(got["rtotal"]) == want_rtotal
        if rank == 0:
            assert np.array_equal(got["keys"], np.asarray(want_keys).reshape(len(want_freq), -1))
            assert np.array_equal(got["freq"], want_freq)
            assert int(got["rows_total"]) == len(want_freq)
            assert got["seqs"].tolist() == [list(want_seqs)[i] for i in order]
            assert np.array_equal(got["rfreq"], np.asarray(want_rfreq)[order])
