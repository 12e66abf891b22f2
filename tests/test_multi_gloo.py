"""N > 1 host path on CPU: two processes over gloo, reads sharded by contiguous range, results
combined with screencounter_b200.multi (the same code runs over NCCL on the GPUs).  The per-rank
counting engine here is the oracle (test infrastructure), so what is being tested is the sharding
and the merges: dense all-reduce, sparse merge by key (SURVEY.md 8.2 row e, 8.1 T24)."""
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
        # dense: single barcodes
        first, count = multi.shard_range(len(reads), rank, world)
        counts, total = multi.count_single_barcodes_sharded(fastq(reads[first:first + count]), TEMPLATE, 2, pool, 1, False, engine=engine)
        # sparse: combinations (merge by key)
        first, count = multi.shard_range(len(combo_reads), rank, world)
        keys, freq, ctotal = engine.count_combo_single(fastq(combo_reads[first:first + count]), COMBO_TEMPLATE, 2, p1, p2, 1, True)
        keys, freq, (ctotal,) = multi.combine_table(keys, freq, [ctotal])
        # sparse: random barcodes (merge by string)
        first, count = multi.shard_range(len(rand_reads), rank, world)
        seqs, rfreq, rtotal = engine.count_random(fastq(rand_reads[first:first + count]), RANDOM_TEMPLATE, 2, 1, True)
        seqs, rfreq, (rtotal,) = multi.combine_table(list(seqs), rfreq, [rtotal])
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), counts=counts, total=total, keys=keys, freq=freq, ctotal=ctotal,
                 seqs=np.array(seqs), rfreq=rfreq, rtotal=rtotal)
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
    keys, freq, _ = multi.combine_table(np.array([[1, 2], [0, 5], [1, 2]], dtype=np.int32), np.array([1, 2, 3], dtype=np.int32))
    assert keys.tolist() == [[0, 5], [1, 2]] and freq.tolist() == [2, 4]
    seqs, freq, _ = multi.combine_table(["T", "AN", "AC", "T"], [1, 1, 1, 4])
    assert seqs == ["AC", "AN", "T"] and freq.tolist() == [1, 1, 5]


@pytest.mark.timeout(300)
def test_two_ranks_gloo_match_unsharded(port, tmp_path):
    import torch.multiprocessing as mp
    from util import fastq
    world = 2
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
        assert np.array_equal(got["keys"], np.asarray(want_keys).reshape(len(want_freq), -1))
        assert np.array_equal(got["freq"], want_freq)
        assert int(got["ctotal"]) == want_ctotal
        assert got["seqs"].tolist() == [list(want_seqs)[i] for i in order]
        assert np.array_equal(got["rfreq"], np.asarray(want_rfreq)[order])
        assert int(got["rtotal"]) == want_rtotal
