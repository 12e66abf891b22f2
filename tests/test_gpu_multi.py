"""The device side of the multi-GPU exchange: sorted tables of sparse results, their merge kernels and their rendering,
through the C ABI (scg_table_*).  One GPU is enough for the table operations (two plans stand in for two ranks); the
NCCL round trip itself runs where two GPUs are visible, and in every N > 1 bench run (bench.py `nccl_check`)."""
import os
import socket
import sys

import numpy as np
import pytest

from util import distinct_pool

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RANDOM_TEMPLATE = "CAGCTACGTACG" + "-" * 16 + "CCAGCTCGATCG"
COMBO_TEMPLATE = "CAGCTACG" + "-" * 20 + "GGTACCTT" + "-" * 20 + "CGATCGAG"


def _random_spec(space=30_000):
    from screencounter_b200.device import SynthSpec
    return SynthSpec(RANDOM_TEMPLATE, [], seed=13, read_len=75, strand=2, random_space=space, sub_per_10k=30)


def test_random_tables_merge_like_one_pass():
    from screencounter_b200.device import RandomPlan
    from screencounter_b200.multi import DeviceTableOps
    spec = _random_spec()
    spans = [(0, 200_000), (200_000, 150_003), (350_003, 1), (350_004, 99_996)]
    ops = DeviceTableOps()
    tables = []
    for first, n in spans:
        plan = RandomPlan(RANDOM_TEMPLATE, 2, 1, True)
        plan.run(spec.on_device(first, n))
        tables.append(ops.from_plan(plan))
        plan.free()
    merged = tables[0]
    for t in tables[1:]:
        nxt = ops.merge(merged, t)
        ops.free(merged)
        ops.free(t)
        merged = nxt
    seqs, freq = ops.render(merged)
    whole = RandomPlan(RANDOM_TEMPLATE, 2, 1, True)
    whole.run(spec.on_device(0, 450_000))
    want_seqs, want_freq = whole.harvest()
    assert np.array_equal(seqs, want_seqs) and np.array_equal(freq, want_freq)
    assert list(seqs) == sorted(seqs)
    # merging with an empty table and with itself
    empty_plan = RandomPlan(RANDOM_TEMPLATE, 2, 1, True)
    empty = ops.from_plan(empty_plan)
    same = ops.merge(merged, empty)
    s2, f2 = ops.render(same)
    assert np.array_equal(s2, seqs) and np.array_equal(f2, freq)
    twice = ops.merge(merged, same)
    s3, f3 = ops.render(twice)
    assert np.array_equal(s3, seqs) and np.array_equal(f3, 2 * freq)
    other = ops.merge(empty, merged)
    s4, f4 = ops.render(other)
    assert np.array_equal(s4, seqs) and np.array_equal(f4, freq)


@pytest.mark.parametrize("npool", [300, 5000])   # dense matrix / device hash
def test_combo_tables_merge_like_one_pass(npool):
    from screencounter_b200.device import SynthSpec, ComboPlan
    from screencounter_b200.multi import DeviceTableOps
    rng = np.random.default_rng(8)
    p1, p2 = distinct_pool(rng, npool, 20), distinct_pool(rng, npool, 20)
    spec = SynthSpec(COMBO_TEMPLATE, [p1, p2], seed=11, read_len=75, strand=2)
    ops = DeviceTableOps()
    parts = []
    for first, n in ((0, 120_000), (120_000, 80_001)):
        plan = ComboPlan(COMBO_TEMPLATE, 2, p1, p2, 1, True)
        plan.run(spec.on_device(first, n))
        parts.append(ops.from_plan(plan))
        plan.free()
    merged = ops.merge(parts[0], parts[1])
    keys, freq = ops.render(merged)
    whole = ComboPlan(COMBO_TEMPLATE, 2, p1, p2, 1, True)
    whole.run(spec.on_device(0, 200_001))
    want_keys, want_freq = whole.harvest()
    assert np.array_equal(keys, want_keys) and np.array_equal(freq, want_freq)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from screencounter_b200 import multi
    from screencounter_b200.device import RandomPlan
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        spec = _random_spec()
        first, n = multi.shard_range(450_000, rank, world)
        plan = RandomPlan(RANDOM_TEMPLATE, 2, 1, True, device=rank)
        plan.run(spec.on_device(first, n, device=rank))
        ops = multi.DeviceTableOps(rank)
        merged, info = multi.merge_tables_across_ranks(ops, ops.from_plan(plan))
        if rank == 0:
            seqs, freq = ops.render(merged)
            np.savez(os.path.join(out_dir, "merged.npz"), seqs=seqs, freq=freq, rows_total=info["rows_total"])
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpus_nccl_merge(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (bench.py checks the same on every N > 1 run)")
    import torch.multiprocessing as mp
    from screencounter_b200.device import RandomPlan
    mp.spawn(_nccl_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "merged.npz"))
    whole = RandomPlan(RANDOM_TEMPLATE, 2, 1, True)
    whole.run(_random_spec().on_device(0, 450_000))
    want_seqs, want_freq = whole.harvest()
    assert np.array_equal(got["seqs"], want_seqs) and np.array_equal(got["freq"], want_freq)
    assert int(got["rows_total"]) == len(want_freq)
