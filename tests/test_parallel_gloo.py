"""N>1 host logic on CPU: world_size-2 gloo process group (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marl_dpp_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        q = torch.randn(257, 8, generator=g)
        sync = parallel.ReplicaSync(interval=4)
        history = []
        for rounds_done in range(1, 13):
            q += 0.01 * (rank + 1)              # stands in for a round of local learning
            if sync.step(q, rounds_done):
                history.append(rounds_done)
        vis = torch.tensor([1 << rank, 0, 3 * rank, 16], dtype=torch.int32)   # updated-key bits of this rank
        parallel.merge_visited_(vis)
        np.save(os.path.join(out_dir, "vis%d.npy" % rank), vis.numpy())
        base, n = parallel.shard_env_ids(1 << 10, rank, world)
        np.save(os.path.join(out_dir, "q%d.npy" % rank), q.numpy())
        np.save(os.path.join(out_dir, "meta%d.npy" % rank), np.array(history + [base, n, sync.syncs]))
    finally:
        dist.destroy_process_group()


def test_replica_averaging_two_ranks(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    q = [np.load(tmp_path / ("q%d.npy" % r)) for r in range(world)]
    meta = [np.load(tmp_path / ("meta%d.npy" % r)) for r in range(world)]
    # NumPy restatement of the same schedule: average of independently updated replicas
    ref = [torch.randn(257, 8, generator=torch.Generator().manual_seed(100 + r)).numpy() for r in range(world)]
    for rounds_done in range(1, 13):
        for r in range(world):
            ref[r] = ref[r] + np.float32(0.01 * (r + 1))
        if rounds_done % 4 == 0:
            mean = (ref[0] + ref[1]) / np.float32(2)
            ref = [mean.copy() for _ in range(world)]
    for r in range(world):
        np.testing.assert_allclose(q[r], ref[r], rtol=1e-6, atol=1e-6)
        assert list(meta[r][:3]) == [4, 8, 12] and meta[r][-1] == 3
        assert (meta[r][3], meta[r][4]) == (r * 1024, 1024)
    np.testing.assert_array_equal(q[0], q[1])
    for r in range(world):   # bitwise OR over the ranks
        assert np.load(tmp_path / ("vis%d.npy" % r)).tolist() == [3, 0, 3, 16]


def test_env_id_partitions_cover_exactly_once():
    for total, world in ((1 << 20, 8), (1000, 3), (7, 8)):
        spans = [parallel.split_env_ids(total, r, world) for r in range(world)]
        ids = [i for s, n in spans for i in range(s, s + n)] if total <= 1000 else None
        assert sum(n for _, n in spans) == total
        assert spans[0][0] == 0 and all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        if ids is not None:
            assert ids == list(range(total))
    with pytest.raises(ValueError):
        parallel.shard_env_ids(10, 3, 2)
    assert not parallel.ReplicaSync(0).due(8) and parallel.ReplicaSync(4).due(8) and not parallel.ReplicaSync(4).due(6)


# ---- sharded reachability: the exchange logic on a toy monotone system (the sweeps themselves are CUDA) ----------
def _toy_system(n=3000, seed=5):
    """State i is live iff it is a goal or one of its successors is: random forward edges, a few goals."""
    rng = np.random.default_rng(seed)
    succ = [rng.integers(0, n, size=rng.integers(0, 4)) for _ in range(n)]
    goal = rng.random(n) < 0.01
    return succ, goal


def _toy_sweep(bits, succ, goal, lo, hi):
    """One in-place sweep over states [lo, hi): reads any bit, sets own bits; returns whether one was set."""
    changed = False
    for i in range(lo, min(hi, len(succ))):
        if (int(bits[i >> 5]) >> (i & 31)) & 1:
            continue
        if goal[i] or any((int(bits[j >> 5]) >> (j & 31)) & 1 for j in succ[i]):
            bits[i >> 5] = np.int32(np.uint32(int(bits[i >> 5]) & 0xFFFFFFFF | (1 << (i & 31))).astype(np.int32))
            changed = True
    return changed


def _reach_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        succ, goal = _toy_system()
        words = (len(succ) + 31) // 32
        first, per = parallel.shard_words(words, rank, world)
        bitmap = torch.zeros(per * world, dtype=torch.int32)
        view = bitmap.numpy()
        sweeps = parallel.sharded_fixpoint(
            bitmap, lambda bm: _toy_sweep(view, succ, goal, first * 32, (first + per) * 32), rank, world)
        np.save(os.path.join(out_dir, "reach%d.npy" % rank), np.concatenate([view[:words], [sweeps]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_fixpoint_equals_the_single_process_fixpoint(tmp_path, world):
    port = _free_port()
    mp.spawn(_reach_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    succ, goal = _toy_system()
    words = (len(succ) + 31) // 32
    want = np.zeros(words, dtype=np.int32)
    while _toy_sweep(want, succ, goal, 0, len(succ)):
        pass
    got = [np.load(tmp_path / ("reach%d.npy" % r)) for r in range(world)]
    for r in range(world):
        np.testing.assert_array_equal(got[r][:words], want)      # every rank ends with the whole bitmap
        assert got[r][-1] == got[0][-1] >= 2                       # the same number of sweeps everywhere
    assert 0 < int(np.unpackbits(want.view(np.uint8)).sum()) < len(succ)
    for words_, world_ in ((10, 3), (7, 8), (4096, 2)):
        spans = [parallel.shard_words(words_, r, world_) for r in range(world_)]
        assert spans[0][0] == 0 and all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world_ - 1))
        assert spans[-1][0] + spans[-1][1] >= words_
