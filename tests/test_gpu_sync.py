"""GPU: Q-replica averaging by the library's own peer-memory kernel (csrc/q_sync.cuh).

No reference counterpart, so the oracle is the NumPy restatement of the definition: float32 sum of the
ranks' tables in rank order, divided by the rank count -- the same bits on every rank.

Kernels that wait on one another must never be separate launches on one GPU, so with ONE GPU:
  * the protocol between ranks (both handshakes, epochs, slice ownership, arithmetic) runs as the
    library's cooperative group launch, every rank's slices with the code its own kernel runs
    (mdpp_sync_emulate);
  * the exchange INSIDE mdpp_q_train_rounds (averaging points, multi-round launches that stop at them,
    the copy home of an odd ping-pong parity, the overlap with the next env kernel) runs with real
    rank kernels against peers whose flags are saturated, i.e. peers that never have to be waited for.
One process per GPU with CUDA IPC between them runs when the box has 2+ GPUs.
"""
import os
import socket

import numpy as np
import pytest

from conftest import bad_states_text

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _mods():
    import marl_dpp_b200 as m
    from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner
    from marl_dpp_b200.parallel import PeerExchange
    return m, BatchedEnv, BatchedQLearner, PeerExchange


def _mean_in_rank_order(tables):
    s = tables[0].astype(np.float32).copy()
    for t in tables[1:]:
        s = (s + t).astype(np.float32)
    return (s / np.float32(len(tables))).astype(np.float32)


def _scenario(m, block=2, cap=300):
    boxes, idx, choose = m.layout_3[block]["initial_states"][0]
    return m.Scenario(m.layout_3[block], boxes, idx, choose, bad_states_text=bad_states_text(block), step_cap=cap)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_exchange_protocol_equals_mean_in_rank_order(world):
    """`world` replicas trained on different env shards, then two exchanges (one cooperative launch each): every
    table equals the float32 mean in rank order, bit for bit and on every rank; the second exchange (equal
    inputs, next epoch of the flags) is the identity."""
    m, BatchedEnv, BatchedQLearner, PeerExchange = _mods()
    sc = _scenario(m)
    n = 2048
    ex = PeerExchange.local(sc, world)
    lrs = []
    for r in range(world):
        env = BatchedEnv(sc, n, env_id_base=r * n, seed=3)
        lr = BatchedQLearner(env, episodes=5000, seed=3, q=ex.tables[r])
        ex.attach(lr, 0, rank=r)
        lr.train_rounds(6 + r)                  # different amounts of learning per rank
        lrs.append(lr)
    torch.cuda.synchronize()
    before = [t.cpu().numpy().copy() for t in ex.tables]
    assert any(not np.array_equal(before[0], b) for b in before[1:])
    for rep in range(2):
        ex.emulate(lrs)
        torch.cuda.synchronize()
        want = _mean_in_rank_order(before)
        for r in range(world):
            assert np.array_equal(ex.tables[r].cpu().numpy(), want), (rep, r)
        before = [want] * world
    for r, lr in enumerate(lrs):
        assert lr.env.stats()["internal_errors"] == 0 and lr.sync_count() == 2
        fl = ex.flags(r).cpu().numpy().reshape(2, 148, 8)
        ctas = (sc.n_states * 2 + 127) // 128          # 128 float4 per CTA while the table fits 148 CTAs that way
        assert (fl[:, :ctas, :world] == 2).all() and (fl[:, ctas:, :] == 0).all()   # both phases, every slice, epoch 2
    ex2 = PeerExchange.local(sc, 2)
    with pytest.raises(ValueError):                 # a learner whose table is not the exchange's
        ex2.attach(BatchedQLearner(BatchedEnv(sc, 64)), 4, rank=0)
    del lrs
    ex.close()
    ex2.close()


@pytest.mark.parametrize("block,interval,chunks,rank,world",
                         [(2, 4, (10, 1, 5), 0, 2), (0, 3, (7, 4), 1, 3), (2, 1, (6,), 2, 3), (0, 1, (5,), 0, 2),
                          (2, 64, (64, 64, 8), 1, 2)])
def test_interval_exchange_inside_train_rounds(block, interval, chunks, rank, world):
    """mdpp_q_train_rounds averages before every round R > 0 with R % interval == 0 as a member of its kernel
    chain (multi-round launches stop at the averaging points; the 1-car table is copied home when its ping-pong
    parity is odd; the next env kernel is released while the exchange runs).  One real rank at the BASELINE-like
    launch shape (148 CTAs) among peers that hold constant tables and never have to be waited for, against a
    plain handle trained round by round with the NumPy mean applied at the same rounds: bit-identical."""
    m, BatchedEnv, BatchedQLearner, PeerExchange = _mods()
    sc = _scenario(m, block)
    n = 1 << 15
    ex = PeerExchange.local(sc, world)
    g = torch.Generator().manual_seed(11)
    for r in range(world):
        ex.flags(r).fill_(0x7FFFFFFF)              # saturated: no rank ever waits for another
        if r != rank:
            t = torch.randn((sc.n_states, 8), generator=g) * 50.0
            t[:, 5:] = 0.0
            ex.tables[r].copy_(t)
    env = BatchedEnv(sc, n, env_id_base=rank * n, seed=5)
    lr = BatchedQLearner(env, episodes=5000, seed=5, q=ex.tables[rank])
    ex.attach(lr, interval, rank=rank)
    for k in chunks:
        lr.train_rounds(k)
    torch.cuda.synchronize()
    total = sum(chunks)
    ref = BatchedQLearner(BatchedEnv(sc, n, env_id_base=rank * n, seed=5), episodes=5000, seed=5)
    peers = [t.cpu().numpy() for t in ex.tables]
    n_sync = 0
    for R in range(total):
        if R > 0 and R % interval == 0:
            tabs = [ref.q.cpu().numpy() if r == rank else peers[r] for r in range(world)]
            ref.q.copy_(torch.from_numpy(_mean_in_rank_order(tabs)).cuda())
            n_sync += 1
        ref.train_rounds(1)
    torch.cuda.synchronize()
    assert torch.equal(lr.q, ref.q)
    assert torch.equal(lr.env.records, ref.env.records)
    a, b = lr.env.stats(), ref.env.stats()
    assert a["internal_errors"] == 0
    for key in ("agent_steps", "q_updates", "episodes_done", "touched_keys"):
        assert a[key] == b[key], key
    assert lr.sync_count() == n_sync == (total - 1) // interval
    for r in range(world):                          # the peers' tables were only read
        if r != rank:
            assert np.array_equal(ex.tables[r].cpu().numpy(), peers[r])
    del lr, ref
    ex.close()


# ---- one process per GPU, buffers exchanged as CUDA IPC handles (needs 2+ GPUs) -------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ipc_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        m, BatchedEnv, BatchedQLearner, PeerExchange = _mods()
        sc = _scenario(m)
        n = 1 << 16
        ex = PeerExchange(sc)
        env = BatchedEnv(sc, n, env_id_base=rank * n, seed=7)
        lr = BatchedQLearner(env, episodes=5000, seed=7, q=ex.q)
        ex.attach(lr, 4)
        lr.train_rounds(10)
        lr.train_rounds(3)
        lr.sync_replicas()
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, "q%d.npy" % rank), lr.q.cpu().numpy())
        np.save(os.path.join(out_dir, "err%d.npy" % rank), np.array([env.stats()["internal_errors"], lr.sync_count()]))
        dist.barrier()
        del lr, env
        ex.close()
    finally:
        dist.destroy_process_group()


def test_ipc_exchange_two_processes(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_ipc_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    q = [np.load(tmp_path / ("q%d.npy" % r)) for r in range(world)]
    assert np.array_equal(q[0], q[1])                              # replicas bit-equal after the exchange
    for r in range(world):
        assert np.load(tmp_path / ("err%d.npy" % r)).tolist() == [0, 4]   # rounds 4, 8, 12 + the explicit one
    # the same two shards in this process on plain handles, NumPy mean at rounds 4, 8, 12 and at the end
    m, BatchedEnv, BatchedQLearner, _ = _mods()
    sc = _scenario(m)
    n = 1 << 16
    refs = [BatchedQLearner(BatchedEnv(sc, n, env_id_base=r * n, seed=7), episodes=5000, seed=7) for r in range(world)]
    for R in range(13):
        if R > 0 and R % 4 == 0:
            mean = torch.from_numpy(_mean_in_rank_order([x.q.cpu().numpy() for x in refs])).cuda()
            for x in refs:
                x.q.copy_(mean)
        for x in refs:
            x.train_rounds(1)
    want = _mean_in_rank_order([x.q.cpu().numpy() for x in refs])
    assert np.array_equal(q[0], want)
