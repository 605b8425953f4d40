"""GPU: -dd reachability bitmap (new definition) bit-exact against the CPU brute force
built from the reference-pinned scalar transition functions."""
import numpy as np
import pytest

from conftest import bad_states_text

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CASES = [
    (2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]]),
    (2, ["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]]),
    (2, ["D5", "D2", "D6"], [0, 0, 0], [[2, 1, 0], [2, 1, 0], [0, 2, 0]]),
    (0, ["U4", "U8", "U7"], [0, 1, 1], [[0, 0, 0]] * 3),
    (0, ["U4"], [1], [[0, 0, 0]]),
]


@pytest.mark.parametrize("block,boxes,idx,choose", CASES)
def test_reachability_bitmap_bit_exact(block, boxes, idx, choose):
    import marl_dpp_b200 as m
    from marl_dpp_b200.batched import BatchedEnv
    from oracle import oracle as orc
    sc = m.Scenario(m.layout_3[block], boxes, idx, choose)
    env = BatchedEnv(sc, 32)
    want, live = orc.reachability(orc.Config(m.layout_3[block]), len(boxes), choose, sc.n_local_nodes, sc.node_base)
    # both implementations: the backward breadth-first search (one cooperative launch) and the forward sweeps
    for method in ("bfs", "sweep", "bfs"):
        bitmap, sweeps = env.deadlock_reachability(method=method)
        got = bitmap.cpu().numpy().view(np.uint32)
        assert np.array_equal(got, want), method
        assert int(sum(bin(int(w)).count("1") for w in got)) == live and sweeps >= 2
    V = 2 * sc.n_local_nodes + 1
    goal = V ** len(boxes) - 1
    assert (int(got[goal >> 5]) >> (goal & 31)) & 1
    # every start heading of the scenario is live (the task is solvable)
    for h in "NESW":
        j = env.joint_index([(b + h, d) for b, d in zip(boxes, idx)])
        assert (int(got[j >> 5]) >> (j & 31)) & 1


def test_reachability_four_cars_properties_and_bad_states_cross_report():
    """4 cars (28.4 M joint states): fixpoint reached, deterministic, and a cross-report against the
    reference's shipped bad-states file (informative: the two definitions differ)."""
    import marl_dpp_b200 as m
    from marl_dpp_b200.batched import BatchedEnv
    boxes, idx = ["D2", "D6", "D4", "D8"], [0, 0, 0, 0]
    choose = [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]]
    sc = m.Scenario(m.layout_3[2], boxes, idx, choose)
    env = BatchedEnv(sc, 32)
    b1, s1 = env.deadlock_reachability(method="sweep")
    b2, s2 = env.deadlock_reachability(method="bfs")
    assert torch.equal(b1, b2)       # the fixpoint is unique: forward sweeps == backward search
    assert 2 <= s1 <= 200 and 2 <= s2 <= 200
    words = b1.cpu().numpy().view(np.uint32)
    live = int(np.unpackbits(words.view(np.uint8)).sum())
    n = (2 * sc.n_local_nodes + 1) ** 4
    assert 0 < live < n
    j = env.joint_index([(b + "N", d) for b, d in zip(boxes, idx)])
    assert (int(words[j >> 5]) >> (j & 31)) & 1
    # the file's canonical 2-car deadlock "D5xD6zD6xD8": car A in D5 bound for D6, car B in D6 bound for D8
    text = bad_states_text(2)
    assert "D5xD6zD6xD8" in text.split("\n")


# ---- the search sharded over ranks: every rank sweeps its own word range, the ranges are exchanged ----
@pytest.mark.parametrize("shards", [1, 3])
def test_sweep_ranges_reach_the_same_fixpoint(shards):
    """mdpp_deadlock_sweep_range on one GPU: the joint space cut into word ranges swept one after the other (what
    the ranks of parallel.sharded_reachability do side by side) ends in the bitmap of the backward search."""
    import marl_dpp_b200 as m
    from marl_dpp_b200 import parallel
    from marl_dpp_b200.batched import BatchedEnv
    block, boxes, idx, choose = CASES[1]
    sc = m.Scenario(m.layout_3[block], boxes, idx, choose)
    env = BatchedEnv(sc, 32)
    want, _ = env.deadlock_reachability(method="bfs")
    n = env.joint_state_count()
    words = (n + 31) // 32
    per = parallel.shard_words(words, 0, shards)[1]
    bitmap = torch.zeros(per * shards, dtype=torch.int32, device=env.device)
    changed = torch.zeros(1, dtype=torch.int32, device=env.device)
    sweeps = 0
    while True:
        changed.zero_()
        for r in range(shards):
            env.deadlock_sweep_range(bitmap, r * per * 32, per * 32, changed)
        sweeps += 1
        if int(changed.item()) == 0:
            break
        assert sweeps < 1000
    assert torch.equal(bitmap[:words], want) and sweeps >= 2
    # a range that does not start on a word boundary is refused
    with pytest.raises(Exception):
        env.deadlock_sweep_range(bitmap, 7, 64, changed)
    # and the one-rank form of the sharded driver (no process group)
    got, s1 = parallel.sharded_reachability(env)
    assert torch.equal(got, want) and s1 >= 2


def _reach_worker(rank, world, port, out_dir):
    import os
    import torch.distributed as dist
    import marl_dpp_b200 as m
    from marl_dpp_b200 import parallel
    from marl_dpp_b200.batched import BatchedEnv
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        boxes, idx = ["D2", "D6", "D4", "D8"], [0, 0, 0, 0]
        choose = [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]]
        env = BatchedEnv(m.Scenario(m.layout_3[2], boxes, idx, choose), 32)
        got, sweeps = parallel.sharded_reachability(env)
        want, _ = env.deadlock_reachability(method="bfs")
        np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([int(torch.equal(got, want)), sweeps]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_reachability_two_gpus(tmp_path):
    """4 cars (28.4 M joint states) with the joint-state range split over two GPUs (NCCL all-gather of the ranks'
    word ranges after every sweep): every rank ends with the bitmap of the single-GPU search."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_reach_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [np.load(tmp_path / ("ok%d.npy" % r)).tolist() for r in range(2)]
    assert res[0][0] == 1 and res[1][0] == 1 and res[0][1] == res[1][1] >= 2
