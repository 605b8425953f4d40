"""GPU parity at the BASELINE batch size: 2^20 envs per GPU, compared with the CPU oracle bit for bit.

The small-batch parity tests give every CTA at most one tile of 32 envs.  At 2^20 envs the persistent
kernels run 148 CTAs x 32 warps with ~7 tiles per warp, the multi-lap tile loops with their record
prefetch, the tiles_q / tiles_r split and the OR over 148 per-CTA bitmap slices all do real work, so a
skipped tile or a dropped slice would change the table, the records or the counters compared here.
The oracle's per-env phase runs on every host core (oracle.set_threads) and finishes in seconds.
"""
import os

import numpy as np
import pytest

from conftest import bad_states_text

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

A5 = "SLRFT"
N_FULL = 1 << 20


def _mods():
    import marl_dpp_b200 as m
    from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner
    from oracle import oracle as orc
    orc.set_threads()
    return m, BatchedEnv, BatchedQLearner, orc


def _assert_records_equal(env, ob):
    """Every env: car fields, ghost counters, episode and step counters."""
    cars = {k: v.cpu().numpy() for k, v in env.cars().items()}
    ocars, oep, osteps = ob.export()
    assert np.array_equal(cars["node"], ocars[:, :, 0])
    assert np.array_equal(cars["dest_index"], ocars[:, :, 1])
    assert np.array_equal(cars["done"], ocars[:, :, 2])
    want_ghost = np.where(ocars[:, :, 2] == 0, 5, np.where(ocars[:, :, 3] >= 0, ocars[:, :, 3], 7))
    assert np.array_equal(cars["ghost"], want_ghost)
    assert np.array_equal(cars["episode"], oep)
    assert np.array_equal(cars["steps"], osteps)


def _assert_table_equal(lr, sc, oq):
    q = lr.q.cpu().numpy()
    vis = lr.visited.cpu().numpy()
    items = [x for x in oq.items() if x[1] != "B"]
    want_vis = np.zeros_like(vis)
    for s, a, v in items:
        i, k = sc.encode(s), A5.index(a)
        assert q[i, k] == np.float32(v), (s, a)
        want_vis[i] |= 1 << k
    assert np.count_nonzero(q[:, :5]) == sum(1 for _, _, v in items if v != 0.0)
    assert np.array_equal(vis, want_vis)          # updated-key set (what the qvalue file lists)


def _assert_counters_equal(env, ob, keys=("agent_steps", "episodes_done", "episodes_capped", "explore_steps")):
    gs, ws = env.stats(), ob.stats()
    assert gs["internal_errors"] == 0
    for k in keys:
        assert gs[k] == ws[k], k
    assert gs["q_updates"] == gs["agent_steps"]
    assert gs["touched_keys"] == ob.touched_keys()
    assert ob.disagreements() == 0
    return gs


def test_fullsize_pure_exploration_rounds_bit_exact():
    """Config 2 at its real size, the epsilon = 1 segment: 8 rounds in one multi-round launch (+ bitmap
    reduce + cluster pull), one round through the single-round kernels (round kernel + two dense
    pulls: the bench step), then 7 more in one launch -- 16 rounds, 3.3e7 agent-steps."""
    m, BatchedEnv, BatchedQLearner, orc = _mods()
    boxes, idx, choose = m.layout_3[2]["initial_states"][0]
    bad = bad_states_text(2)
    sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
    env = BatchedEnv(sc, N_FULL, seed=31)
    lr = BatchedQLearner(env, episodes=5000, seed=31)
    lr.train_rounds(8)
    lr.train_rounds(1)
    lr.train_rounds(7)
    ob = orc.Batch(orc.Config(m.layout_3[2]), N_FULL, boxes, idx, choose, bad, 10000)
    ob.reset(seed=31)
    oq = orc.QTable(f32=True)
    olp = orc.learner(5000, seed=31)
    for r in range(16):
        ob.q_round(oq, olp, r)
    _assert_table_equal(lr, sc, oq)
    _assert_records_equal(env, ob)
    gs = _assert_counters_equal(env, ob)
    assert gs["agent_steps"] > 30_000_000 and gs["explore_steps"] == gs["agent_steps"]
    assert int((env.records[:, 3] & 0x700).sum()) == 0


@pytest.mark.parametrize("variant", ["default", "substep_kernels", "k1_rounds"])
def test_fullsize_greedy_rounds_bit_exact(variant):
    """Rounds WITH greedy decisions at full size: an episode budget of 2 puts episode 0 at epsilon 0.8 and
    episode 1 at epsilon 0, a step cap of 24 ends episodes quickly, so 40 rounds cross both segments and
    finish with every env frozen.  default = q_rounds_mixed_kernel (cooperative, 148 CTAs, grid barriers between
    env passes and pulls); MDPP_MIXED_ROUNDS=0 = one sub-step kernel + dense pull per car; MDPP_K1_ROUNDS=1 = the
    fused K0 / K1 pair."""
    m, BatchedEnv, BatchedQLearner, orc = _mods()
    boxes, idx, choose = m.layout_3[2]["initial_states"][0]
    bad = bad_states_text(2)
    cap, episodes, rounds = 24, 2, 40
    sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=cap)
    envvars = {"substep_kernels": {"MDPP_MIXED_ROUNDS": "0"}, "k1_rounds": {"MDPP_K1_ROUNDS": "1"}}.get(variant, {})
    os.environ.update(envvars)
    try:
        env = BatchedEnv(sc, N_FULL, seed=37)
        lr = BatchedQLearner(env, episodes=episodes, seed=37)
        for k in (16, 1, 16, 7):
            lr.train_rounds(k)
        torch.cuda.synchronize()
    finally:
        for k in envvars:
            os.environ.pop(k, None)
    ob = orc.Batch(orc.Config(m.layout_3[2]), N_FULL, boxes, idx, choose, bad, cap)
    ob.reset(seed=37)
    oq = orc.QTable(f32=True)
    olp = orc.learner(episodes, seed=37)
    assert list(olp.eps_threshold) == list(lr.params.eps_threshold)
    for r in range(rounds):
        ob.q_round(oq, olp, r)
    _assert_table_equal(lr, sc, oq)
    _assert_records_equal(env, ob)
    gs = _assert_counters_equal(env, ob)
    assert 0 < gs["explore_steps"] < gs["agent_steps"]           # both kinds of decisions happened
    assert int((env.cars()["episode"] >= episodes).sum()) == N_FULL  # every env used up its budget
    assert int((env.records[:, 3] & 0x700).sum()) == 0            # no lane is left parked


def test_fullsize_four_car_learner_bit_exact():
    """Max agents at full size: 4 cars, 1.65 M-state table (2^20 envs), 8 rounds through mdpp_q_train_rounds: scan
    mode, every env exploring -- groups of two rounds per env pass, the pulls of a group on the handle's own stream
    beside the env pass of the next (7 rounds: 2 + 2 + 2 + 1), then a single round in a call of its own."""
    m, BatchedEnv, BatchedQLearner, orc = _mods()
    boxes, idx = ["D2", "D6", "D4", "D8"], [0, 0, 0, 0]
    choose = [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]]
    bad = bad_states_text(2)
    sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=300)
    env = BatchedEnv(sc, N_FULL, seed=41)
    lr = BatchedQLearner(env, episodes=5000, seed=41)
    lr.train_rounds(7)
    lr.train_rounds(1)
    ob = orc.Batch(orc.Config(m.layout_3[2]), N_FULL, boxes, idx, choose, bad, 300)
    ob.reset(seed=41)
    oq = orc.QTable(f32=True)
    olp = orc.learner(5000, seed=41)
    for r in range(8):
        ob.q_round(oq, olp, r)
    _assert_table_equal(lr, sc, oq)
    _assert_records_equal(env, ob)
    _assert_counters_equal(env, ob)


def test_fullsize_four_car_rollout_bit_exact():
    """Config 3 (random-policy sweep, max agents = 4 cars) at 2^20 envs: per-step outputs of 4 rounds in one
    launch, final records of every env and the counters against the oracle."""
    m, BatchedEnv, _, orc = _mods()
    boxes, idx = ["D2", "D6", "D4", "D8"], [0, 0, 0, 0]
    choose = [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]]
    bad = bad_states_text(2)
    rounds = 4
    sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
    env = BatchedEnv(sc, N_FULL, seed=43)
    outs = env.alloc_rollout_outputs(rounds)
    env.rollout_random(rounds, 0, "q_reward", outs)
    ob = orc.Batch(orc.Config(m.layout_3[2]), N_FULL, boxes, idx, choose, bad, 10000)
    ob.reset(seed=43)
    want = ob.rollout_random(rounds, 0, 43, False, threads=os.cpu_count() or 1)
    assert np.array_equal(outs["action"].cpu().numpy(), want["action"])
    assert np.array_equal(outs["reward"].cpu().numpy(), want["reward"])
    assert np.array_equal(outs["done"].cpu().numpy(), want["done"])
    got_next = outs["next_state"].cpu().numpy().reshape(-1)
    moved = want["action"].reshape(-1) != 0xFF
    assert np.array_equal(got_next == -1, ~moved)
    for i in range(0, got_next.size, 1021):        # state ids of a strided sample (string encode on the host)
        if moved[i]:
            assert got_next[i] == sc.encode(orc.state_str(want["next"][i])), i
    _assert_records_equal(env, ob)
    gs, ws = env.stats(), ob.stats()
    for k in ("agent_steps", "episodes_done", "episodes_capped", "reward_sum"):
        assert gs[k] == ws[k], k
    # a second launch continues from the written-back records
    env.rollout_random(2, rounds, "q_reward")
    ob.rollout_random(2, rounds, 43, False, threads=os.cpu_count() or 1, outputs=False)
    _assert_records_equal(env, ob)


def test_reachability_four_cars_bit_exact():
    """-dd precompute, 4 cars: all 28.4 M joint states against the CPU brute force (forward enumeration through
    the scalar oracle env on every host core, backward closure over the reverse adjacency)."""
    m, BatchedEnv, _, orc = _mods()
    boxes, idx = ["D2", "D6", "D4", "D8"], [0, 0, 0, 0]
    choose = [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]]
    sc = m.Scenario(m.layout_3[2], boxes, idx, choose)
    env = BatchedEnv(sc, 32)
    want, live = orc.reachability(orc.Config(m.layout_3[2]), 4, choose, sc.n_local_nodes, sc.node_base)
    for method in ("bfs", "sweep"):
        bitmap, _ = env.deadlock_reachability(method=method)
        got = bitmap.cpu().numpy().view(np.uint32)
        assert np.array_equal(got, want), method
        assert int(np.unpackbits(got.view(np.uint8)).sum()) == live == 13968601


def test_fullsize_three_car_hybrid_bit_exact():
    """Mid-size table (3 cars, 126 540 states) at full size through every change-over of its two roads: the scan
    mode's fused env pass + scanning pulls while the host's bound says "every env explores" (tightened by each
    counter fetch), shared-memory bit maps + dense pulls otherwise, a table copy / settle pass in between.  An
    episode budget of 3 (epsilon 1, then 0.8 .. 0) and a step cap of 30 take every env to frozen in ~45 rounds."""
    m, BatchedEnv, BatchedQLearner, orc = _mods()
    boxes, idx, choose = ["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]]
    bad = bad_states_text(2)
    cap, episodes = 30, 3
    sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=cap)
    env = BatchedEnv(sc, N_FULL, seed=47)
    lr = BatchedQLearner(env, episodes=episodes, seed=47)
    ob = orc.Batch(orc.Config(m.layout_3[2]), N_FULL, boxes, idx, choose, bad, cap)
    ob.reset(seed=47)
    oq = orc.QTable(f32=True)
    olp = orc.learner(episodes, seed=47)
    assert olp.eps_threshold[0] == 1                       # episode 0 explores, the others decide greedily too
    done = 0
    for k in (3, 4, 1, 6):                                 # fused pass(es) first, dense sub-steps once the bound says so;
        lr.train_rounds(k)                                 # every stats fetch lets the fused pass back in while all
        env.stats()                                        # envs are still in episode 0
        done += k
    while int(env.cars()["episode"].min()) < episodes:
        assert done < episodes * (cap + 2)
        lr.train_rounds(16)
        done += 16
    for r in range(done):
        ob.q_round(oq, olp, r)
    _assert_table_equal(lr, sc, oq)
    _assert_records_equal(env, ob)
    gs = _assert_counters_equal(env, ob)
    assert 0 < gs["explore_steps"] < gs["agent_steps"]


def test_fullsize_four_car_greedy_rounds_bit_exact():
    """Scan mode with greedy decisions at full size: 4 cars (1.65 M states), an episode budget of 2 (epsilon 0.8,
    then 0) and a step cap of 24 -- every round takes the sub-step kernels (4 096 CTAs marking the global bitmaps,
    three banks in rotation) and one scanning pull per car, to frozen envs."""
    m, BatchedEnv, BatchedQLearner, orc = _mods()
    boxes, idx = ["D2", "D6", "D4", "D8"], [0, 0, 0, 0]
    choose = [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]]
    bad = bad_states_text(2)
    cap, episodes = 24, 2
    sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=cap)
    env = BatchedEnv(sc, N_FULL, seed=53)
    lr = BatchedQLearner(env, episodes=episodes, seed=53)
    done = 0
    while int(env.cars()["episode"].min()) < episodes:
        assert done < episodes * (cap + 2)
        lr.train_rounds(13)
        done += 13
    ob = orc.Batch(orc.Config(m.layout_3[2]), N_FULL, boxes, idx, choose, bad, cap)
    ob.reset(seed=53)
    oq = orc.QTable(f32=True)
    olp = orc.learner(episodes, seed=53)
    assert olp.eps_threshold[0] == 0                       # no episode explores only
    for r in range(done):
        ob.q_round(oq, olp, r)
    _assert_table_equal(lr, sc, oq)
    _assert_records_equal(env, ob)
    gs = _assert_counters_equal(env, ob)
    assert 0 < gs["explore_steps"] < gs["agent_steps"]
