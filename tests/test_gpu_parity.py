"""GPU parity: the CUDA path (through the C ABI) against (a) vectors recorded from the live
reference and (b) the CPU oracle on the same seeded inputs.  Integer / index work is bit-exact;
Q values are bit-exact against the float32 oracle and within 1e-5 relative of the reference's
float64 values."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, bad_states_text, load_golden, q_close

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TRAJ = sorted(os.path.basename(p)[:-8] for p in glob.glob(os.path.join(GOLDEN, "traj_*.json.gz")))
QRUN = sorted(os.path.basename(p)[:-8] for p in glob.glob(os.path.join(GOLDEN, "qrun_*.json.gz")))
A5 = "SLRFT"
HEADS = "NESW"


def _mods():
    import marl_dpp_b200 as m
    from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner
    from oracle import oracle as orc
    return m, BatchedEnv, BatchedQLearner, orc


def _scenario(m, g, step_cap=0):
    bad = bad_states_text(g["block"]) if g["bad_states"] else None
    return m.Scenario(m.layout_3[g["block"]], g["boxes"], g["idx"], g["choose"], bad_states_text=bad,
                      step_cap=step_cap)


def _mask_str(x):
    return "".join(A5[a] for a in range(5) if (x >> a) & 1)


def _rounds_of(steps, n_cars):
    """Split one episode's step list into round-robin rounds: [{car: record}]."""
    rounds, cur, last = [], {}, -1
    for rec in steps:
        car = rec[0]
        if car <= last:
            rounds.append(cur)
            cur = {}
        cur[car] = rec
        last = car
    if cur:
        rounds.append(cur)
    return rounds


@pytest.mark.parametrize("name", TRAJ)
def test_reference_trajectories_bit_exact(name):
    """Every recorded reference episode becomes one env of a batch; the recorded headings and
    actions are replayed and state string, legal set, next state and reward must match at every
    agent-step (reference src/env_gym.py encodestate/getlegalactions/step)."""
    m, BatchedEnv, _, _ = _mods()
    g = load_golden(name)
    sc = _scenario(m, g)
    eps = g["episodes"]
    n, nc = len(eps), len(g["boxes"])
    env = BatchedEnv(sc, n)
    heads = torch.tensor([[HEADS.index(h[-1]) for h in ep["headings"]] for ep in eps], dtype=torch.uint8)
    env.reset(headings=heads)
    per_env = [_rounds_of(ep["steps"], nc) for ep in eps]
    checked = 0
    for r in range(max(len(x) for x in per_env)):
        for c in range(nc):
            recs = [per_env[e][r].get(c) if r < len(per_env[e]) else None for e in range(n)]
            if not any(recs):
                continue
            acts = torch.tensor([A5.index(x[3]) if x else 0xFF for x in recs], dtype=torch.uint8)
            out = env.step_car(c, acts, deadlock_id=g["deadlock_id"], reward=g["reward"])
            out = {k: v.cpu().numpy() for k, v in out.items()}
            for e, rec in enumerate(recs):
                if rec is None:
                    assert out["status"][e] == 1
                    continue
                _, state, legal, action, nxt, reward = rec
                assert out["status"][e] == 0, (name, e, r, c)
                assert sc.decode(out["state"][e]) == state
                assert _mask_str(out["legal"][e]) == legal, state
                assert sc.decode(out["next_state"][e]) == nxt
                assert out["reward"][e] == np.float32(reward), (state, action)
                checked += 1
    assert checked == g["n_steps"]
    done = env.cars()["done"].cpu().numpy().all(axis=1)
    assert [bool(x) for x in done] == [ep["done"] for ep in eps]
    assert env.stats()["agent_steps"] == g["n_steps"]


ROLLOUTS = [
    ("b2_2car_bad", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], True, 0, 2048, 96),
    ("b2_4car_bad_cap", 2, ["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]],
     True, 300, 1024, 128),
    ("b0_1car", 0, ["U4"], [1], [[0, 0, 0]], False, 0, 1024, 64),
    ("b0_3car", 0, ["U4", "U8", "U7"], [0, 1, 1], [[0, 0, 0]] * 3, True, 500, 1024, 96),
    ("b2_5car_cap", 2, ["D1", "D2", "D3", "D5", "D7"], [0, 1, 0, 1, 0],
     [[0, 1, 0], [2, 1, 0], [1, 0, 0], [2, 2, 0], [3, 4, 0]], True, 200, 512, 96),
]


def _ghost_of(done, count):
    if not done:
        return 5
    return count if count >= 0 else 7


@pytest.mark.parametrize("case", ROLLOUTS, ids=[c[0] for c in ROLLOUTS])
@pytest.mark.parametrize("dqn", [False, True], ids=["q_reward", "dqn_reward"])
def test_random_rollout_matches_oracle(case, dqn):
    """Philox-driven random-policy rollout with auto-reset: actions, rewards, done flags, state
    ids, final packed records and counters must equal the oracle's."""
    m, BatchedEnv, _, orc = _mods()
    _, block, boxes, idx, choose, bad, cap, n, rounds = case
    badtxt = bad_states_text(block) if bad else None
    sc = m.Scenario(m.layout_3[block], boxes, idx, choose, bad_states_text=badtxt, step_cap=cap)
    env = BatchedEnv(sc, n, env_id_base=1000, seed=77)
    outs = env.alloc_rollout_outputs(rounds)
    reward = "dqn_reward" if dqn else "q_reward"
    half = rounds // 2  # two launches: the round counter and the records carry over
    env.rollout_random(half, 0, reward, {k: v[:half] for k, v in outs.items()})
    env.rollout_random(rounds - half, half, reward, {k: v[half:] for k, v in outs.items()})
    got = {k: v.cpu().numpy() for k, v in outs.items()}

    ob = orc.Batch(orc.Config(m.layout_3[block]), n, boxes, idx, choose, badtxt, cap, env_id_base=1000)
    ob.reset(seed=77)
    want = ob.rollout_random(rounds, 0, 77, dqn, threads=4)
    assert np.array_equal(got["action"], want["action"])
    assert np.array_equal(got["reward"], want["reward"])
    assert np.array_equal(got["done"], want["done"])
    flat_next = got["next_state"].reshape(-1)
    flat_act = got["action"].reshape(-1)
    rng = np.random.default_rng(0)
    for i in rng.choice(flat_next.size, size=4000, replace=False):
        if flat_act[i] == 0xFF:
            assert flat_next[i] == -1
        else:
            assert sc.decode(int(flat_next[i])) == orc.state_str(want["next"][int(i)])
    cars = {k: v.cpu().numpy() for k, v in env.cars().items()}
    for e in range(0, n, 7):
        oc = ob.env_cars(e)
        assert [int(x) for x in cars["node"][e]] == [c[0] for c in oc]
        assert [int(x) for x in cars["dest_index"][e]] == [c[1] for c in oc]
        assert [int(x) for x in cars["done"][e]] == [c[2] for c in oc]
        assert [int(x) for x in cars["ghost"][e]] == [_ghost_of(c[2], c[3]) for c in oc]
        assert (int(cars["episode"][e]), int(cars["steps"][e])) == ob.counters(e)
    gs, ws = env.stats(), ob.stats()
    assert gs["internal_errors"] == 0
    for k in ("agent_steps", "episodes_done", "episodes_capped", "reward_sum"):
        assert gs[k] == ws[k], k
    assert gs["agent_steps"] == int((got["action"] != 0xFF).sum())


@pytest.mark.parametrize("name", QRUN)
def test_reference_qlearning_runs(name):
    """RL.train_given_state replayed on ONE env with the recorded coin / choice / heading stream.
    Greedy steps are chosen by the CUDA argmax and must equal the reference's action; every
    state string and reward must match; the final table must be bit-identical to the float32
    oracle and within 1e-5 relative of the reference's float64 Counter."""
    m, BatchedEnv, BatchedQLearner, orc = _mods()
    g = load_golden(name)
    sc = _scenario(m, g)
    nc = len(g["boxes"])
    env = BatchedEnv(sc, 1)
    lr = BatchedQLearner(env, episodes=g["episodes_total"], alpha=g["alpha"], discount=g["discount"])
    ocfg = orc.Config(m.layout_3[g["block"]])
    oenv = orc.Env(ocfg, bad_states_text(g["block"]) if g["bad_states"] else None)
    oq = orc.QTable(f32=True, alpha=g["alpha"], discount=g["discount"])
    for ep in g["log"]:
        heads = torch.tensor([[HEADS.index(h[-1]) for h in ep["headings"]]], dtype=torch.uint8)
        env.reset(headings=heads, zero_counters=False)
        oenv.reset()
        oenv.initialize_cars(nc)
        oenv.initialize(g["idx"], ep["headings"], g["choose"])
        it = iter(ep["steps"])
        done = False
        while not done:
            for c in range(nc):
                if oenv.car_done(c):
                    continue
                rec = next(it, None)
                if rec is None:
                    done = True
                    break
                state, explore, action, nxt, reward, _ = rec
                forced = torch.tensor([A5.index(action) if explore else 0xFE], dtype=torch.uint8)
                out = lr.substep(c, forced=forced, outputs=True)
                lr.finalize()
                assert A5[int(out["action"][0])] == action, (name, state)
                assert sc.decode(int(out["state"][0])) == state
                assert float(out["reward"][0]) == float(reward)
                st = oenv.encodestate(c, 1)
                a6 = orc.ACTIONS.index(action)
                ns = oenv.successor_state(st, a6)
                oq.update(oenv, st, a6, ns, oenv.reward(st, ns, c, False))
                oenv.update_car(c, ns.pnode)
            done = done or oenv.is_game_ended()
    q = lr.q.cpu().numpy()
    ref64 = {(s, a): float(v) for s, a, v in g["qvals"]}
    n_checked = 0
    for s, a, v32 in oq.items():
        if a == "B":
            continue
        got = q[sc.encode(s), A5.index(a)]
        assert got == np.float32(v32), (s, a)
        want = ref64[(s, a)]
        assert q_close(float(got), want), (s, a, got, want)
        n_checked += 1
    assert n_checked == len(ref64)
    # keys the CUDA path marks as updated == keys the reference updated (nonzero or not)
    updated = {(s[0], s[2]) for ep in g["log"] for s in ep["steps"]}
    assert {(s, a) for s, a, _ in lr.qvalue_items()} == updated


def torch_u8(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint8))


QBATCH = [
    ("b2_2car_bad", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], True, 400, 512, 40, 300),
    ("b0_1car", 0, ["U4"], [1], [[0, 0, 0]], False, 0, 256, 30, 200),
    ("b2_3car_bad", 2, ["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]], True, 300, 384, 30, 150),
    # 1.65 M states: scan mode (global touched bitmaps + mirror table + scanning pull) instead of the dense ping-pong
    ("b2_4car_bad_list", 2, ["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]],
     True, 300, 320, 20, 80),
    # the same small table forced through the scan path (global touched bitmaps + mirror table + scanning pull)
    ("b2_2car_bad_forced_list", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], True, 400, 512, 40, 120),
]


@pytest.mark.parametrize("case", QBATCH, ids=[c[0] for c in QBATCH])
def test_batched_qlearning_matches_oracle(case):
    """Many envs sharing one table, Philox decisions, per-env epsilon schedule, auto-reset and
    episode budget: per-sub-step actions and the final table are bit-identical to the oracle's
    restatement of the batch semantics."""
    m, BatchedEnv, BatchedQLearner, orc = _mods()
    name, block, boxes, idx, choose, bad, cap, n, episodes, rounds = case
    badtxt = bad_states_text(block) if bad else None
    sc = m.Scenario(m.layout_3[block], boxes, idx, choose, bad_states_text=badtxt, step_cap=cap)
    if name.endswith("forced_list"):
        os.environ["MDPP_DENSE"] = "0"
    try:
        env = BatchedEnv(sc, n, seed=5)
    finally:
        os.environ.pop("MDPP_DENSE", None)
    lr = BatchedQLearner(env, episodes=episodes, seed=5)
    ob = orc.Batch(orc.Config(m.layout_3[block]), n, boxes, idx, choose, badtxt, cap)
    ob.reset(seed=5)
    oq = orc.QTable(f32=True)
    olp = orc.learner(episodes, seed=5)
    assert list(olp.eps_threshold) == list(lr.params.eps_threshold)
    assert list(olp.eps_q16) == list(lr.params.eps_q16)
    nc = len(boxes)
    for r in range(rounds):
        for c in range(nc):
            out = lr.substep(c, outputs=True)
            lr.finalize()
            want = ob.q_substep(oq, olp, c, r)
            assert np.array_equal(out["action"].cpu().numpy(), want["action"]), (r, c)
            assert np.array_equal(out["reward"].cpu().numpy(), want["reward"]), (r, c)
        lr.round += 1
    q = lr.q.cpu().numpy()
    items = [x for x in oq.items() if x[1] != "B"]
    for s, a, v in items:
        assert q[sc.encode(s), A5.index(a)] == np.float32(v), (s, a)
    assert np.count_nonzero(q[:, :5]) == sum(1 for _, _, v in items if v != 0.0)
    gs, ws = env.stats(), ob.stats()
    assert gs["internal_errors"] == 0
    for k in ("agent_steps", "episodes_done", "episodes_capped", "reward_sum", "explore_steps"):
        assert gs[k] == ws[k], k
    assert gs["q_updates"] == gs["agent_steps"]
    # the batch definition rests on every contributor of a key computing the same value
    assert ob.disagreements() == 0
    assert gs["touched_keys"] == ob.touched_keys()
    # the fused multi-round entry point continues identically
    lr.train_rounds(5)
    for r in range(rounds, rounds + 5):
        ob.q_round(oq, olp, r)
    q = lr.q.cpu().numpy()
    for s, a, v in [x for x in oq.items() if x[1] != "B"]:
        assert q[sc.encode(s), A5.index(a)] == np.float32(v), (s, a)


RANDOM_ORDER = [
    ("b2_2car_bad", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], True, 400, 512, 40, 300),
    ("b0_1car", 0, ["U4"], [1], [[0, 0, 0]], False, 0, 256, 30, 120),
    ("b2_3car_bad", 2, ["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]], True, 300, 384, 30, 150),
    ("b2_4car_bad_list", 2, ["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]],
     True, 300, 320, 20, 60),
]


@pytest.mark.parametrize("case", RANDOM_ORDER, ids=[c[0] for c in RANDOM_ORDER])
def test_batched_random_order_matches_oracle(case):
    """MDPP_LEARN_RANDOM_ORDER (train_given_state_random_order, qlearning.py:241-295): every env draws the
    car that acts, a finished car means no agent-step, the episode may end after any agent-step.  Per
    sub-step the encoded state, action and reward of every env, then the table and the counters, are
    bit-identical to the oracle's restatement; every third round the car and/or the action are forced
    through the high / low nibble of the forced byte."""
    m, BatchedEnv, BatchedQLearner, orc = _mods()
    from marl_dpp_b200.tables import LEARN_RANDOM_ORDER
    name, block, boxes, idx, choose, bad, cap, n, episodes, rounds = case
    badtxt = bad_states_text(block) if bad else None
    sc = m.Scenario(m.layout_3[block], boxes, idx, choose, bad_states_text=badtxt, step_cap=cap)
    env = BatchedEnv(sc, n, seed=9)
    lr = BatchedQLearner(env, episodes=episodes, seed=9, flags=LEARN_RANDOM_ORDER)
    ob = orc.Batch(orc.Config(m.layout_3[block]), n, boxes, idx, choose, badtxt, cap)
    ob.reset(seed=9)
    oq = orc.QTable(f32=True)
    olp = orc.learner(episodes, seed=9, flags=LEARN_RANDOM_ORDER)
    nc = len(boxes)
    rng = np.random.default_rng(1)
    cars_seen = set()
    for r in range(rounds):
        for c in range(nc):
            forced = None
            if r % 3 == 2:  # high nibble: car or 0xF = own draw; low nibble: action, 0xE greedy, 0xF own decision
                hi = rng.choice(np.array(list(range(nc)) + [0xF, 0xF], dtype=np.uint8), size=n)
                lo = rng.choice(np.array([0, 1, 2, 3, 4, 0xE, 0xF, 0xF], dtype=np.uint8), size=n)
                forced = ((hi.astype(np.uint8) << 4) | lo).astype(np.uint8)
            out = lr.substep(c, forced=None if forced is None else torch_u8(forced), outputs=True)
            lr.finalize()
            want = ob.q_substep(oq, olp, c, r, forced=forced)
            assert np.array_equal(out["action"].cpu().numpy(), want["action"]), (r, c)
            assert np.array_equal(out["reward"].cpu().numpy(), want["reward"]), (r, c)
            got_state = out["state"].cpu().numpy()
            for e in range(0, n, 37):
                if want["action"][e] != 0xFF:
                    st = orc.state_str(want["state"][e])
                    assert got_state[e] == sc.encode(st), (r, c, e)
                    cars_seen.add(st.split("x")[1].split("z")[0])
        lr.round += 1
    if nc > 1:
        assert len(cars_seen) > 1  # different cars (different destinations) took agent-steps
    q = lr.q.cpu().numpy()
    items = [x for x in oq.items() if x[1] != "B"]
    for s, a, v in items:
        assert q[sc.encode(s), A5.index(a)] == np.float32(v), (s, a)
    assert np.count_nonzero(q[:, :5]) == sum(1 for _, _, v in items if v != 0.0)
    gs, ws = env.stats(), ob.stats()
    assert gs["internal_errors"] == 0
    for k in ("agent_steps", "episodes_done", "episodes_capped", "reward_sum", "explore_steps"):
        assert gs[k] == ws[k], k
    assert gs["episodes_done"] > 0
    assert ob.disagreements() == 0
    assert gs["touched_keys"] == ob.touched_keys()
    # the multi-round entry point continues identically (n_cars sub-steps per round)
    lr.train_rounds(5)
    for r in range(rounds, rounds + 5):
        ob.q_round(oq, olp, r)
    q = lr.q.cpu().numpy()
    for s, a, v in [x for x in oq.items() if x[1] != "B"]:
        assert q[sc.encode(s), A5.index(a)] == np.float32(v), (s, a)
    assert env.stats()["agent_steps"] == ob.stats()["agent_steps"]
    # back to round-robin rounds on the same handle: the host's bound on the episode counters must have followed
    # the random-order sub-steps (any of them may end an episode), or greedy lanes would be taken for explorers
    lr.params.flags = 0
    olp.flags = 0
    lr.train_rounds(6)
    for r in range(rounds + 5, rounds + 11):
        ob.q_round(oq, olp, r)
    q = lr.q.cpu().numpy()
    for s, a, v in [x for x in oq.items() if x[1] != "B"]:
        assert q[sc.encode(s), A5.index(a)] == np.float32(v), (s, a)
    gs, ws = env.stats(), ob.stats()
    for k in ("agent_steps", "episodes_done", "explore_steps"):
        assert gs[k] == ws[k], k
    assert int((env.records[:, 3] & 0x700).sum()) == 0


FUSED = [
    ("b2_2car_bad", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], True, 300, 1024, 10, 1400, 7, 0),
    ("b2_2car", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], False, 0, 777, 8, 900, 50, 0),
    ("b0_1car", 0, ["U4"], [1], [[0, 0, 0]], False, 200, 1000, 12, 600, 3, 0),
    # 50 episodes: the epsilon = 1 segment ends at episode 20.  The host's bound on the episode counters reaches 20
    # after 20 rounds although no env is near it (rounds then run with the catch-up kernel), every stats fetch
    # tightens it again (rounds without), until envs really leave the segment.
    ("b2_2car_bad_stats", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], True, 40, 512, 50, 900, 11, 4),
]


@pytest.mark.parametrize("k1", [False, True, "substep"], ids=["default", "k1_rounds", "substep_kernels"])
@pytest.mark.parametrize("case", FUSED, ids=[c[0] for c in FUSED])
def test_train_rounds_fused_path_matches_oracle(case, k1):
    """mdpp_q_train_rounds on 1- and 2-car tables takes the fused round kernels (q_round.cuh: table-driven
    steps, car 1 stepped in the same pass when it explores; several rounds per launch while every env
    explores) and, once greedy decisions appear, the cooperative multi-round kernel (env pass / grid barrier / pull
    per sub-step) -- or, with MDPP_K1_ROUNDS=1, the fused pair K0 / K1 (car 1 parked and finished by the catch-up
    kernel when it acts greedily), or, with MDPP_MIXED_ROUNDS=0, one sub-step kernel + pull per car.  The episode budget is small so that the run crosses every epsilon segment
    (1, .8, .5, .3, .15, 0) and ends with frozen envs; table, records and counters must equal the
    oracle's sub-step-by-sub-step restatement bit for bit."""
    m, BatchedEnv, BatchedQLearner, orc = _mods()
    name, block, boxes, idx, choose, bad, cap, n, episodes, rounds, chunk, stats_every = case
    badtxt = bad_states_text(block) if bad else None
    sc = m.Scenario(m.layout_3[block], boxes, idx, choose, bad_states_text=badtxt, step_cap=cap)
    # rounds with greedy decisions: default = the cooperative multi-round kernel; the fused K0 / K1 pair; the
    # sub-step kernels
    envvars = {True: {"MDPP_K1_ROUNDS": "1"}, "substep": {"MDPP_MIXED_ROUNDS": "0"}}.get(k1, {})
    os.environ.update(envvars)
    try:
        _fused_path_body(m, BatchedEnv, BatchedQLearner, orc, sc, case, badtxt)
    finally:
        for key in envvars:
            os.environ.pop(key, None)


def _fused_path_body(m, BatchedEnv, BatchedQLearner, orc, sc, case, badtxt):
    name, block, boxes, idx, choose, bad, cap, n, episodes, rounds, chunk, stats_every = case
    env = BatchedEnv(sc, n, seed=17)
    lr = BatchedQLearner(env, episodes=episodes, seed=17)
    ob = orc.Batch(orc.Config(m.layout_3[block]), n, boxes, idx, choose, badtxt, cap)
    ob.reset(seed=17)
    oq = orc.QTable(f32=True)
    olp = orc.learner(episodes, seed=17)
    done = 0
    while done < rounds:
        k = min(chunk, rounds - done)
        lr.train_rounds(k)
        for r in range(done, done + k):
            ob.q_round(oq, olp, r)
        done += k
        if stats_every and (done // chunk) % stats_every == 0:
            env.stats()  # synchronises and tightens the host-side episode bound
        if done in (chunk, rounds) or done % (chunk * 20) == 0:  # checkpoints along the way
            q = lr.q.cpu().numpy()
            items = [x for x in oq.items() if x[1] != "B"]
            for s, a, v in items:
                assert q[sc.encode(s), A5.index(a)] == np.float32(v), (done, s, a)
            assert np.count_nonzero(q[:, :5]) == sum(1 for _, _, v in items if v != 0.0)
    cars = {k: v.cpu().numpy() for k, v in env.cars().items()}
    for e in range(0, n, 3):
        oc = ob.env_cars(e)
        assert [int(x) for x in cars["node"][e]] == [c[0] for c in oc], e
        assert [int(x) for x in cars["dest_index"][e]] == [c[1] for c in oc]
        assert [int(x) for x in cars["done"][e]] == [c[2] for c in oc]
        assert [int(x) for x in cars["ghost"][e]] == [_ghost_of(c[2], c[3]) for c in oc]
        assert (int(cars["episode"][e]), int(cars["steps"][e])) == ob.counters(e)
    gs, ws = env.stats(), ob.stats()
    assert gs["internal_errors"] == 0
    for k in ("agent_steps", "episodes_done", "episodes_capped", "explore_steps"):
        assert gs[k] == ws[k], k
    assert gs["q_updates"] == gs["agent_steps"]
    assert gs["touched_keys"] == ob.touched_keys()
    assert 0 < gs["explore_steps"] < gs["agent_steps"]          # both kinds of decisions happened
    if not stats_every:
        assert int((cars["episode"] >= episodes).sum()) > 0       # some envs used up their budget
    assert ob.disagreements() == 0
    assert int((env.records[:, 3] & 0x700).sum()) == 0            # no lane is left parked


def test_pure_exploration_rounds_share_a_launch():
    """While every env is certainly in the epsilon = 1 segment, mdpp_q_train_rounds runs up to 8 rounds per
    env-kernel launch (q_rounds_pure_kernel) and then all their pulls in two launches (bitmap reduce + one
    cluster that holds the table in distributed shared memory): far fewer launches, the same table, records,
    visited bits and counters as round-by-round launches -- and as the oracle.  Variant: the dense pulls after
    the multi-round kernel (MDPP_MULTI_PULL=0)."""
    import torch
    m, BatchedEnv, BatchedQLearner, orc = _mods()
    variants = [("default", {}), ("round by round", {"MDPP_MULTI_ROUND": "0"}), ("dense pulls", {"MDPP_MULTI_PULL": "0"})]
    for boxes, idx, choose, block, per_round in ((["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], 2, 3), (["U4"], [1], [[0, 0, 0]], 0, 2)):
        nc = len(boxes)
        badtxt = bad_states_text(block)
        sc = m.Scenario(m.layout_3[block], boxes, idx, choose, bad_states_text=badtxt, step_cap=300)
        runs = {}
        for name, envvars in variants:
            os.environ.update(envvars)
            try:
                env = BatchedEnv(sc, 4000, seed=23)
            finally:
                for k in envvars:
                    os.environ.pop(k, None)
            lr = BatchedQLearner(env, episodes=5000, seed=23)
            l0 = env.launch_count()
            lr.train_rounds(19)   # 8 + 8 + 3
            lr.train_rounds(1)    # a single round takes the one-round kernels
            lr.train_rounds(2)
            torch.cuda.synchronize()
            runs[name] = (env.launch_count() - l0, lr.q.clone(), env.records.clone(), env.stats(), lr.visited.clone())
        assert runs["round by round"][0] >= 22 * per_round
        assert runs["default"][0] <= 16                       # 4 multi-round groups x 3 launches + one single round
        assert runs["default"][0] < runs["dense pulls"][0] < runs["round by round"][0]
        ref = runs["round by round"]
        for name in ("default", "dense pulls"):
            got = runs[name]
            assert torch.equal(got[1], ref[1]), name
            assert torch.equal(got[2], ref[2]), name
            assert torch.equal(got[4], ref[4]), name
            for k in ("agent_steps", "episodes_done", "explore_steps", "touched_keys", "q_updates", "internal_errors"):
                assert got[3][k] == ref[3][k], (name, k)
        assert ref[3]["internal_errors"] == 0
        ob = orc.Batch(orc.Config(m.layout_3[block]), 4000, boxes, idx, choose, badtxt, 300)
        ob.reset(seed=23)
        oq = orc.QTable(f32=True)
        olp = orc.learner(5000, seed=23)
        for r in range(22):
            ob.q_round(oq, olp, r)
        q = runs["default"][1].cpu().numpy()
        for s_, a, v in [x for x in oq.items() if x[1] != "B"]:
            assert q[sc.encode(s_), A5.index(a)] == np.float32(v), (s_, a)
        assert runs["default"][3]["agent_steps"] == ob.stats()["agent_steps"]
        assert runs["default"][3]["touched_keys"] == ob.touched_keys()


def test_training_after_other_episode_advancing_calls():
    """A random rollout advances episode counters outside the learner kernels; training afterwards must still
    equal the oracle (the host-side episode bound turns conservative), and the env-kernel profiling hook steps
    every live env once."""
    m, BatchedEnv, BatchedQLearner, orc = _mods()
    boxes, idx, choose = ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]]
    badtxt = bad_states_text(2)
    sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=badtxt, step_cap=60)
    n, episodes = 600, 9
    env = BatchedEnv(sc, n, seed=23)
    ob = orc.Batch(orc.Config(m.layout_3[2]), n, boxes, idx, choose, badtxt, 60)
    ob.reset(seed=23)
    env.rollout_random(90, 0)                 # a few episodes per env, no learner involved
    ob.rollout_random(90, 0, 23, False, threads=2, outputs=False)
    lr = BatchedQLearner(env, episodes=episodes, seed=23)
    oq = orc.QTable(f32=True)
    olp = orc.learner(episodes, seed=23)
    lr.round = 90
    lr.train_rounds(260)
    for r in range(90, 350):
        ob.q_round(oq, olp, r)
    q = lr.q.cpu().numpy()
    for s, a, v in [x for x in oq.items() if x[1] != "B"]:
        assert q[sc.encode(s), A5.index(a)] == np.float32(v), (s, a)
    cars = {k: v.cpu().numpy() for k, v in env.cars().items()}
    for e in range(0, n, 5):
        oc = ob.env_cars(e)
        assert [int(x) for x in cars["node"][e]] == [c[0] for c in oc], e
        assert (int(cars["episode"][e]), int(cars["steps"][e])) == ob.counters(e)
    assert int((env.records[:, 3] & 0x700).sum()) == 0
    # profiling hook: one launch of the round kernel alone
    env2 = BatchedEnv(sc, 4096, seed=5)
    lr2 = BatchedQLearner(env2, episodes=5000, seed=5)
    lr2.profile_env_kernel()
    st = env2.stats()
    assert st["agent_steps"] == 2 * 4096 and st["q_updates"] == st["agent_steps"]


def test_full_size_properties():
    """BASELINE size (2^20 envs, layout 3 block 2, 2 cars): determinism, shard invariance
    (env ids key the RNG, so two half-size shards equal one full batch) and counter identities."""
    m, BatchedEnv, BatchedQLearner, _ = _mods()
    n = 1 << 20
    sc = m.Scenario(m.layout_3[2], ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]],
                    bad_states_text=bad_states_text(2), step_cap=10000)
    full = BatchedEnv(sc, n, seed=9)
    full.rollout_random(64)
    a = BatchedEnv(sc, n // 2, env_id_base=0, seed=9)
    b = BatchedEnv(sc, n // 2, env_id_base=n // 2, seed=9)
    a.rollout_random(64)
    b.rollout_random(64)
    assert torch.equal(full.records[:n // 2], a.records)
    assert torch.equal(full.records[n // 2:], b.records)
    sf, sa, sb = full.stats(), a.stats(), b.stats()
    for k in ("agent_steps", "episodes_done", "reward_sum"):
        assert sf[k] == sa[k] + sb[k]
    cars = full.cars()
    assert int(cars["steps"].sum()) + 0 <= sf["agent_steps"]
    # learner: two identical runs give bit-identical tables (integer accumulators: order-free)
    tables = []
    for _ in range(2):
        env = BatchedEnv(sc, n, seed=9)
        lr = BatchedQLearner(env, episodes=1000, seed=9)
        lr.train_rounds(24)
        st = env.stats()
        assert st["q_updates"] == st["agent_steps"] > 0 and st["internal_errors"] == 0
        tables.append(lr.q.clone())
    assert torch.equal(tables[0], tables[1])
    assert torch.isfinite(tables[0]).all()


def test_c_abi_error_paths():
    m, BatchedEnv, _, _ = _mods()
    from marl_dpp_b200 import _lib
    sc = m.Scenario(m.layout_3[0], ["U4"], [1], [[0, 0, 0]])
    env = BatchedEnv(sc, 64)
    acts = torch.zeros(64, dtype=torch.uint8, device="cuda")
    with pytest.raises(_lib.MdppError):
        env.step_car(3, acts)
    out = env.step_car(0, torch.full((64,), 3, dtype=torch.uint8))  # U4 facing a wall for some headings
    st = out["status"].cpu().numpy()
    assert set(st.tolist()) <= {0, 2}
    out = env.step_car(0, torch.full((64,), 0xFF, dtype=torch.uint8))
    assert (out["status"].cpu().numpy() == 1).all()
