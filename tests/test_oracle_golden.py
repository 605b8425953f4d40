"""CPU oracle (oracle/mdpp_oracle.c) pinned against vectors recorded from the live reference."""
import glob
import os

import pytest

from conftest import GOLDEN, bad_states_text, load_golden, q_close
from oracle import oracle as orc

TRAJ = sorted(os.path.basename(p)[:-8] for p in glob.glob(os.path.join(GOLDEN, "traj_*.json.gz")))
QRUN = sorted(os.path.basename(p)[:-8] for p in glob.glob(os.path.join(GOLDEN, "qrun_*.json.gz")))
QRAND = sorted(os.path.basename(p)[:-8] for p in glob.glob(os.path.join(GOLDEN, "qrand_*.json.gz")))


def cfg_for(block):
    return orc.Config(load_golden("tables_b%d" % block)["layout"])


@pytest.mark.parametrize("block", [0, 2])
def test_static_graph_matches_reference(block):
    g = load_golden("tables_b%d" % block)
    cfg = cfg_for(block)
    for n, succ in g["action_dictionary"].items():
        nid = orc.node_id(n)
        for a, s in enumerate(succ):
            got = cfg.successor(nid, a)
            assert (got == orc.ORC_NONE) == s.endswith("YY"), (n, a, s)
            if got != orc.ORC_NONE:
                assert orc.node_str(got) == s, (n, a)
        assert orc.mask_str(cfg.base_legal(nid)) == g["base_legal"][n], n
        assert cfg.in_no_node_list(nid) == (n in g["no_node_list"])
    for d, row in g["bfs_len"].items():
        for n, length in row.items():
            assert cfg.bfs_len(orc.node_id(n), orc.node_id(d)) == length, (n, d)


def test_state_string_roundtrip():
    for s in ["D2NxD6NzD3xD6", "U4WxU2W", "D5ExD6NzD1xD6zD8xD8", "D1SxD4EzD2xD6zD3xD8zD5xD4zD7xD7"]:
        assert orc.state_str(orc.state_from_str(s)) == s


def replay_traj(g):
    cfg = cfg_for(g["block"])
    env = orc.Env(cfg, bad_states_text(g["block"]) if g["bad_states"] else None)
    n = len(g["boxes"])
    dqn = g["reward"] == "dqn_reward"
    checked = 0
    for ep in g["episodes"]:
        env.reset()
        env.initialize_cars(n)
        env.initialize(g["idx"], ep["headings"], g["choose"])
        it = iter(ep["steps"])
        done = False
        exhausted = False
        while not done and not exhausted:
            for car in range(n):
                if env.car_done(car):
                    continue
                rec = next(it, None)
                if rec is None:
                    exhausted = True
                    break
                num, state, legal, action, nxt, reward = rec
                assert num == car
                st = env.encodestate(car, g["deadlock_id"])
                assert orc.state_str(st) == state
                assert orc.mask_str(env.legal(st)) == legal, state
                a = orc.ACTIONS.index(action)
                ns = env.successor_state(st, a)
                assert orc.state_str(ns) == nxt
                assert env.reward(st, ns, car, dqn) == reward, (state, action)
                env.update_car(car, ns.pnode)
                checked += 1
            done = env.is_game_ended()
        assert next(it, None) is None
        assert done == ep["done"]
    assert checked == g["n_steps"]


@pytest.mark.parametrize("name", TRAJ)
def test_trajectory_bit_exact(name):
    replay_traj(load_golden(name))


@pytest.mark.parametrize("name", QRUN)
def test_qlearning_run_bit_exact(name):
    """train_given_state replayed with the recorded coin / choice / heading stream: every greedy
    action, every updated Q value (float64, compared exactly) and the final Counter incl. its
    insertion order must match the reference."""
    g = load_golden(name)
    cfg = cfg_for(g["block"])
    env = orc.Env(cfg, bad_states_text(g["block"]) if g["bad_states"] else None)
    q = orc.QTable(f32=False, alpha=g["alpha"], discount=g["discount"])
    n = len(g["boxes"])
    for e, ep in enumerate(g["log"]):
        assert orc.epsilon(g["episodes_total"], e, 3) == ep["eps"]
        env.reset()
        env.initialize_cars(n)
        env.initialize(g["idx"], ep["headings"], g["choose"])
        it = iter(ep["steps"])
        done = False
        stop = False
        while not done and not stop:
            for car in range(n):
                if env.car_done(car):
                    continue
                rec = next(it, None)
                if rec is None:
                    stop = True
                    break
                state, explore, action, nxt, reward, qafter = rec
                st = env.encodestate(car, 1)
                assert orc.state_str(st) == state
                if explore:
                    a = orc.ACTIONS.index(action)
                else:
                    a = q.greedy(env, st)
                    assert orc.ACTIONS[a] == action, (e, state)
                ns = env.successor_state(st, a)
                r = env.reward(st, ns, car, False)
                assert r == reward
                q.update(env, st, a, ns, r)
                assert q.get(st, a) == float(qafter)
                env.update_car(car, ns.pnode)
            done = env.is_game_ended()
    got = q.items()
    want = [(s, a, float(v)) for s, a, v in g["qvals"]]
    assert got == want


@pytest.mark.parametrize("name", QRAND)
def test_qlearning_random_order_run_bit_exact(name):
    """train_given_state_random_order (reference src/algorithms/qlearning.py:241-295) replayed with the
    recorded car draws, coins, choices and headings: a draw that hits a finished car does nothing, every
    other draw is one agent-step of that car; states, greedy actions, rewards, every updated Q value
    (float64, exact) and the final Counter incl. its insertion order must match the reference."""
    g = load_golden(name)
    cfg = cfg_for(g["block"])
    env = orc.Env(cfg, bad_states_text(g["block"]) if g["bad_states"] else None)
    q = orc.QTable(f32=False, alpha=g["alpha"], discount=g["discount"])
    n = len(g["boxes"])
    for e, ep in enumerate(g["log"]):
        assert orc.epsilon(g["episodes_total"], e, 3) == ep["eps"]
        env.reset()
        env.initialize_cars(n)
        env.initialize(g["idx"], ep["headings"], g["choose"])
        it = iter(ep["steps"])
        for car in ep["draws"]:
            assert not env.is_game_ended()
            if env.car_done(car):
                continue
            num, state, explore, action, nxt, reward, qafter = next(it)
            assert num == car
            st = env.encodestate(car, 1)
            assert orc.state_str(st) == state
            if explore:
                a = orc.ACTIONS.index(action)
            else:
                a = q.greedy(env, st)
                assert orc.ACTIONS[a] == action, (e, state)
            ns = env.successor_state(st, a)
            assert orc.state_str(ns) == nxt
            r = env.reward(st, ns, car, False)
            assert r == reward
            q.update(env, st, a, ns, r)
            assert q.get(st, a) == float(qafter)
            env.update_car(car, ns.pnode)
        assert next(it, None) is None
        assert env.is_game_ended()
    got = q.items()
    want = [(s, a, float(v)) for s, a, v in g["qvals"]]
    assert got == want


@pytest.mark.parametrize("name", QRAND)
def test_batch_oracle_random_order_reproduces_reference_run(name):
    """The batch restatement's random-order mode (flags & 16) with ONE env, the reference's recorded car draws and
    decisions fed through the forced byte (high nibble car, low nibble action / 0xE greedy): it visits the
    reference's (car, state, action, reward) sequence and learns its table (float32 vs float64: 1e-5)."""
    np = pytest.importorskip("numpy")
    g = load_golden(name)
    cfg = cfg_for(g["block"])
    n = len(g["boxes"])
    ob = orc.Batch(cfg, 1, g["boxes"], g["idx"], g["choose"], bad_states_text(g["block"]) if g["bad_states"] else None, 0)
    ob.reset(seed=1)
    q = orc.QTable(f32=True)
    lp = orc.learner(g["episodes_total"], seed=1, flags=16)
    A5 = "SLRFT"
    rnd = 0
    for ep in g["log"]:
        heads = [[orc.HEADS.index(h[-1]) for h in ep["headings"]]]
        ob.reset(headings=np.array(heads, dtype=np.uint8), zero_counters=False)
        it = iter(ep["steps"])
        for car in ep["draws"]:
            if ob.env_cars(0)[car][2]:  # a finished car consumes the draw, nothing else (qlearning.py:268-269)
                out = ob.q_substep(q, lp, 0, rnd, forced=np.array([(car << 4) | 0xF], dtype=np.uint8))
                assert out["action"][0] == 0xFF
                rnd += 1
                continue
            num, state, explore, action, nxt, reward, qafter = next(it)
            assert num == car
            low = A5.index(action) if explore else 0xE
            out = ob.q_substep(q, lp, 0, rnd, forced=np.array([(car << 4) | low], dtype=np.uint8))
            rnd += 1
            assert orc.state_str(out["state"][0]) == state
            assert A5[out["action"][0]] == action
            assert out["reward"][0] == reward
            got = q.get(orc.state_from_str(state), orc.ACTIONS.index(action))
            assert q_close(got, float(qafter))
        assert next(it, None) is None
    assert ob.stats()["episodes_done"] == len(g["log"])
    want = {(s, a): float(v) for s, a, v in g["qvals"]}
    got = {(s, a): v for s, a, v in q.items()}
    # the reference's Counter also holds the keys it only read (0.0); the batch table the updated ones
    assert set(got) == {(s[1], s[3]) for ep in g["log"] for s in ep["steps"]}
    for k, v in want.items():
        assert q_close(got.get(k, 0.0), v), k


def test_known_answers():
    g = load_golden("known_answers")
    cfg = cfg_for(2)
    env = orc.Env(cfg, None)
    env.initialize_cars(2)
    env.initialize([0, 0], ["D2N", "D3E"], [[0, 1, 0], [0, 1, 0]])
    st = env.encodestate(0, 1)
    t = g["transition"]
    assert orc.state_str(st) == t["state"] and orc.mask_str(env.legal(st)) == t["legal"]
    ns = env.successor_state(st, orc.ACTIONS.index(t["action"]))
    assert orc.state_str(ns) == t["next"] and env.reward(st, ns, 0) == t["reward"]
    # ghost of a finished car: 5 encodes, then it vanishes
    env = orc.Env(cfg, None)
    env.initialize_cars(2)
    env.initialize([0, 1], ["D5E", "D8N"], [[0, 1, 0], [0, 1, 0]])
    env.update_car(1, orc.node_id("D8E"))
    for want_state, want_legal, want_count in g["ghost"]["encodes"]:
        st = env.encodestate(0, 1)
        assert orc.state_str(st) == want_state
        assert orc.mask_str(env.legal(st)) == want_legal
        assert env.car(1)[3] == want_count
    assert orc.node_str(env.car(1)[0]) == g["ghost"]["parked"]
    assert env.car(1)[1] == g["ghost"]["parked_di"]
    assert orc.state_str(env.encodestate(0, 0)) == g["ghost"]["mode0"]
    for E, sched in g["eps_model3"].items():
        E = int(E)
        if len(sched) == E and not isinstance(sched[0], list):
            assert [orc.epsilon(E, e, 3) for e in range(E)] == sched
        else:
            full = [orc.epsilon(E, e, 3) for e in range(E)]
            assert [[i, full[i]] for i in range(E) if i == 0 or full[i] != full[i - 1]] == sched


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # ... 7 rounds: what the batch semantics draw from (MDPP_PHILOX_ROUNDS in env_core.cuh and mdpp_oracle.c)
    assert orc.philox_rounds() == 7
    assert orc.philox([0, 0, 0, 0], [0, 0], 7) == [0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48]
    assert orc.philox([0xffffffff] * 4, [0xffffffff] * 2, 7) == [0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662]
    assert orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], 7) == \
        [0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a]


def test_dqn_convert_and_mask_match_reference():
    """RL.convert (80 feature bits) and give_mask recorded from the reference's dqn_open_source."""
    g = load_golden("dqn_convert")
    n = 0
    for name, d in g.items():
        env = orc.Env(cfg_for(d["block"]), None)
        for s, f, mask in zip(d["states"], d["features"], d["masks"]):
            st = orc.state_from_str(s)
            assert "".join(map(str, orc.convert(st))) == f, s
            lm = env.legal(st)
            assert [i for i, a in enumerate([0, 1, 2, 3, 5]) if (lm >> a) & 1] == mask, s
            n += 1
    assert n >= 400


def test_float32_oracle_tracks_float64_reference_within_1e5():
    """Same recorded action stream, float32 arithmetic: every table entry within 1e-5 relative of the
    reference's float64 value (the tolerance the north star states for the GPU table)."""
    g = load_golden("qrun_b2_2car_bad")
    env = orc.Env(cfg_for(g["block"]), bad_states_text(g["block"]))
    q = orc.QTable(f32=True, alpha=g["alpha"], discount=g["discount"])
    n = len(g["boxes"])
    for ep in g["log"]:
        env.reset()
        env.initialize_cars(n)
        env.initialize(g["idx"], ep["headings"], g["choose"])
        it = iter(ep["steps"])
        done = False
        while not done:
            for car in range(n):
                if env.car_done(car):
                    continue
                rec = next(it, None)
                if rec is None:
                    done = True
                    break
                st = env.encodestate(car, 1)
                a = orc.ACTIONS.index(rec[2])
                ns = env.successor_state(st, a)
                q.update(env, st, a, ns, env.reward(st, ns, car, False))
                env.update_car(car, ns.pnode)
            done = done or env.is_game_ended()
    want = {(s, a): float(v) for s, a, v in g["qvals"]}
    got = {(s, a): v for s, a, v in q.items()}
    assert set(got) <= set(want)
    for k, v in got.items():
        assert q_close(v, want[k]), k


BATCH_CLAIM = [
    (2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], True, 1024, 150, 0),
    (2, ["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]], True, 512, 120, 0),
    (0, ["U4", "U8", "U5", "U7"], [0, 1, 0, 0], [[0, 0, 0]] * 4, False, 512, 120, 0),
    (2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], True, 512, 150, 2 | 4),  # encodestate(.,0) + stall detection
    # random car order: different cars of different envs write the same table in one sub-step
    (2, ["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]], True, 512, 150, 16),
    (2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], False, 1024, 150, 16),
]


@pytest.mark.parametrize("case", BATCH_CLAIM, ids=lambda c: "b%d_%dcar_f%d" % (c[0], len(c[1]), c[7]))
def test_batch_contributors_of_a_key_agree(case):
    """The batch semantics define the value a touched key receives as "the value its contributing
    envs computed": (state, action) and the table snapshot determine it (the reward's clash term is
    0 under legal actions).  The oracle computes every contributor's value independently from its
    own env and counts the keys whose contributors differ: there must be none."""
    import marl_dpp_b200 as m
    block, boxes, idx, choose, bad, n, rounds, flags = case
    badtxt = bad_states_text(block) if bad else None
    b = orc.Batch(orc.Config(m.layout_3[block]), n, boxes, idx, choose, badtxt, 300)
    b.reset(seed=21)
    q = orc.QTable(f32=True)
    lp = orc.learner(6, seed=21, flags=flags)
    for r in range(rounds):
        b.q_round(q, lp, r)
    assert b.stats()["agent_steps"] > 10000
    assert b.touched_keys() > 0
    assert b.disagreements() == 0


def test_batch_oracle_is_thread_count_invariant():
    """The full-size GPU parity tests run the oracle's per-env phase on every host core: the table (values AND
    insertion order), the envs, the counters and the reachability bitmap must not depend on the thread count."""
    import marl_dpp_b200 as m
    from oracle import oracle as orc
    boxes, idx, choose = m.layout_3[2]["initial_states"][0]
    bad = bad_states_text(2)
    runs = []
    try:
        for threads in (1, 4):
            orc.set_threads(threads)
            ob = orc.Batch(orc.Config(m.layout_3[2]), 6000, boxes, idx, choose, bad, 30)
            ob.reset(seed=13)
            oq = orc.QTable(f32=True)
            olp = orc.learner(3, seed=13)          # greedy decisions from episode 1 on
            for r in range(60):
                ob.q_round(oq, olp, r)
            cars, ep, steps = ob.export()
            runs.append((oq.items(), cars.copy(), ep.copy(), steps.copy(), ob.stats(), ob.touched_keys(),
                         ob.disagreements()))
        a, b = runs
        assert a[0] == b[0] and a[4] == b[4] and a[5] == b[5] and a[6] == b[6] == 0
        for i in (1, 2, 3):
            assert (a[i] == b[i]).all()
        assert 0 < a[4]["explore_steps"] < a[4]["agent_steps"] and a[4]["episodes_done"] > 6000
        # export agrees with the per-env accessors
        for e in (0, 17, 5999):
            assert [tuple(int(x) for x in c[:3]) for c in a[1][e]] == [c[:3] for c in ob.env_cars(e)]
            assert (int(a[2][e]), int(a[3][e])) == ob.counters(e)
        choose3 = [[0, 1, 0], [1, 0, 0], [2, 1, 0]]
        orc.set_threads(1)
        w1, l1 = orc.reachability(orc.Config(m.layout_3[2]), 3, choose3)
        orc.set_threads(3)
        w3, l3 = orc.reachability(orc.Config(m.layout_3[2]), 3, choose3)
        assert l1 == l3 and (w1 == w3).all()
    finally:
        orc.set_threads(1)
