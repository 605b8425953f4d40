"""GPU: the reference-facing surfaces -- string-API Environment, RL plugin, learner flags."""
import os

import numpy as np
import pytest

from conftest import bad_states_text, load_golden, q_close

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
A5 = "SLRFT"


@pytest.fixture()
def scratch(tmp_path, monkeypatch):
    """cwd laid out like the reference expects: bad-states under data/, qvalues under src/data/."""
    d = tmp_path / "data" / "bad_states_files" / "layout-3"
    d.mkdir(parents=True)
    (d / "down_right.txt").write_text(bad_states_text(2))
    (d / "up_left.txt").write_text(bad_states_text(0))
    (tmp_path / "src" / "data" / "qvalue_files" / "layout-3").mkdir(parents=True)
    monkeypatch.chdir(tmp_path)
    return tmp_path


@pytest.mark.parametrize("name,episodes", [("traj_b2_2car_bad", 2), ("traj_b0_3car_bad", 2),
                                           ("traj_b2_2car_dqn_step", 3), ("traj_b2_4car_bad", 1)])
def test_environment_string_api_replays_reference(scratch, name, episodes):
    """The reference's own call sequence (qlearning.py:197-229 / Environment.step) against the facade."""
    import marl_dpp_b200 as m
    from marl_dpp_b200.env_gym import Environment
    g = load_golden(name)
    if not g["bad_states"]:
        os.remove(scratch / "data" / "bad_states_files" / "layout-3" / ("down_right.txt" if g["block"] == 2 else "up_left.txt"))
    env = Environment(m.layout_3[g["block"]])
    assert env.give_bad_state_reward == g["bad_states"]
    n = len(g["boxes"])
    use_step = name.endswith("dqn_step")
    for ep in g["episodes"][:episodes]:
        env.reset()
        env.initialize_cars(n)
        env.initialize(g["idx"], g["boxes"], g["choose"])
        env.given_nodes = list(ep["headings"])          # replay the recorded random.shuffle outcome
        env.reinitialize(g["idx"], g["boxes"], g["choose"])
        env.select_reward_function(g["reward"])
        it = iter(ep["steps"][:400])
        stop = False
        while not stop and not env.is_game_ended():
            for num in range(n):
                if env.check_car_training(num):
                    continue
                rec = next(it, None)
                if rec is None:
                    stop = True
                    break
                car, state, legal, action, nxt, reward = rec
                assert car == num
                assert env.encodestate(num, g["deadlock_id"]) == state
                assert "".join(env.getlegalactions(env.state_to_array(state))) == legal
                if use_step:
                    s, a, ns, r, arr = env.step(state, action, num)
                    assert (s, a, arr) == (state, action, env.state_to_array(nxt))
                else:
                    ns = env.get_successor_state(state, action)
                    r = env.reward_function(state, ns, num, action)
                    arr = env.state_to_array(ns)
                    env.update_car(num, arr[0][0], arr[0][1], 1)
                assert ns == nxt and r == reward and type(r) is type(reward), (state, action, r, reward)


def test_environment_known_answers(scratch):
    import marl_dpp_b200 as m
    from marl_dpp_b200.env_gym import Environment
    g = load_golden("known_answers")
    os.remove(scratch / "data" / "bad_states_files" / "layout-3" / "down_right.txt")
    env = Environment(m.layout_3[2])
    env.initialize_cars(2)
    env.initialize([0, 1], ["D5", "D8"], [[0, 1, 0], [0, 1, 0]])
    env.given_nodes = ["D5E", "D8N"]
    env.reinitialize([0, 1], ["D5", "D8"], [[0, 1, 0], [0, 1, 0]])
    env.update_car(1, "D8E", "D8E", 1)               # car 1 reaches its final destination
    cars = env.get_cars()
    assert cars[1].training_done == 1 and cars[1].source_node == g["ghost"]["parked"]
    assert cars[1].destination_index == g["ghost"]["parked_di"]
    for want_state, want_legal, want_count in g["ghost"]["encodes"]:
        s = env.encodestate(0, 1)
        assert s == want_state and "".join(env.getlegalactions(env.state_to_array(s))) == want_legal
        assert max(env.get_cars()[1].training_done_count, -1) == max(want_count, -1)
    assert env.encodestate(0, 0) == g["ghost"]["mode0"]
    assert env.get_successor_node("D6N", "F") == "D5N" and env.get_successor_node("D6E", "F") == "DYY"


def test_rl_plugin_contract(scratch):
    """settings.params + sess/env_config construction, train -> file -> reload -> greedy validation."""
    import marl_dpp_b200 as m
    from marl_dpp_b200 import app, formats, settings
    ed = app.give_configuration_object(3, 0)
    from marl_dpp_b200.algorithms.qlearning import RL
    rl = RL(**settings.params, sess=None, env_config=ed, n_envs=64, step_cap=2000)
    boxes, idx, choose = ed.initial_states[0]
    rl.train_given_state(60, '0', 'unit', boxes, len(boxes), idx, choose)
    assert rl.last_stats["episodes_done"] == 64 * 60
    path = formats.qvalue_path('unit')
    items = formats.read_qvalues(path)
    assert items and all(np.isfinite(v) for _, _, v in items)
    assert {(s, a) for s, a, _ in items} == set(rl.qvals)
    sc = rl.learner.env.scenario
    for s, a, v in items:
        assert sc.decode(sc.encode(s)) == s and a in A5
    # reload into a fresh plugin and roll the greedy policy out: the 1-car task must be solved
    rl2 = RL(**settings.params, sess=None, env_config=ed, n_envs=16, step_cap=500)
    res = rl2.validate_given_state(1, 'unit', '0', boxes, len(boxes), idx, choose, 0)
    assert res["success"] == 1.0 and res["capped"] == 0
    before = {k: v for k, v in rl2.qvals.items()}
    rl2._pull()
    assert rl2.qvals == before                         # frozen learner: the table did not move
    assert rl2.computeactionfromqvalues(sc.decode(int(rl2.learner.env.step_car(
        0, torch.tensor([0xFE] * 16, dtype=torch.uint8))["state"][0]))) in A5
    # the reference's -dd / -v flag surface
    out = app.main(["-i", "unit", "-o", "0", "-dd", "0", "-v", "1", "-l", "3", "-bn", "0", "--envs", "8"])
    assert out is not None


def test_learner_flags_match_oracle(scratch):
    """detect_deadlocks_v0 mode (encodestate(.,0) + stall heuristic) and the frozen mode, bit-exact
    against the batched oracle."""
    import marl_dpp_b200 as m
    from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner
    from oracle import oracle as orc
    boxes, idx, choose = ["D5", "D6", "D8"], [0, 0, 0], [[2, 1, 0], [0, 2, 0], [2, 1, 0]]
    bad = bad_states_text(2)
    n, episodes, cap = 256, 10, 200
    sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=cap)
    env = BatchedEnv(sc, n, seed=21)
    lr = BatchedQLearner(env, episodes=episodes, seed=21, flags=2 | 4)
    ob = orc.Batch(orc.Config(m.layout_3[2]), n, boxes, idx, choose, bad, cap)
    ob.reset(seed=21)
    oq = orc.QTable(f32=True)
    olp = orc.learner(episodes, seed=21)
    olp.flags = 2 | 4
    for r in range(400):
        for c in range(3):
            out = lr.substep(c, outputs=True)
            lr.finalize()
            want = ob.q_substep(oq, olp, c, r)
            assert np.array_equal(out["action"].cpu().numpy(), want["action"]), (r, c)
        lr.round += 1
    flagged = env.deadlock_states().cpu().numpy()
    L = orc._batch_lib()
    L.orc_batch_deadlock_state.restype = orc.C.POINTER(orc.State)
    L.orc_batch_deadlock_state.argtypes = [orc.C.c_void_p, orc.C.c_int64]
    L.orc_batch_deadlock_episode.restype = orc.C.c_int32
    L.orc_batch_deadlock_episode.argtypes = [orc.C.c_void_p, orc.C.c_int64]
    flag_episode = (env.records[:, 3].cpu().numpy().astype(np.int64) >> 16) & 0xFFFF
    n_flagged = 0
    for e in range(n):
        st = L.orc_batch_deadlock_state(ob.h, e).contents
        if st.pnode == orc.ORC_NONE:
            assert flagged[e] == -1 and L.orc_batch_deadlock_episode(ob.h, e) == -1
        else:
            assert sc.decode(int(flagged[e])) == orc.state_str(st)
            assert flag_episode[e] == L.orc_batch_deadlock_episode(ob.h, e)   # the value detect_deadlocks_v0 returns
            n_flagged += 1
    q = lr.q.cpu().numpy()
    for s, a, v in [x for x in oq.items() if x[1] != "B"]:
        assert q[sc.encode(s), A5.index(a)] == np.float32(v)
    gs, ws = env.stats(), ob.stats()
    assert gs["internal_errors"] == 0
    for k in ("agent_steps", "episodes_done", "episodes_capped", "explore_steps"):
        assert gs[k] == ws[k], k
    # frozen: same decisions, table untouched
    snap = lr.q.clone()
    env2 = BatchedEnv(sc, n, seed=22)
    lr2 = BatchedQLearner(env2, episodes=episodes, seed=22, flags=1, q=lr.q)
    lr2.train_rounds(50)
    assert torch.equal(lr2.q, snap) and env2.stats()["q_updates"] == 0 and env2.stats()["agent_steps"] > 0


def test_rl_detect_deadlocks_writes_bad_states(scratch):
    import marl_dpp_b200 as m
    from marl_dpp_b200 import app, formats, settings
    from marl_dpp_b200.algorithms.qlearning import RL
    ed = app.give_configuration_object(3, 2)
    path = ed.environment_bad_states_file
    before = set(formats.read_bad_states(path))
    rl = RL(**settings.params, sess=None, env_config=ed, n_envs=512, step_cap=300)
    res = rl.detect_deadlocks_v0(20, '0', '0', ["D5", "D6", "D8"], 3, [0, 0, 0], [[2, 1, 0], [0, 2, 0], [2, 1, 0]])
    after = set(formats.read_bad_states(path))
    assert before <= after
    for s in rl.detected_deadlocks:
        assert formats.strip_headings(s) in after
    # reference return value (qlearning.py:350-382): the detection episode (epsilon is 0 only in the last
    # 1 % of the episodes: 19 of 20) when a new state was flagged, the last episode index otherwise
    assert res == 19 and (rl.detected_deadlocks or after == before)
    # bad_states passed as TEXT with no file on disk: the file is created instead of raising
    os.remove(path)
    rl = RL(**settings.params, sess=None, env_config=ed, n_envs=512, step_cap=300, bad_states="\n".join(sorted(before)) + "\n")
    res = rl.detect_deadlocks_v0(20, '0', '0', ["D5", "D6", "D8"], 3, [0, 0, 0], [[2, 1, 0], [0, 2, 0], [2, 1, 0]])
    assert res == 19 and (not rl.detected_deadlocks or before <= set(formats.read_bad_states(path)))


@pytest.mark.parametrize("name", ["qrun_b0_1car", "qrun_b2_2car"])
def test_rl_plugin_reproduces_reference_run_from_the_same_seed(scratch, name):
    """The strongest drop-in statement: seed Python's `random` like the reference's recorded run and
    the plugin (one env, reference_rng) visits the same (state, action) sequence and ends with the same
    table (float32 vs the reference's float64: <= 1e-5 relative), with the env, the argmax and the TD
    update all on the GPU."""
    import random
    import marl_dpp_b200 as m
    from marl_dpp_b200 import settings
    from marl_dpp_b200.algorithms.qlearning import RL
    g = load_golden(name)
    os.remove(scratch / "data" / "bad_states_files" / "layout-3" / ("down_right.txt" if g["block"] == 2 else "up_left.txt"))
    rl = RL(**settings.params, sess=None, env_config=m.layout_3[g["block"]], n_envs=1, step_cap=0, reference_rng=True)
    random.seed(3)                                   # gen_qrun seeds with 3 right before train_given_state
    rl.train_given_state(g["episodes_total"], '0', '0', g["boxes"], len(g["boxes"]), g["idx"], g["choose"])
    sc = rl.learner.env.scenario
    want_trace = [(s[0], s[2]) for ep in g["log"] for s in ep["steps"]]
    got_trace = [(sc.decode(s), A5[a]) for s, a in rl.trace]
    assert got_trace == want_trace
    want = {(s, a): float(v) for s, a, v in g["qvals"]}
    updated = set(want_trace)
    assert set(rl.qvals) == updated
    for k, v in rl.qvals.items():
        assert q_close(v, want[k]), k


@pytest.mark.parametrize("name", ["qrand_b2_2car", "qrand_b2_3car_bad"])
def test_rl_plugin_reproduces_reference_random_order_run(scratch, name):
    """train_given_state_random_order seeded like the reference's recorded run (one env, reference_rng): the
    same cars are drawn, the same (state, action) sequence is visited and the same table is learnt (float32
    vs float64: <= 1e-5 relative)."""
    import random
    import marl_dpp_b200 as m
    from marl_dpp_b200 import settings
    from marl_dpp_b200.algorithms.qlearning import RL
    g = load_golden(name)
    bad_file = scratch / "data" / "bad_states_files" / "layout-3" / "down_right.txt"
    if not g["bad_states"]:
        os.remove(bad_file)
    rl = RL(**settings.params, sess=None, env_config=m.layout_3[g["block"]], n_envs=1, step_cap=0, reference_rng=True)
    random.seed(g["seed"])
    res = rl.train_given_state_random_order(g["episodes_total"], '0', '0', g["boxes"], len(g["boxes"]), g["idx"],
                                            g["choose"])
    assert res is None and g["result"] is None
    sc = rl.learner.env.scenario
    want_trace = [(s[0], s[1], s[3]) for ep in g["log"] for s in ep["steps"]]
    got_trace = [(c, sc.decode(s), A5[a]) for c, s, a in rl.trace]
    assert got_trace == want_trace
    want = {(s, a): float(v) for s, a, v in g["qvals"]}
    assert set(rl.qvals) == {(s, a) for _, s, a in want_trace}
    for k, v in rl.qvals.items():
        assert q_close(v, want[k]), k


def test_rl_plugin_random_order_batched(scratch):
    """The batched form: 256 envs, every env draws its own car; the run ends with every env through its
    episode budget, a table over the same states as the reference's run, and the file written."""
    import marl_dpp_b200 as m
    from marl_dpp_b200 import settings, formats
    from marl_dpp_b200.algorithms.qlearning import RL
    g = load_golden("qrand_b2_2car")
    os.remove(scratch / "data" / "bad_states_files" / "layout-3" / "down_right.txt")
    rl = RL(**settings.params, sess=None, env_config=m.layout_3[2], n_envs=256)
    res = rl.train_given_state_random_order(40, '0', '7', g["boxes"], 2, g["idx"], g["choose"])
    st = rl.last_stats
    assert st["episodes_done"] >= 256 * 40 and st["internal_errors"] == 0
    assert res is None and st["episodes_capped"] == 0
    ref_states = {s for s, a, v in g["qvals"]}
    got_states = {s for (s, a) in rl.qvals}
    assert len(ref_states & got_states) > 0.6 * len(ref_states)
    assert os.path.exists(formats.qvalue_path('7', rl.qvalue_dir))


def test_rl_counter_style_methods_match_oracle(scratch):
    """getaction / update / computevaluefromqvalues / getvalue / getpolicy on state strings (reference
    qlearning.py:34-121) against the float32 oracle table, on the (state, action, next, reward) tuples of a
    recorded reference run."""
    import random
    import marl_dpp_b200 as m
    from marl_dpp_b200 import settings
    from marl_dpp_b200.algorithms.qlearning import RL
    from oracle import oracle as orc
    g = load_golden("qrun_b2_2car_bad")
    rl = RL(**settings.params, sess=None, env_config=m.layout_3[2], n_envs=1)
    with pytest.raises(RuntimeError):
        rl.getaction("D2NxD6NzD3xD6", 0.0)
    rl.prepare(g["boxes"], 2, g["idx"], g["choose"])
    cfg = orc.Config(m.layout_3[2])
    env = orc.Env(cfg, bad_states_text(2))
    env.initialize_cars(2)
    oq = orc.QTable(f32=True)
    steps = [s for ep in g["log"][:6] for s in ep["steps"]][:400]
    for state, explore, action, nxt, reward, qafter in steps:
        st, ns = orc.state_from_str(state), orc.state_from_str(nxt)
        assert rl.computevaluefromqvalues(nxt) == oq.value(env, ns)
        rl.update(state, action, nxt, reward)
        oq.update(env, st, orc.ACTIONS.index(action), ns, reward)
        assert rl.getqvalue(state, action) == oq.get(st, orc.ACTIONS.index(action))
        assert rl.getpolicy(state) == orc.ACTIONS[oq.greedy(env, st)]
        assert rl.getvalue(state) == oq.value(env, st)
    assert rl.getaction(steps[0][0], 0.0) == rl.computeactionfromqvalues(steps[0][0])
    random.seed(5)
    legal = orc.mask_str(env.legal(orc.state_from_str(steps[0][0])))
    picks = {rl.getaction(steps[0][0], 1.0) for _ in range(60)}
    assert picks == set(legal) - {"B"}
