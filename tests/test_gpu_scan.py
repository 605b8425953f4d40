"""GPU parity of the large-table learner (scan mode: 3-5 cars): touched bitmaps in global memory, a mirror of the
table, one scanning pull per sub-step, and all cars of several pure-exploration rounds in one env pass
(q_rounds_pure_scan_kernel).  Compared with the CPU oracle's sub-step-by-sub-step restatement, bit for bit:
table, updated-key bits, every env's record, counters -- over runs that start in the epsilon = 1 segment (fused
passes), leave it (sub-step kernels), end with frozen envs, and are cut into calls of different sizes; the caller
also rewrites the table between two calls (mdpp_q_table_written through BatchedQLearner.q)."""
import os

import numpy as np
import pytest

from conftest import bad_states_text
from test_gpu_fullsize import _assert_counters_equal, _assert_records_equal, _assert_table_equal

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

A5 = "SLRFT"
CASES = [
    # name, boxes, dest_index, choose, force scan mode, envs, episodes, step cap
    ("b2_1car_forced_scan", ["D3"], [0], [[0, 1, 0]], True, 1000, 30, 40),
    ("b2_2car_forced_scan", ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], True, 2049, 24, 44),
    ("b2_3car_forced_scan", ["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]], True, 3001, 20, 45),
    # mid-size table, default: shared-memory bit maps + dense pulls for rounds with greedy decisions, the scan mode's
    # fused pass + scanning pulls for pure-exploration rounds, a table copy / settle pass at every change-over
    ("b2_3car_hybrid", ["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]], False, 3001, 20, 45),
    ("b2_4car", ["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]], False, 4100, 16, 48),
    ("b2_5car", ["D1", "D2", "D3", "D5", "D7"], [0, 1, 0, 1, 0], [[0, 1, 0], [2, 1, 0], [1, 0, 0], [2, 2, 0], [3, 4, 0]],
     False, 2500, 12, 50),
]


def _mods():
    import marl_dpp_b200 as m
    from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner
    from oracle import oracle as orc
    orc.set_threads()
    return m, BatchedEnv, BatchedQLearner, orc


def _make(case, seed, pure_scan=True, overlap=True):
    m, BatchedEnv, BatchedQLearner, orc = _mods()
    name, boxes, idx, choose, force, n, episodes, cap = case
    badtxt = bad_states_text(2)
    sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=badtxt, step_cap=cap)
    if force:
        os.environ["MDPP_DENSE"] = "0"
    if not pure_scan:
        os.environ["MDPP_PURE_SCAN"] = "0"
    if not overlap:
        os.environ["MDPP_SCAN_OVERLAP"] = "0"
    try:
        env = BatchedEnv(sc, n, seed=seed)
    finally:
        os.environ.pop("MDPP_DENSE", None)
        os.environ.pop("MDPP_PURE_SCAN", None)
        os.environ.pop("MDPP_SCAN_OVERLAP", None)
    lr = BatchedQLearner(env, episodes=episodes, seed=seed)
    ob = orc.Batch(orc.Config(m.layout_3[2]), n, boxes, idx, choose, badtxt, cap)
    ob.reset(seed=seed)
    return sc, env, lr, ob, orc.QTable(f32=True), orc.learner(episodes, seed=seed), orc


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_scan_mode_rounds_match_oracle(case):
    sc, env, lr, ob, oq, olp, orc = _make(case, seed=11)
    done = 0

    def run(k):
        nonlocal done
        lr.train_rounds(k)
        for r in range(done, done + k):
            ob.q_round(oq, olp, r)
        done += k

    # fused passes while every env certainly explores (the host's bound grows by one per round and is tightened by
    # every stats fetch): groups that fill a launch, leave a remainder, and single rounds
    for k in (5, 1, 3):
        run(k)
    _assert_table_equal(lr, sc, oq)
    _assert_records_equal(env, ob)
    _assert_counters_equal(env, ob)               # (fetching the counters tightens the bound)
    run(7)
    _assert_table_equal(lr, sc, oq)
    # the caller rewrites the table between two calls: the mirror must follow
    lr.q.mul_(0.5)
    for s, a, v in oq.items():
        oq.set(orc.state_from_str(s), orc.ACTIONS.index(a), float(np.float32(v) * np.float32(0.5)))
    run(2)
    _assert_table_equal(lr, sc, oq)
    # through every epsilon segment to frozen envs
    episodes, cap = case[6], case[7]
    while int(env.cars()["episode"].min()) < episodes:      # (an episode lasts at most `cap` rounds)
        assert done < episodes * (cap + 2)
        run(37)
        env.stats()
    _assert_table_equal(lr, sc, oq)
    _assert_records_equal(env, ob)
    _assert_counters_equal(env, ob)
    assert int(env.cars()["episode"].min()) == episodes     # every env used its budget: frozen
    assert ob.disagreements() == 0
    assert env.stats()["touched_keys"] == ob.touched_keys()


def test_scan_mode_grouping_and_fused_pass_do_not_change_the_result():
    """The same 4-car run as one call, as single-round calls, with the fused pass switched off, and with the pulls on
    the caller's stream instead of beside the next env pass."""
    case = [c for c in CASES if c[0] == "b2_4car"][0]
    tables, records = [], []
    for pure_scan, overlap, groups in ((True, True, [24]), (True, True, [1] * 24), (False, True, [24]),
                                       (True, True, [3, 8, 2, 11]), (True, False, [24])):
        sc, env, lr, ob, oq, olp, orc = _make(case, seed=4, pure_scan=pure_scan, overlap=overlap)
        for k in groups:
            lr.train_rounds(k)
        tables.append(lr.q.cpu().numpy().copy())
        records.append(env.records.cpu().numpy().copy())
        assert env.stats()["internal_errors"] == 0
    for t, r in zip(tables[1:], records[1:]):
        assert np.array_equal(tables[0], t)
        assert np.array_equal(records[0], r)
