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
