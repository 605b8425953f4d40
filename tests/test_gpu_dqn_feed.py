"""GPU: DQN feed (convert featurizer, give_mask, dqn_reward through observe/act) against the oracle's
restatement, which is pinned to reference-recorded RL.convert outputs."""
import numpy as np
import pytest

from conftest import bad_states_text, load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
A6 = [0, 1, 2, 3, 5]

CASES = [
    (2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], 96, 80),
    (2, ["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]], 64, 120),
    (2, ["D1", "D2", "D3", "D5", "D7"], [0, 1, 0, 1, 0], [[0, 1, 0], [2, 1, 0], [1, 0, 0], [2, 2, 0], [3, 4, 0]], 40, 150),
    (0, ["U4", "U8", "U7"], [0, 1, 1], [[0, 0, 0]] * 3, 50, 80),
]


@pytest.mark.parametrize("block,boxes,idx,choose,n,rounds", CASES)
def test_dqn_feed_matches_oracle(block, boxes, idx, choose, n, rounds):
    import marl_dpp_b200 as m
    from marl_dpp_b200.batched import BatchedEnv
    from oracle import oracle as orc
    bad = bad_states_text(block)
    sc = m.Scenario(m.layout_3[block], boxes, idx, choose, bad_states_text=bad, step_cap=0)
    env = BatchedEnv(sc, n, seed=31)
    nc = len(boxes)
    ocfg = orc.Config(m.layout_3[block])
    heads = env.cars()["node"].cpu().numpy() & 3
    oenvs = []
    for e in range(n):
        oe = orc.Env(ocfg, bad)
        oe.initialize_cars(nc)
        oe.initialize(idx, [orc.box_id(b) * 4 + int(heads[e, i]) for i, b in enumerate(boxes)], choose)
        oenvs.append(oe)
    checked = 0
    for r in range(rounds):
        for c in range(nc):
            feats, legal, state = env.dqn_observe(c)
            feats, legal, state = feats.cpu().numpy(), legal.cpu().numpy(), state.cpu().numpy()
            acts = np.full(n, 0xFF, dtype=np.uint8)
            want_next = {}
            for e in range(n):
                oe = oenvs[e]
                if oe.car_done(c):
                    assert legal[e] == 0 and not feats[e].any()
                    continue
                st = oe.encodestate(c, 1)
                assert sc.decode(int(state[e])) == orc.state_str(st)
                assert list(feats[e]) == orc.convert(st), (r, c, e, orc.state_str(st))
                lm6 = oe.legal(st)
                lm5 = sum(((lm6 >> a6) & 1) << i for i, a6 in enumerate(A6))
                assert legal[e] == lm5
                if lm5 == 0:
                    continue
                options = [i for i in range(5) if (lm5 >> i) & 1]
                a = options[(e * 7 + r * 3 + c) % len(options)]
                acts[e] = a
                ns = oe.successor_state(st, A6[a])
                rw = oe.reward(st, ns, c, True)
                lm6n = oe.legal(ns)
                want_next[e] = (orc.convert(ns), rw, sum(((lm6n >> a6) & 1) << i for i, a6 in enumerate(A6)))
                oe.update_car(c, ns.pnode)
            out = env.dqn_act(c, torch.from_numpy(acts), reward="dqn_reward")
            out = {k: v.cpu().numpy() for k, v in out.items()}
            for e in range(n):
                if e in want_next:
                    f, rw, lm = want_next[e]
                    assert out["status"][e] == 0
                    assert list(out["next_features"][e]) == f
                    assert out["reward"][e] == np.float32(rw)
                    assert out["next_legal"][e] == lm
                    assert bool(out["done"][e]) == oenvs[e].is_game_ended()
                    checked += 1
                else:
                    assert out["status"][e] == 1 and not out["next_features"][e].any()
        # envs that finished are restarted on both sides with the same headings
        done = env.cars()["done"].cpu().numpy().all(axis=1)
        if done.any():
            hd = np.random.default_rng(r).integers(0, 4, size=(n, nc)).astype(np.uint8)
            env.reset(headings=torch.from_numpy(hd), mask=torch.from_numpy(done.astype(np.uint8)), zero_counters=False)
            for e in np.nonzero(done)[0]:
                oe = oenvs[e]
                oe.reset()
                oe.initialize_cars(nc)
                oe.initialize(idx, [orc.box_id(b) * 4 + int(hd[e, i]) for i, b in enumerate(boxes)], choose)
    assert checked > n * nc * rounds // 4


def test_collector_fills_ring_and_feeds_a_torch_model():
    import marl_dpp_b200 as m
    from marl_dpp_b200.batched import BatchedEnv
    from marl_dpp_b200.dqn_feed import DQNCollector, ReplayRing, masked_greedy_policy, random_legal_policy
    sc = m.Scenario(m.layout_3[2], ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], step_cap=500)
    n = 4096
    env = BatchedEnv(sc, n, seed=1)
    ring = ReplayRing(8 * n, "cuda")
    col = DQNCollector(env, ring, random_legal_policy())
    for _ in range(6):
        col.collect_round()
    assert ring.filled == 8 * n and ring.valid.any()
    s, a, ns, r, nm = ring.sample(256)
    assert s.shape == (256, 80) and ns.shape == (256, 80) and set(torch.unique(s).tolist()) <= {0, 1}
    assert (a < 5).all() and ((nm >> a.to(torch.uint8)) >= 0).all()
    assert set(np.round(torch.unique(r).cpu().numpy() * 100).astype(int).tolist()) <= {-110, -60, -10, 40, 100}
    # the reference's network shape (80-24-24-24-24-5), in PyTorch, driven by the masked greedy policy
    model = torch.nn.Sequential(torch.nn.Linear(80, 24), torch.nn.ReLU(), torch.nn.Linear(24, 24), torch.nn.ReLU(),
                                torch.nn.Linear(24, 24), torch.nn.ReLU(), torch.nn.Linear(24, 24), torch.nn.ReLU(),
                                torch.nn.Linear(24, 5)).cuda()
    col2 = DQNCollector(env, ring, masked_greedy_policy(model))
    before = env.stats()["agent_steps"]
    col2.collect_round()
    assert env.stats()["agent_steps"] > before
