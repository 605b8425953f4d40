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


def test_unique_replay_equals_the_references_ordered_set():
    """remember() + `deque(OrderedSet(self.memory))` (reference dqn_open_source.py:322,343) on the device: after a
    few thousand envs x rounds the de-duplicated memory holds exactly the distinct (state, action, next_state,
    reward) tuples a Python set of the same samples holds, its key list is sorted, every payload is the tuple's,
    and a minibatch is `batch_size` distinct members (None while the set is smaller than the batch, :344-345)."""
    import marl_dpp_b200 as m
    from marl_dpp_b200.batched import BatchedEnv
    from marl_dpp_b200.dqn_feed import DQNCollector, ReplayRing, UniqueReplay, random_legal_policy
    sc = m.Scenario(m.layout_3[2], ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], step_cap=500)
    n = 4096
    env = BatchedEnv(sc, n, seed=2)
    ring = ReplayRing(2 * n, "cuda")
    mem = UniqueReplay(sc, "cuda")
    assert mem.sample(32) is None and len(mem) == 0
    g = torch.Generator(device="cuda").manual_seed(5)
    col = DQNCollector(env, ring, random_legal_policy(g), memory=mem)
    seen = {}
    offered = 0
    for _ in range(40):
        # replay the collector's sub-steps by hand so that every sample can also go into a Python set
        for c in range(env.n_cars):
            sl = ring.reserve(n)
            feats, legal, state = env.dqn_observe(c, features=ring.state[sl])
            actions = col.policy(feats, legal)
            actions = torch.where(legal != 0, actions, torch.full_like(actions, 0xFF))
            out = env.dqn_act(c, actions, next_features=ring.next_state[sl], auto_reset=True, round=col.round)
            ok = out["status"] == 0
            nxt = env.query(c, torch.where(ok, state, torch.zeros_like(state)),
                            torch.where(ok, actions, torch.full_like(actions, 0xFF)), reward="dqn_reward")["next_state"]
            mem.insert(state, actions, nxt, out["reward"], out["next_legal"], out["status"], ring.state[sl], ring.next_state[sl])
            okc = ok.cpu().numpy()
            tup = zip(state.cpu().numpy()[okc].tolist(), actions.cpu().numpy()[okc].tolist(),
                      nxt.cpu().numpy()[okc].tolist(), out["reward"].cpu().numpy()[okc].tolist(),
                      out["next_legal"].cpu().numpy()[okc].tolist(),
                      [bytes(r) for r in ring.state[sl].cpu().numpy()[okc]],
                      [bytes(r) for r in ring.next_state[sl].cpu().numpy()[okc]])
            for s_, a_, n_, r_, mk, f1, f2 in tup:
                offered += 1
                seen.setdefault((s_, a_, n_, r_), (mk, f1, f2))
        col.round += 1
    st = mem.stats()
    assert st["offered"] == offered and st["differing_duplicates"] == 0
    assert len(mem) == st["unique"] == len(seen) > 1000
    assert len({(s_, a_) for s_, a_, _, _ in seen}) == len(seen)        # next state and reward follow from the key
    keys = mem.keys[:len(mem)].cpu().numpy()
    assert np.array_equal(keys, np.array(sorted(s_ * 5 + a_ for s_, a_, _, _ in seen)))
    nxt_all, rew_all, mask_all = mem.next.cpu().numpy(), mem.reward.cpu().numpy(), mem.mask.cpu().numpy()
    feat_all = mem.features.cpu().numpy()
    for (s_, a_, n_, r_), (mk, f1, f2) in seen.items():
        k = s_ * 5 + a_
        assert (nxt_all[k], rew_all[k], mask_all[k]) == (n_, np.float32(r_), mk)
        assert bytes(feat_all[k, :80]) == f1 and bytes(feat_all[k, 80:]) == f2
    # the same set as strings, the way the reference's OrderedSet sees it
    strings = {(sc.decode(s_), "SLRFT"[a_], sc.decode(n_), r_) for s_, a_, n_, r_ in seen}
    assert len(strings) == len(seen)
    batch = mem.sample(32, generator=g)
    assert len(set(batch["key"].tolist())) == 32
    for k, s_, a_, n_, r_ in zip(batch["key"].tolist(), batch["state"].tolist(), batch["action"].tolist(),
                                 batch["next_state"].tolist(), batch["reward"].tolist()):
        assert k == s_ * 5 + a_ and (s_, a_, n_, r_) in seen
    assert batch["features"].shape == (32, 80) and batch["next_features"].shape == (32, 80)
    assert mem.sample(len(mem) + 1) is None


@pytest.mark.parametrize("index_mode", ["reference", "action"])
def test_double_dqn_target_matches_numpy_restatement(index_mode):
    """The Double-DQN target (reference dqn_open_source.py:360-379) against its NumPy statement, line for line
    (TensorFlow is absent, so the two networks' outputs are random float32 arrays): bit-identical y and target,
    the same invalid_count.  index_mode "reference" keeps the reference's use of the POSITION inside the masked
    sub-array as the action index; "action" maps it back."""
    from marl_dpp_b200.dqn_feed import double_dqn_target
    rng = np.random.default_rng(3)
    n, discount = 4096, 0.99
    model_predictions = rng.normal(size=(n, 5)).astype(np.float32)       # self.predict(next_states)
    model_predictions[::7, 1] = model_predictions[::7, 3]                # ties: np.argmax takes the first
    target_predictions = rng.normal(size=(n, 5)).astype(np.float32)      # self.target_predict(next_states)
    predictions = rng.normal(size=(n, 5)).astype(np.float32)             # self.predict(states)
    bits = rng.integers(1, 32, size=n).astype(np.uint8)
    masks = [np.nonzero([(b >> k) & 1 for k in range(5)])[0] for b in bits]   # self.masks[next_state]
    rewards = rng.choice(np.array([-110, -60, -10, 40, 100, -210]), size=n) / 100   # float64 k / 100 (env_gym.py:553)
    actions = rng.integers(0, 5, size=n)
    # ---- the reference's lines ----
    max_actions, invalid_count = [], 0
    for i, qvalues in enumerate(model_predictions):
        overall_max_action = np.argmax(qvalues)
        max_action = np.argmax(qvalues[masks[i]])
        max_actions.append(max_action if index_mode == "reference" else masks[i][max_action])
        if overall_max_action != max_action:
            invalid_count += 1
    max_actions = np.array(max_actions)
    max_qvalues = np.choose(max_actions, target_predictions.T)
    true = rewards + discount * max_qvalues
    want = predictions.copy()
    for i in range(len(want)):
        want[i][actions[i]] = true[i]
    # ---- the kernel ----
    counters = torch.zeros(2, dtype=torch.int64, device="cuda")
    dev = lambda a: torch.from_numpy(a).cuda()
    y, t = double_dqn_target(dev(model_predictions), dev(target_predictions), dev(bits), dev(rewards.astype(np.float32)),
                             dev(actions), dev(predictions), discount, index_mode=index_mode, counters=counters)
    # the kernel gets the float32 rewards the env kernels emit and recovers the reference's float64 k / 100 from them;
    # `discount * max_qvalues` is a float32 product in NumPy (Python scalar x float32 array), the sum is float64
    assert (discount * max_qvalues).dtype == np.float32 and true.dtype == np.float64
    assert np.array_equal(y.cpu().numpy(), want)
    assert np.array_equal(t.cpu().numpy(), true.astype(np.float32))
    assert counters.tolist() == [invalid_count, 0]
