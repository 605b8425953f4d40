"""Timing experiment (GPU): random-policy rollout (4 cars, outputs on) and 2-car Q-learning vs env count."""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner

bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))["down_right"]
sc4 = m.Scenario(m.layout_3[2], ["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]],
                 bad_states_text=bad, step_cap=10000)
b2, i2, c2 = m.layout_3[2]["initial_states"][0]
sc2 = m.Scenario(m.layout_3[2], b2, i2, c2, bad_states_text=bad, step_cap=10000)
for logn in (20, 22, 24, 26):
    n = 1 << logn
    env = BatchedEnv(sc4, n, seed=5)
    outs = env.alloc_rollout_outputs(1)
    for w in range(3):
        env.rollout_random(1, w, outputs=outs)
    torch.cuda.synchronize(); env.stats(reset=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for w in range(8):
        env.rollout_random(1, 3 + w, outputs=outs)
    b.record(); torch.cuda.synchronize()
    s = env.stats(); ms = a.elapsed_time(b)
    r = s["agent_steps"] / (ms * 1e-3)
    del env, outs
    env = BatchedEnv(sc2, n, seed=3)
    lr = BatchedQLearner(env, episodes=5000, seed=3)
    lr.train_rounds(60); torch.cuda.synchronize(); env.stats(reset=True)
    a.record(); lr.train_rounds(40); b.record(); torch.cuda.synchronize()
    s = env.stats(); ms2 = a.elapsed_time(b)
    print("2^%d envs: rollout 4 cars + outputs %.2e steps/s (%.0f GB/s alg.) | q-learning 2 cars %.2e steps/s (%.1f us/round)"
          % (logn, r, 14 * r / 1e9, s["agent_steps"] / (ms2 * 1e-3), ms2 * 1e3 / 40), flush=True)
    del lr, env
