"""Timing experiment (GPU): the large-table learner (scan mode, 3-5 cars) at 2^20 envs -- pure-exploration rounds
(fused env pass + one scanning pull per sub-step) and rounds with epsilon pinned to 0.5 (sub-step kernels).
usage: exp_scan.py [n_cars=4] [log2_envs=20] [rounds=128] [preroll_rounds=512]      env: MDPP_PURE_SCAN=0 -> sub-step kernels throughout"""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner

bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))["down_right"]
STARTS = {2: m.layout_3[2]["initial_states"][0],
          3: (["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]]),
          4: (["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]]),
          5: (["D1", "D2", "D3", "D5", "D7"], [0, 1, 0, 1, 0], [[0, 1, 0], [2, 1, 0], [1, 0, 0], [2, 2, 0], [3, 4, 0]])}
k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 128
boxes, idx, choose = STARTS[k]
sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
n = 1 << lg
env = BatchedEnv(sc, n, seed=3)
lr = BatchedQLearner(env, episodes=5000, seed=3)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(label, r, per_call=None):
    torch.cuda.synchronize()
    env.stats(reset=True)
    l0 = env.launch_count()
    a.record()
    for _ in range(r // (per_call or r)):
        lr.train_rounds(per_call or r)
    b.record()
    torch.cuda.synchronize()
    s = env.stats()
    ms = a.elapsed_time(b)
    print("%d cars 2^%d envs, %s: %.1f us/round  %.3e agent-steps/s  %.0f%% of the HBM roofline  active %.3f  touched/sub-step %.0f  "
          "launches/round %.1f  greedy %.3f" % (k, lg, label, ms / r * 1e3, s["agent_steps"] / (ms * 1e-3),
          100 * (44.0 + 16.0 / k) * s["agent_steps"] / (ms * 1e-3) / 6545.9e9, s["agent_steps"] / float(r * n * k),
          s["touched_keys"] / float(r * k), (env.launch_count() - l0) / float(r), 1 - s["explore_steps"] / max(s["agent_steps"], 1)),
          flush=True)


lr.train_rounds(int(sys.argv[4]) if len(sys.argv) > 4 else 512)   # to the stationary mix of active and finished cars
timed("pure exploration", rounds)
timed("pure exploration, one round per call", 32, 1)
for i in range(6):
    lr.params.eps_q16[i] = 32768
lr.train_rounds(16)
timed("epsilon 0.5", rounds)
assert env.stats()["internal_errors"] == 0
