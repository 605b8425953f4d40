"""Timing experiment (GPU): random-policy rollout kernel, n cars x n envs, with / without materialised outputs.
usage: exp_rollout.py [n_cars=4] [log2_envs=24] [reps=10] [preroll_rounds=1024]"""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv

bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))["down_right"]
STARTS = {1: (["D3"], [0], [[0, 1, 0]]),
          2: m.layout_3[2]["initial_states"][0],
          3: (["D2", "D6", "D4"], [0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0]]),
          4: (["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]]),
          5: (["D2", "D6", "D4", "D8", "D1"], [0, 0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0], [0, 1, 0]])}
k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 24
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
boxes, idx, choose = STARTS[k]
sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
n = 1 << lg
env = BatchedEnv(sc, n, seed=5)
pre = int(sys.argv[4]) if len(sys.argv) > 4 else 1024   # untimed rounds first: the stationary mix of active / finished cars
for w in range(pre // 16):
    env.rollout_random(16, w * 16)
base = (pre // 16) * 16
outs = env.alloc_rollout_outputs(1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for label, o, rounds in (("outputs, 1 round/launch", outs, 1), ("no outputs, 1 round/launch", None, 1),
                         ("no outputs, 16 rounds/launch", None, 16)):
    for w in range(3):
        env.rollout_random(rounds, base + w * rounds, outputs=o)
    torch.cuda.synchronize()
    env.stats(reset=True)
    a.record()
    for w in range(reps):
        env.rollout_random(rounds, base + (3 + w) * rounds, outputs=o)
    base += (3 + reps) * rounds
    b.record()
    torch.cuda.synchronize()
    s = env.stats(reset=True)
    ms = a.elapsed_time(b)
    v = s["agent_steps"] / (ms * 1e-3)
    print("%d cars 2^%d envs, %s: %.3f ms/launch  %.3e agent-steps/s  (%.0f GB/s algorithmic with outputs; active fraction %.3f)"
          % (k, lg, label, ms / reps, v, (10.0 + 16.0 / k) * v / 1e9, s["agent_steps"] / float(reps * rounds * n * k)), flush=True)
