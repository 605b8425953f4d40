"""Timing experiment (GPU): learner throughput across scenarios (1..5 cars, small and large tables)."""
import gzip, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner

bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))
CASES = [
    ("b0 1 car", 0, ["U4"], [1], [[0, 0, 0]]),
    ("b2 2 cars", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]]),
    ("b2 3 cars", 2, ["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]]),
    ("b2 4 cars", 2, ["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]]),
    ("b2 5 cars", 2, ["D1", "D2", "D3", "D5", "D7"], [0, 1, 0, 1, 0], [[0, 1, 0], [2, 1, 0], [1, 0, 0], [2, 2, 0], [3, 4, 0]]),
]
n = 1 << 20
for name, block, boxes, idx, choose in CASES:
    sc = m.Scenario(m.layout_3[block], boxes, idx, choose,
                    bad_states_text=bad["down_right" if block == 2 else "up_left"], step_cap=10000)
    env = BatchedEnv(sc, n, seed=3)
    lr = BatchedQLearner(env, episodes=5000, seed=3)
    lr.train_rounds(100)
    torch.cuda.synchronize()
    env.stats(reset=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); lr.train_rounds(100); b.record(); torch.cuda.synchronize()
    st = env.stats()
    ms = a.elapsed_time(b)
    print("%-10s states %9d  ws %6.0f MB  round %7.1f us  %.2e agent-steps/s  touched/substep %.0f" %
          (name, sc.n_states, env.workspace.numel() / 2**20, ms * 10, st["agent_steps"] / (ms * 1e-3),
           st["touched_keys"] / (100 * len(boxes))), flush=True)
    del lr, env
