"""Run the learner on one scenario (argv: n_cars) for profiling."""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner
bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))["down_right"]
CASES = {2: (["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]]),
         3: (["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]]),
         4: (["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]]),
         5: (["D1", "D2", "D3", "D5", "D7"], [0, 1, 0, 1, 0], [[0, 1, 0], [2, 1, 0], [1, 0, 0], [2, 2, 0], [3, 4, 0]])}
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 120
boxes, idx, choose = CASES[nc]
sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
env = BatchedEnv(sc, 1 << 20, seed=3)
lr = BatchedQLearner(env, episodes=5000, seed=3)
lr.train_rounds(rounds)
torch.cuda.synchronize()
print(env.stats())
