"""Launch-list helper (GPU): a few multi-round groups of the 2-car learner, for `ncu --metrics gpu__time_duration.sum`."""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner
bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))["down_right"]
boxes, idx, choose = m.layout_3[2]["initial_states"][0]
sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
env = BatchedEnv(sc, 1 << 20, seed=3)
lr = BatchedQLearner(env, episodes=5000, seed=3)
lr.train_rounds(int(sys.argv[1]) if len(sys.argv) > 1 else 48)
torch.cuda.synchronize()
print(env.stats()["agent_steps"])
