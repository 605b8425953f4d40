"""Launch-list helper (GPU): 2-car learner with epsilon pinned (argv[1], default 0.5): the K0 / pull / K1 / pull rounds."""
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
lr.train_rounds(200)
eps = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
for i in range(6):
    lr.params.eps_q16[i] = int(eps * 65536)
lr.train_rounds(100)
torch.cuda.synchronize()
env.stats(reset=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); lr.train_rounds(200); b.record()
torch.cuda.synchronize()
s = env.stats()
print("eps %.2f: %.2f us/round, %.3e agent-steps/s, greedy fraction %.3f" % (eps, a.elapsed_time(b) / 200 * 1e3, s["agent_steps"] / (a.elapsed_time(b) * 1e-3), 1 - s["explore_steps"] / max(s["agent_steps"], 1)))
