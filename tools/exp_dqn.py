"""Run the DQN feed kernels (observe + act) on 2^24 envs for profiling / timing."""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv
bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))["down_right"]
boxes, idx, choose = m.layout_3[2]["initial_states"][0]
sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
n = 1 << 24
env = BatchedEnv(sc, n, seed=6)
dev = env.device
f_s = torch.empty((n, 80), dtype=torch.uint8, device=dev)
f_n = torch.empty((n, 80), dtype=torch.uint8, device=dev)
lg = torch.empty(n, dtype=torch.uint8, device=dev)
sid = torch.empty(n, dtype=torch.int32, device=dev)
o = {"reward": torch.empty(n, dtype=torch.float32, device=dev), "next_legal": torch.empty(n, dtype=torch.uint8, device=dev),
     "done": torch.empty(n, dtype=torch.uint8, device=dev), "status": torch.empty(n, dtype=torch.uint8, device=dev)}
ms = 0.0
for w in range(6):
    if w == 2:
        env.stats(reset=True)
    for c in range(2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(); env.dqn_observe(c, features=f_s, legal=lg, state=sid); e[1].record()
        acts = torch.where(lg != 0, torch.log2((lg & (0 - lg)).float()).to(torch.uint8), torch.full_like(lg, 0xFF))
        e[2].record(); env.dqn_act(c, acts, next_features=f_n, out=o, auto_reset=True, round=w); e[3].record()
        torch.cuda.synchronize()
        if w >= 2:
            ms += e[0].elapsed_time(e[1]) + e[2].elapsed_time(e[3])
            print("car %d observe %.1f us act %.1f us" % (c, e[0].elapsed_time(e[1]) * 1e3, e[2].elapsed_time(e[3]) * 1e3))
s = env.stats()
print("%.3e steps/s, %.0f GB/s algorithmic" % (s["agent_steps"] / (ms * 1e-3), 174.0 * s["agent_steps"] / (ms * 1e-3) / 1e9))
