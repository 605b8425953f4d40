import gzip, json, os, sys
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner
bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))["down_right"]
boxes, idx, choose = m.layout_3[2]["initial_states"][0]
sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
env = BatchedEnv(sc, 1 << 20, seed=3)
lr = BatchedQLearner(env, episodes=5000, seed=3)
lr.train_rounds(200)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
L, h = env._L, env._h
vis = L.mdpp_visited_bits(h) - env.workspace.data_ptr()
ctl_off = (vis + 4 * sc.n_states + 255) & ~255
dbg_off = ctl_off + 128 * 64 + 64 + 512
dbg = env.workspace[dbg_off:dbg_off + 128].view(torch.int64)
MIN = [0, 2, 3, 9, 10]
for cold in (1, 0):
    acc = []
    for it in range(30):
        if cold: flush.zero_()
        torch.cuda.synchronize()
        dbg[:] = 0
        for i in MIN: dbg[i] = 1 << 62
        torch.cuda.synchronize()
        lr.train_rounds(1)
        torch.cuda.synchronize()
        d = dbg.cpu().tolist()
        acc.append([x - d[0] for x in d])
    mm = np.median(np.array(acc[5:], dtype=np.float64), axis=0) / 1e3
    print("cold" if cold else "warm", "K0 end %.1f" % mm[1])
    for name, b in (("pull0", 2), ("pull1", 9)):
        print("   %s: first resident %.1f, first past wait %.1f, last past wait %.1f, last bitmaps reduced %.1f, last stores issued %.1f, last exit %.1f"
              % (name, mm[b], mm[b + 1], mm[b + 2], mm[b + 3], mm[b + 4], mm[b + 5]))
