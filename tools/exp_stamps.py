"""Phase timing of q_rounds_resident_kernel (GPU): runs rounds with epsilon pinned (argv[1], default 0.5) under
MDPP_STAMPS=1 and prints, for three CTAs, the time of every phase of the first sub-steps of the last launch."""
import ctypes as C, gzip, json, os, sys
os.environ["MDPP_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200 import _lib
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
lr.train_rounds(160)
torch.cuda.synchronize()
buf = (C.c_uint64 * 96)()
_lib.check(_lib.lib().mdpp_debug_stamps(buf, 96))
names = ["pass", "flush", "barrier", "pull", "clear"]
for cta in range(3):
    t = list(buf[cta * 32:(cta + 1) * 32])
    print("CTA slot %d: staged at 0" % cta)
    for sub in range(6):
        parts = []
        for k in range(5):
            a, b = t[5 * sub + k], t[5 * sub + k + 1]
            parts.append("%s %.2f" % (names[k], (b - a) / 1e3))
        print("  sub-step %d: %s us   (ends at %.2f)" % (sub, "  ".join(parts), (t[5 * sub + 5] - t[0]) / 1e3))
