"""Timing experiment (ONE GPU): what the replica exchange costs INSIDE the kernel chain when no peer ever has to be
waited for (world = 2, the peer's table constant, every flag saturated): 256 rounds back to back per sync interval.
Separates the chain / launch effects of q_sync_kernel from cross-GPU latency and rank skew."""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner
from marl_dpp_b200.parallel import PeerExchange

bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))["down_right"]
boxes, idx, choose = m.layout_3[2]["initial_states"][0]
sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ex = PeerExchange.local(sc, world)
for r in range(world):
    ex.flags(r).fill_(0x7FFFFFFF)
env = BatchedEnv(sc, 1 << 20, seed=3)
lr = BatchedQLearner(env, episodes=5000, seed=3, q=ex.tables[0])
ex.attach(lr, 0, rank=0)
lr.train_rounds(256)
env.stats()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for beside in ("1", "0"):
  os.environ["MDPP_SYNC_BESIDE"] = beside
  print("exchange %s" % ("on its own stream beside the env kernel (default)" if beside == "1" else "in the caller's stream (MDPP_SYNC_BESIDE=0)"))
  for iv in (0, 1, 4, 64):
    ex.attach(lr, iv, rank=0)
    lr.round = ((lr.round + 63) // 64) * 64
    torch.cuda.synchronize()
    c0 = lr.sync_count()
    a.record()
    for _ in range(4):
        lr.train_rounds(64)
    b.record()
    torch.cuda.synchronize()
    env.stats()
    print("  interval %3d: %.2f us per round (%d exchanges)" % (iv, a.elapsed_time(b) / 256 * 1e3, lr.sync_count() - c0))
os.environ["MDPP_MULTI_ROUND"] = "0"
env2 = BatchedEnv(sc, 1 << 20, seed=3)
lr2 = BatchedQLearner(env2, episodes=5000, seed=3)
lr2.train_rounds(64)
torch.cuda.synchronize()
a.record()
for _ in range(4):
    lr2.train_rounds(64)
b.record()
torch.cuda.synchronize()
print("no exchange attached, one round per launch (MDPP_MULTI_ROUND=0): %.2f us per round" % (a.elapsed_time(b) / 256 * 1e3))
