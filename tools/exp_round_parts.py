"""Timing experiment (GPU): the round kernel alone (warm / cold L2) vs whole rounds (chained / isolated)."""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner

bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))["down_right"]
boxes, idx, choose = m.layout_3[2]["initial_states"][0]
sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def ev():
    return torch.cuda.Event(enable_timing=True)


def learner():
    env = BatchedEnv(sc, n, seed=3)
    lr = BatchedQLearner(env, episodes=5000, seed=3)
    lr.train_rounds(200)
    torch.cuda.synchronize()
    return env, lr


env, lr = learner()
a, b = ev(), ev()
a.record()
for _ in range(200):
    lr.profile_env_kernel()
b.record()
torch.cuda.synchronize()
print("round kernel alone, back to back, warm L2: %.2f us" % (a.elapsed_time(b) / 200 * 1e3))
ts = []
for _ in range(100):
    flush.zero_()
    a, b = ev(), ev()
    a.record(); lr.profile_env_kernel(); b.record()
    ts.append((a, b))
torch.cuda.synchronize()
print("round kernel alone, L2 flushed before each: %.2f us" % (sum(x.elapsed_time(y) for x, y in ts) / len(ts) * 1e3))
ts = []
for _ in range(100):
    a, b = ev(), ev()
    a.record(); lr.profile_env_kernel(); b.record()
    torch.cuda.synchronize()
    ts.append((a, b))
print("round kernel alone, isolated (sync between), warm L2: %.2f us" % (sum(x.elapsed_time(y) for x, y in ts) / len(ts) * 1e3))
del lr, env
env, lr = learner()
a, b = ev(), ev()
a.record(); lr.train_rounds(300); b.record()
torch.cuda.synchronize()
print("whole rounds, chained: %.2f us" % (a.elapsed_time(b) / 300 * 1e3))
ts = []
for _ in range(100):
    a, b = ev(), ev()
    a.record(); lr.train_rounds(1); b.record()
    torch.cuda.synchronize()
    ts.append((a, b))
print("whole round, isolated (sync between), warm L2: %.2f us" % (sum(x.elapsed_time(y) for x, y in ts) / len(ts) * 1e3))
ts = []
for _ in range(100):
    flush.zero_()
    a, b = ev(), ev()
    a.record(); lr.train_rounds(1); b.record()
    ts.append((a, b))
torch.cuda.synchronize()
print("whole round, L2 flushed before each: %.2f us" % (sum(x.elapsed_time(y) for x, y in ts) / len(ts) * 1e3))
# the same with the L2 left full of CLEAN lines (write flush, then a read pass over another 256 MiB)
flush2 = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
flush2.zero_()
ts = []
for _ in range(100):
    flush.zero_()
    sink = flush2.sum()
    a, b = ev(), ev()
    a.record(); lr.train_rounds(1); b.record()
    ts.append((a, b))
torch.cuda.synchronize()
print("whole round, L2 flushed (write pass then read pass: clean lines) before each: %.2f us" % (sum(x.elapsed_time(y) for x, y in ts) / len(ts) * 1e3))
a, b = ev(), ev()
torch.cuda.synchronize()
a.record(); b.record(); torch.cuda.synchronize()
print("empty event pair: %.2f us" % (a.elapsed_time(b) * 1e3))
