"""Timing experiment (GPU): where a learner round goes -- sub-steps alone, finalize alone, full round."""
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
lr.train_rounds(150)
torch.cuda.synchronize()


def timed(fn, reps=200):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def full():
    lr.train_rounds(1)


def substeps_only():
    lr.substep(0); lr.substep(1); lr.round += 1


def finalize_only():
    lr.finalize(); lr.finalize()


print("full round        %.1f us" % timed(full))
print("2 sub-steps only  %.1f us" % timed(substeps_only))
lr.finalize()
print("2 finalizes only  %.1f us" % timed(finalize_only))
print("full round again  %.1f us" % timed(full))
