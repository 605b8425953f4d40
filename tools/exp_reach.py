"""Timing / capture helper (GPU): -dd reachability of the 4-car scenario, backward BFS (one cooperative launch) and
forward sweeps.  usage: exp_reach.py [reps=3]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sc = m.Scenario(m.layout_3[2], ["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]])
env = BatchedEnv(sc, 1024, seed=7)
for method in ("bfs", "sweep"):
    env.deadlock_reachability(method=method)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        bm, n = env.deadlock_reachability(method=method)
    torch.cuda.synchronize()
    print("%s: %.3f ms per call, %d levels/sweeps, 28.4 M joint states" % (method, (time.perf_counter() - t0) / reps * 1e3, n))
