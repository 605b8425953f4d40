"""Timing experiment (GPU): q_substep / finalize duration vs accumulator replicas and training phase."""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner

bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))["down_right"]
boxes, idx, choose = m.layout_3[2]["initial_states"][0]
sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
n = 1 << 20
reps_list = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,8,64").split(",")]
for reps in reps_list:
    os.environ["MDPP_REPLICAS"] = str(reps)
    env = BatchedEnv(sc, n, seed=3)
    lr = BatchedQLearner(env, episodes=5000, seed=3)
    out = []
    for phase, warm in (("start", 0), ("steady", 150)):
        lr.train_rounds(warm)
        torch.cuda.synchronize()
        env.stats(reset=True)
        ts, tf = [], []
        for r in range(20):
            for c in range(2):
                a, b, d = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                a.record(); lr.substep(c); b.record(); lr.finalize(); d.record()
                ts.append((a, b)); tf.append((b, d))
            lr.round += 1
        torch.cuda.synchronize()
        st = env.stats(reset=True)
        s_ms = sum(x.elapsed_time(y) for x, y in ts) / len(ts)
        f_ms = sum(x.elapsed_time(y) for x, y in tf) / len(tf)
        out.append("%s: substep %.1f us finalize %.1f us units %.0f -> %.2e steps/s" %
                   (phase, s_ms * 1e3, f_ms * 1e3, st["agent_steps"] / len(ts), st["agent_steps"] / ((s_ms + f_ms) * len(ts) * 1e-3)))
    print("replicas", reps, " | ".join(out), flush=True)
    del lr, env
