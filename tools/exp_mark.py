"""Timing experiment (GPU): env-kernel / pull-kernel duration, dense (1) vs list (0) mode, by training phase."""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner

bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))["down_right"]
boxes, idx, choose = m.layout_3[2]["initial_states"][0]
sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
n = 1 << 20
modes = [x for x in (sys.argv[1] if len(sys.argv) > 1 else "1").split(",")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode in modes:
    os.environ["MDPP_DENSE"] = mode
    env = BatchedEnv(sc, n, seed=3)
    lr = BatchedQLearner(env, episodes=5000, seed=3)
    out = []
    for phase, warm in (("start", 0), ("r20", 20), ("r200", 180), ("r1000", 800)):
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
        # whole rounds through the chained entry point, L2 warm
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); lr.train_rounds(50); b.record()
        torch.cuda.synchronize()
        st2 = env.stats(reset=True)
        out.append("%s: env %.1f us pull+copy %.1f us | chained round %.1f us (%.2e steps/s, %d keys/substep)" %
                   (phase, s_ms * 1e3, f_ms * 1e3, a.elapsed_time(b) / 50 * 1e3,
                    st2["agent_steps"] / (a.elapsed_time(b) * 1e-3), st2["touched_keys"] / 100))
    print("dense", mode, "\n   " + "\n   ".join(out), flush=True)
    del lr, env
