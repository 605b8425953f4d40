"""BASELINE config 3: random-policy env-step throughput, max agents (4 cars), outputs materialised, env-count sweep.
Single GPU: python tools/sweep_rollout.py.  N GPUs (independent env shards, no collective on the data path):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/sweep_rollout.py
Device time by CUDA events, max over ranks; agent-steps summed over ranks.  Every batch is first rolled PREROLL rounds
(argv[1], default 1024; 0 = from a fresh reset) to its stationary mix of moving and finished cars."""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
bad = json.load(gzip.open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bad_states.json.gz")))["down_right"]
sc = m.Scenario(m.layout_3[2], ["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]],
                bad_states_text=bad, step_cap=10000)


def allr(x, op):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=op)
    return float(t[0])


if rank == 0:
    print("# %d GPU(s); layout 3 block 2, 4 cars, bad states on; action u8 + reward f32 + done u8 + next-state u32 per agent-step" % world)
for lg in (20, 22, 24, 26):
    n = 1 << lg
    env = BatchedEnv(sc, n, device=dev, env_id_base=rank * n, seed=5)
    outs = env.alloc_rollout_outputs(1)
    pre = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    for w in range(pre // 16):
        env.rollout_random(16, w * 16)
    r0 = (pre // 16) * 16
    for w in range(3):
        env.rollout_random(1, r0 + w, outputs=outs)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    env.stats(reset=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 16 if lg <= 22 else 8
    a.record()
    for w in range(reps):
        env.rollout_random(1, r0 + 3 + w, outputs=outs)
    b.record()
    torch.cuda.synchronize()
    ms = allr(a.elapsed_time(b), dist.ReduceOp.MAX if world > 1 else None)
    steps = allr(float(env.stats()["agent_steps"]), dist.ReduceOp.SUM if world > 1 else None)
    v = steps / (ms * 1e-3)
    if rank == 0:
        slots = reps * 4.0 * n * world / (ms * 1e-3)
        print("2^%d envs per GPU (%d in all): %.3e agent-env-steps/s (%.0f %% of the car slots move; %.3e car slots/s), %.0f GB/s "
              "algorithmic (14 B/step) = %.1f %% of %d x 6545.9 GB/s" % (lg, n * world, v, 100 * v / slots, slots, 14 * v / 1e9,
              100 * 14 * v / 1e9 / (6545.9 * world), world), flush=True)
    del env, outs
if world > 1:
    dist.destroy_process_group()
