"""BASELINE config 4: Q-learning with per-GPU Q replicas, NCCL all-reduce(AVG) sync-interval sweep.
Run under torchrun (one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/sweep_sync.py [rounds=1024] [log2_envs_per_gpu=20]
Each point: `rounds` training rounds launched back to back on every rank (2 cars, layout 3 block 2, bad states on),
the replicas averaged every `interval` rounds; device time by CUDA events, max over ranks; also the spread between
the replicas before the last averaging (max |Q_rank - Q_mean| / max |Q_mean|)."""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
bad = json.load(gzip.open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bad_states.json.gz")))["down_right"]
boxes, idx, choose = m.layout_3[2]["initial_states"][0]
sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gmax(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gsum(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


if rank == 0:
    print("# %d GPU(s), %d envs per GPU, %d rounds per point" % (world, n, rounds))
# the all-reduce alone
env = BatchedEnv(sc, n, device=dev, env_id_base=rank * n, seed=3)
lr = BatchedQLearner(env, episodes=5000, seed=3)
for _ in range(5):
    lr.sync_replicas()
barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    lr.sync_replicas()
b.record()
barrier()
ar_us = gmax(a.elapsed_time(b) / 50 * 1e3)
if rank == 0:
    print("# all-reduce(AVG) of the %d KiB replica alone, back to back: %.1f us" % (lr.q.numel() * 4 // 1024, ar_us))
del lr, env
for interval in (1, 4, 16, 64, 256, 0):
    env = BatchedEnv(sc, n, device=dev, env_id_base=rank * n, seed=3)
    lr = BatchedQLearner(env, episodes=5000, seed=3)
    lr.train_rounds(64)
    if world > 1:
        lr.sync_replicas()
    barrier()
    env.stats(reset=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    done = 0
    chunk = interval if interval > 0 else rounds
    while done < rounds:
        c = min(chunk, rounds - done)
        lr.train_rounds(c)
        done += c
        if world > 1 and interval > 0 and done < rounds:
            lr.sync_replicas()
    b.record()
    barrier()
    ms = gmax(a.elapsed_time(b))
    steps = gsum(float(env.stats()["agent_steps"]))
    spread = 0.0
    if world > 1:
        mean = lr.q.clone()
        dist.all_reduce(mean, op=dist.ReduceOp.AVG)
        spread = gmax(float((lr.q - mean).abs().max() / mean.abs().max().clamp_min(1e-30)))
    if rank == 0:
        print("interval %4s  %.2f us/round  %.3e agent-env-steps/s (all GPUs)  replica spread before averaging %.2e"
              % (interval if interval else "inf", ms / rounds * 1e3, steps / (ms * 1e-3), spread))
    del lr, env
if world > 1:
    dist.destroy_process_group()
