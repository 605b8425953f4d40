"""Timing experiment (GPUs, under torchrun): the replica exchange alone, back to back, per protocol variant
(MDPP_SYNC_MODE), and the library all-reduce of the same table.  Device-timed, max over ranks."""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import marl_dpp_b200 as m
from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner
from marl_dpp_b200.parallel import PeerExchange

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
bad = json.load(gzip.open("tests/golden/bad_states.json.gz"))["down_right"]
boxes, idx, choose = m.layout_3[2]["initial_states"][0]
sc = m.Scenario(m.layout_3[2], boxes, idx, choose, bad_states_text=bad, step_cap=10000)
ex = PeerExchange(sc)
env = BatchedEnv(sc, 1 << 16, env_id_base=rank << 16, seed=3)
lr = BatchedQLearner(env, episodes=5000, seed=3, q=ex.q)
ex.attach(lr, 0)
lr.train_rounds(16)


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


for mode in ("0", "1"):
    os.environ["MDPP_SYNC_MODE"] = mode
    us = timed(lr.sync_replicas)
    if rank == 0:
        print("exchange kernel, %s: %.2f us per call (back to back, incl. launch)"
              % ("relaxed flags (default)" if mode == "0" else "fenced flags (MDPP_SYNC_MODE=1)", us))
os.environ["MDPP_SYNC_MODE"] = "0"
qc = lr.q.clone()
us = timed(lambda: dist.all_reduce(qc, op=dist.ReduceOp.AVG))
if rank == 0:
    print("ncclAllReduce(AVG) of the same %d KiB table: %.2f us per call" % (qc.numel() * 4 // 1024, us))
    print("internal errors:", env.stats()["internal_errors"])
a = torch.zeros(1, device="cuda")
us = timed(lambda: a.add_(1))
if rank == 0:
    print("empty torch kernel per call: %.2f us (host launch floor)" % us)
dist.barrier()
del lr, env
ex.close()
dist.destroy_process_group()
