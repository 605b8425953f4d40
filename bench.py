#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched MARL-DPP hot path on B200.

Workload (BASELINE.json configs[1]): tabular Q-learning on layout 3 block 2 (down_right), the
layout's live initial state (2 cars), bad-states ON, 2^20 lockstep envs per GPU, reference
hyper-parameters (alpha 0.9, gamma 0.15, epsilon model 3, 5000 episodes per env, step cap
10000).  One "step" = one call of the library's training entry point at the granularity it launches
at while every env explores: 8 round-robin rounds over every env of the batch = ONE env-kernel launch
(q_rounds_pure_kernel: every car of every env, 8 rounds: select, step, mark the touched keys) + the
pulls of its 16 sub-steps in two more launches (bitmap reduce + cluster pull: the reference TD update,
once per touched key).  The former definition -- one round per step, round kernel + two dense pulls --
is measured the same way and reported as aux.one_round_per_step.  Metric: agent-env-steps/s
(== Q-updates/s on this path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (contract in the task statement): value = device-timed whole-job
throughput, e2e = the same metric through the public host-facing call (parameters from host
memory in, counters back, synchronised every step), roofline = the dominant kernel against the
measured HBM peak, cpu_baseline = the CPU oracle port on this box's host cores.

Before anything is timed the batch is rolled forward (untimed) to its stationary mix of active and
finished cars: a round costs the same whether or not a car still moves, so a window that starts at a
fresh reset (99 % of the car slots active) over-reports the long-run rate (72 % active).  With N > 1
the replicas are averaged INSIDE the timed region (at least four times) by the library's own
peer-memory kernel, and the sync-interval sweep of BASELINE config 4 is reported in aux.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "agent-env-steps/sec"
UNIT = "agent-env-steps/s"
BLOCK = 2
STEP_CAP = 10000
EPISODES = 5000
WORKLOAD = "qlearning-layout3-block2-2car-badstates-1Mi-envs-per-gpu"
FUSED_Q_BYTES_PER_STEP = lambda k: 44.0 + 16.0 / k  # fused Q-learning, per agent-step
PREROLL_ROUNDS = 1024        # untimed rounds before any timed region: the active-car fraction is stationary by then
# The UNMODIFIED Python reference cannot run on the GPU box (/root/reference is absent there).  Its rate was
# measured in the build container (SURVEY.md section 6 [probe]): src/algorithms/qlearning.py train_given_state,
# layout 3 block 2, 2 cars, Display stubbed, stdout discarded, Python 3.12, ONE core (the reference is single-threaded).
PYTHON_REFERENCE = {"value": 1.28e4, "unit": UNIT, "cores": 1, "kind": "labelled constant, not measured in this run",
                    "provenance": "SURVEY.md section 6 [probe]: 84 532 agent-steps of the unmodified reference's "
                                  "RL.train_given_state in 6.6 s (300 episodes, block 2, 2 cars, cv2 Display stubbed) "
                                  "in the build container; 7.2e2 /s as shipped (Display rebuilt at every reset)"}


def workload_config():
    """The workload both arms run (BASELINE.json configs[1]); arm-specific keys are added by each arm."""
    return {"workload": WORKLOAD, "layout": 3, "block": BLOCK, "n_cars": 2, "envs_per_gpu": 1 << 20,
            "bad_states": True, "episodes_per_env": EPISODES, "step_cap": STEP_CAP,
            "alpha": 0.9, "discount": 0.15, "epsilon_model": 3}


def bad_text():
    import gzip
    with gzip.open(os.path.join(ROOT, "tests", "golden", "bad_states.json.gz")) as f:
        return json.load(f)["down_right"]


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
CPU_ENVS_PER_THREAD = 2048   # the same sample in both CPU legs (cpu_baseline of our arm, --impl reference)
CPU_PREROLL_ROUNDS = 512     # the CPU envs are rolled to their stationary mix first as well (untimed)


def _cpu_bench(threads):
    from oracle import oracle as orc
    import marl_dpp_b200 as m
    boxes, idx, choose = m.layout_3[BLOCK]["initial_states"][0]
    lp = orc.learner(EPISODES, seed=3)
    b = orc.CpuBench(orc.Config(m.layout_3[BLOCK]), threads, CPU_ENVS_PER_THREAD, boxes, idx, choose, bad_text(),
                     STEP_CAP, lp)
    b.run(CPU_PREROLL_ROUNDS)
    return b


def cpu_reference(seconds, threads=None, rounds_per_chunk=100):
    """The CPU arm: oracle port (the reference is pure Python and cannot travel to the GPU box; its
    measured rate is carried as a labelled constant).  One independent learner replica per host thread."""
    threads = threads or os.cpu_count() or 1
    b = _cpu_bench(threads)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        steps += b.run(rounds_per_chunk)
    dt = time.perf_counter() - t0
    b.close()
    return {"value": steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d independent oracle learners x %d envs, %d untimed pre-roll rounds, then %.1f s timed, same "
                      "scenario/hyper-parameters" % (threads, CPU_ENVS_PER_THREAD, CPU_PREROLL_ROUNDS, dt),
            "python_reference": PYTHON_REFERENCE}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    b = _cpu_bench(threads)
    # a step = enough rounds for ~0.5 s of CPU work (calibrated once, untimed): the per-step thread fan-out of the
    # oracle driver is then noise, and the whole --steps K --warmup W run ends within a minute or two
    t0 = time.perf_counter()
    b.run(20)
    per_round = (time.perf_counter() - t0) / 20
    rounds_per_step = max(20, min(2000, int(0.5 / max(per_round, 1e-6))))
    for _ in range(args.warmup):
        b.run(rounds_per_step)
    steps, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        steps += b.run(rounds_per_step)
    dt = time.perf_counter() - t0
    b.close()
    v = steps / dt
    sample = ("each step = %d rounds of %d independent oracle learners x %d envs after %d untimed pre-roll rounds "
              "(bounded sample of the workload; same sample as cpu_baseline of the GPU arm)"
              % (rounds_per_step, threads, CPU_ENVS_PER_THREAD, CPU_PREROLL_ROUNDS))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(), arm="CPU oracle port on host cores (the reference is pure Python): every "
                                              "step is a bounded sample of the workload, see cpu_baseline.sample"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "python_reference": PYTHON_REFERENCE},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--sync-interval", type=int, default=-1,
                    help="rounds between Q-replica averagings (N>1); default min(64, max(1, steps // 4)): at least four "
                         "inside the timed region; 0 = never")
    ap.add_argument("--preroll", type=int, default=PREROLL_ROUNDS, help="untimed rounds before the timed region")
    ap.add_argument("--rounds-per-step", type=int, default=8, help="rounds per timed step (one train_rounds call); 1..8")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-aux", action="store_true", help="skip the secondary (random-rollout, gym e2e) lines")
    ap.add_argument("--full-training", type=int, default=1,
                    help="aux: also run the whole 5000-episode training job once (~20 s on one B200); 0 skips it")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import ctypes as C
    import torch
    import torch.distributed as dist
    import marl_dpp_b200 as m
    from marl_dpp_b200 import _lib
    from marl_dpp_b200.batched import BatchedEnv, BatchedQLearner

    cpu = None
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        cpu = cpu_reference(args.cpu_seconds)  # before CUDA start-up: host cores are otherwise idle

    if _lib.needs_build():
        if local == 0:
            _lib.build()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    boxes, idx, choose = m.layout_3[BLOCK]["initial_states"][0]
    sc = m.Scenario(m.layout_3[BLOCK], boxes, idx, choose, bad_states_text=bad_text(), step_cap=STEP_CAP)
    n, k = args.envs, sc.n_cars
    env = BatchedEnv(sc, n, device=dev, env_id_base=rank * n, seed=3)
    # N > 1: the table lives in a peer-mapped exchange buffer and the library averages the replicas with its own
    # kernel inside train_rounds (csrc/q_sync.cuh); a table too large for that would use NCCL (not this workload)
    RPS = max(1, min(8, args.rounds_per_step))
    # default: at least four exchanges inside the timed region, at most every 64 rounds, on step boundaries
    interval = args.sync_interval if args.sync_interval >= 0 else min(64, max(RPS, (args.steps * RPS // 4) // RPS * RPS))
    exchange, sync_kind = None, None
    if world > 1:
        from marl_dpp_b200.parallel import PeerExchange
        try:
            exchange = PeerExchange(sc, device=dev)
            sync_kind = "q_sync_kernel: one-shot peer-memory exchange over NVLink (CUDA IPC), fused 1/N"
        except Exception as e:  # noqa: BLE001 -- reported in the line, never silent
            exchange, sync_kind = None, "ncclAllReduce(AVG) fallback: %s" % (str(e)[:120],)
    lr = BatchedQLearner(env, episodes=EPISODES, seed=3, q=exchange.q if exchange else None)
    if exchange:
        exchange.attach(lr, interval)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nccl_syncs = [0]

    def after_step():
        """NCCL fallback only: the library's own exchange happens inside train_rounds."""
        if world > 1 and not exchange and interval > 0 and lr.round % interval == 0:
            lr.sync_replicas()
            nccl_syncs[0] += 1

    # ---- untimed pre-roll to the stationary mix of active / finished cars --------------------------------
    done = 0
    while done < args.preroll:
        c = min(256, args.preroll - done)
        if world > 1 and not exchange and interval > 0:
            for _ in range(c):
                lr.train_rounds(1)
                after_step()
        else:
            lr.train_rounds(c)
        done += c
        env.stats()   # lets the library tighten its bound on the episode counters
    for i in range(args.warmup):
        lr.train_rounds(RPS)
        after_step()
    barrier()
    env.stats(reset=True)

    # ---- timed region: K steps, per-step CUDA events on the launching stream, L2 flushed between
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    wall0 = time.perf_counter()
    launches0 = env.launch_count()
    syncs0 = lr.sync_count() + nccl_syncs[0]
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        lr.train_rounds(RPS)
        after_step()
        ev[i][1].record()
        if (i + 1) % 32 == 0:
            # outside the event pair: a counter fetch (synchronises) lets the library tighten its bound on the
            # envs' episode counters, i.e. keep skipping the catch-up kernel while every env explores
            st_mid = env.stats()
    gpu_launches = env.launch_count() - launches0
    syncs_timed = lr.sync_count() + nccl_syncs[0] - syncs0
    barrier()
    wall = time.perf_counter() - wall0
    clk = clocks.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    st = env.stats(reset=True)
    t = torch.tensor([dev_ms, float(st["agent_steps"]), float(st["q_updates"])], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dev_ms = float(tmax[0])
    total_steps, total_upd = float(t[1]), float(t[2])
    value = total_steps / (dev_ms * 1e-3)
    active_fraction = total_steps / (float(args.steps) * RPS * n * k * world)

    # the exchange alone, in this run: back-to-back explicit averagings, device-timed, max over ranks
    sync_us, nccl_us = None, None
    if world > 1:
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            lr.sync_replicas()
        b.record()
        torch.cuda.synchronize()
        ts = torch.tensor([a.elapsed_time(b) / 50 * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sync_us = float(ts[0])
        # the library all-reduce of the same table, for comparison (what round 1 used: ncclAllReduce AVG)
        qc = lr.q.clone()
        for _ in range(5):
            dist.all_reduce(qc, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize()
        a.record()
        for _ in range(50):
            dist.all_reduce(qc, op=dist.ReduceOp.AVG)
        b.record()
        torch.cuda.synchronize()
        ts = torch.tensor([a.elapsed_time(b) / 50 * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        nccl_us = float(ts[0])
        del qc

    # ---- e2e: the public host-facing call, one round per call, counters read back every step ----
    barrier()
    env.stats(reset=True)
    e2e_dt = 0.0
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()           # the L2 flush is outside the timed window
        t0 = time.perf_counter()
        lr.train_rounds_host(RPS)  # learner params from host memory in, the counters back in host memory
        if world > 1 and not exchange and interval > 0 and lr.round % interval == 0:
            lr.sync_replicas()
            torch.cuda.synchronize()
        e2e_dt += time.perf_counter() - t0
    barrier()
    st2 = env.stats(reset=True)
    t2 = torch.tensor([e2e_dt, float(st2["agent_steps"])], dtype=torch.float64, device=dev)
    if world > 1:
        t2max = t2.clone()
        dist.all_reduce(t2max, op=dist.ReduceOp.MAX)
        dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        e2e_dt = float(t2max[0])
    e2e_value = float(t2[1]) / e2e_dt

    # ---- roofline: the kernels of one round against the HBM peak ------------------------------------
    # achieved = algorithmic bytes of the round (44 + 16/k per agent-step) / device time of the
    # round's kernels, i.e. of the timed steps above.  The dominant kernel (the round kernel) is also timed
    # alone, live, with CUDA events and a flushed L2 through the profiling hook.
    roof = None
    if rank == 0:
        peak, how = measured_peak()
        units = total_steps / world / args.steps
        step_ms = dev_ms / args.steps
        achieved = FUSED_Q_BYTES_PER_STEP(k) * units / (step_ms * 1e-3) / 1e9
        def env_kernel_alone(rounds, reps):
            """The env kernel of `rounds` rounds alone (profiling hook: its marks are never pulled), on a batch
            rolled to the same stationary mix, CUDA events, L2 flushed before every launch -> ms per launch."""
            envp = BatchedEnv(sc, n, device=dev, env_id_base=rank * n, seed=3)
            lrp = BatchedQLearner(envp, episodes=EPISODES, seed=3)
            for _ in range(0, args.preroll, 256):
                lrp.train_rounds(min(256, args.preroll))
                envp.stats()
            kev = []
            for i in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                lrp.profile_env_kernel(rounds)
                b.record()
                kev.append((a, b))
                if (i + 1) % 32 == 0:
                    envp.stats()
            torch.cuda.synchronize()
            return sum(a.elapsed_time(b) for a, b in kev) / len(kev)
        k0_ms = env_kernel_alone(RPS, min(args.steps, 200))
        # what an event pair measures with NOTHING between the records, after the same flush: every per-step number
        # above contains this much launch / event latency (chained rounds hide it, see aux)
        eev = []
        for i in range(100):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            b.record()
            eev.append((a, b))
        torch.cuda.synchronize()
        empty_ms = sum(a.elapsed_time(b) for a, b in eev) / len(eev)
        # DRAM bytes of the round's kernels from this round's committed `ncu --set full` capture of this command
        # (per launch, stationary mix); null when no capture of the current kernels is tracked
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic_r02.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic, traffic_src = tj.get("round_dram_bytes"), tj.get("source")
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": ("one step = %d rounds = q_rounds_pure_kernel<%d> + q_marks_reduce_kernel + q_multi_pull_kernel"
                           % (RPS, k)) if RPS > 1 else
                          "one round = q_round_kernel<%d,0> + %d x q_pull_dense_kernel" % (k, k),
                "avg_launch_ms": step_ms, "units_per_launch": units, "empty_event_pair_ms": empty_ms,
                "algorithmic_bytes_per_unit": FUSED_Q_BYTES_PER_STEP(k), "peak_source": how,
                "dominant_kernel": {"name": ("q_rounds_pure_kernel<%d>" if RPS > 1 else "q_round_kernel<%d,0>") % k,
                                    "avg_launch_ms": k0_ms,
                                    "share_of_step": k0_ms / step_ms,
                                    "how": "timed alone with CUDA events, L2 flushed before every launch"},
                "note": "frac counts the whole step (env kernel + pulls) against the algorithmic bytes of fused "
                        "Q-learning; the env kernel is bound by instruction issue / the integer ALU pipe (ncu), "
                        "records and table are L2-resident at this size, so DRAM traffic is below the algorithmic bytes"}

    # ---- secondary lines (not the headline): random-policy rollout and gym-style host stepping ----
    aux = {}
    if world > 1 and exchange and not args.no_aux:
        # BASELINE config 4: sync-interval sweep, every rank takes part (the exchange is collective).  256 rounds
        # per point, launched back to back in calls of 64 (so that interval 64 averages inside the library's
        # kernel chain), device-timed, max over ranks.
        sweep = {}
        for iv in (1, 4, 16, 64, 0):
            exchange.attach(lr, iv)
            lr.round = ((lr.round + 63) // 64) * 64      # the same phase of the interval at every point
            barrier()
            env.stats(reset=True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0 = lr.sync_count()
            a.record()
            for _ in range(4):
                lr.train_rounds(64)
            b.record()
            torch.cuda.synchronize()
            sv = env.stats(reset=True)
            tt = torch.tensor([a.elapsed_time(b), float(sv["agent_steps"])], dtype=torch.float64, device=dev)
            tm = tt.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            sweep["never" if iv == 0 else str(iv)] = {
                "value": float(tt[1]) / (float(tm[0]) * 1e-3), "us_per_round": float(tm[0]) / 256 * 1e3,
                "exchanges": lr.sync_count() - c0}
        exchange.attach(lr, interval)
        if rank == 0:
            aux["sync_interval_sweep"] = {"unit": UNIT, "rounds_per_point": 256, "points": sweep,
                                          "note": "rounds back to back (L2 warm), replicas averaged by q_sync_kernel "
                                                  "before every round R with R % interval == 0"}
    if rank == 0 and world == 1 and not args.no_aux and RPS > 1:
        # the former headline definition, measured the same way: ONE round per step = round kernel + two dense pulls
        env1 = BatchedEnv(sc, n, device=dev, seed=3)
        lr1 = BatchedQLearner(env1, episodes=EPISODES, seed=3)
        for _ in range(0, args.preroll, 256):
            lr1.train_rounds(min(256, args.preroll))
            env1.stats()
        env1.stats(reset=True)
        ev1 = []
        for i in range(min(args.steps * RPS, 400)):
            flush.zero_()
            a1, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a1.record()
            lr1.train_rounds(1)
            b1.record()
            ev1.append((a1, b1))
            if (i + 1) % 128 == 0:
                env1.stats()
        torch.cuda.synchronize()
        s1 = env1.stats(reset=True)
        ms1 = sum(x.elapsed_time(y) for x, y in ev1)
        k1_ms = env_kernel_alone(1, 200)
        aux["one_round_per_step"] = {
            "value": s1["agent_steps"] / (ms1 * 1e-3), "unit": UNIT, "ms_per_step": ms1 / len(ev1),
            "frac_of_hbm_roofline": FUSED_Q_BYTES_PER_STEP(k) * s1["agent_steps"] / (ms1 * 1e-3) / 1e9 / measured_peak()[0],
            "dominant_kernel": {"name": "q_round_kernel<%d,0>" % k, "avg_launch_ms": k1_ms,
                                "share_of_step": k1_ms / (ms1 / len(ev1))},
            "note": "round 1's headline definition: one round per timed step (q_round_kernel + two q_pull_dense_kernel), "
                    "L2 flushed before every step, per-step events, stationary mix"}
        del lr1, env1
    if rank == 0 and world == 1 and not args.no_aux:
        # the same learner, rounds launched back to back (no flush, no per-round events): training throughput
        enva = BatchedEnv(sc, n, device=dev, seed=3)
        lra = BatchedQLearner(enva, episodes=EPISODES, seed=3)
        for _ in range(0, args.preroll, 256):
            lra.train_rounds(min(256, args.preroll))
            enva.stats()
        torch.cuda.synchronize()
        enva.stats(reset=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lra.train_rounds(500)
        b.record()
        torch.cuda.synchronize()
        s = enva.stats(reset=True)
        aux["qlearning_rounds_back_to_back"] = {"value": s["agent_steps"] / (a.elapsed_time(b) * 1e-3), "unit": UNIT,
                                                "us_per_round": a.elapsed_time(b) / 500 * 1e3,
                                                "active_fraction": s["agent_steps"] / (500.0 * n * k),
                                                "note": "500 rounds in one call, L2 warm, epsilon = 1 segment"}
        # the library's own granularity while every env explores: 8 rounds per env-kernel launch, their pulls in two
        # more launches (q_round.cuh); L2 flushed before every call like the headline, per-call events
        evg = []
        for _ in range(60):
            flush.zero_()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            lra.train_rounds(8)
            eb.record()
            evg.append((ea, eb))
        torch.cuda.synchronize()
        s = enva.stats(reset=True)
        gms = sum(x.elapsed_time(y) for x, y in evg)
        aux["qlearning_8_rounds_per_call_l2_flushed"] = {
            "value": s["agent_steps"] / (gms * 1e-3), "unit": UNIT, "us_per_round": gms / (60 * 8) * 1e3,
            "frac_of_hbm_roofline": FUSED_Q_BYTES_PER_STEP(k) * s["agent_steps"] / (gms * 1e-3) / 1e9 / measured_peak()[0],
            "note": "train_rounds(8) per timed call, 256 MiB L2 flush before each call (outside the event pair)"}
        # ... and the same granularity end to end: parameters from host memory, counters back, synchronised
        t_e2e = 0.0
        for _ in range(60):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lra.train_rounds_host(8)
            t_e2e += time.perf_counter() - t0
        s = enva.stats(reset=True)
        aux["e2e_8_rounds_per_call_l2_flushed"] = {
            "value": s["agent_steps"] / t_e2e, "unit": UNIT, "us_per_round": t_e2e / (60 * 8) * 1e6,
            "call": "BatchedQLearner.train_rounds_host(8) (host wall clock, sync + counter read-back every call)"}
        # a later training phase: epsilon pinned to 0.5, so half of the decisions are greedy (Q-row gathers,
        # parked lanes finished by the catch-up kernel)
        for i in range(6):
            lra.params.eps_q16[i] = 32768
        lra.train_rounds(100)
        torch.cuda.synchronize()
        enva.stats(reset=True)
        a.record()
        lra.train_rounds(300)
        b.record()
        torch.cuda.synchronize()
        s = enva.stats(reset=True)
        aux["qlearning_rounds_epsilon_0.5"] = {"value": s["agent_steps"] / (a.elapsed_time(b) * 1e-3), "unit": UNIT,
                                               "us_per_round": a.elapsed_time(b) / 300 * 1e3,
                                               "greedy_fraction": 1.0 - s["explore_steps"] / max(s["agent_steps"], 1)}
        del lra, enva
        if args.full_training:
            # the whole BASELINE training job: 2^20 envs x 5000 episodes each, epsilon schedule 1 -> 0, through the
            # public call sequence of RL._run (256 rounds per call, a counter fetch between calls)
            envf = BatchedEnv(sc, n, device=dev, seed=3)
            lrf = BatchedQLearner(envf, episodes=EPISODES, seed=3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rounds_f = 0
            while True:
                lrf.train_rounds(256)
                rounds_f += 256
                sf = envf.stats()
                if int((envf.records[:, 2] & 0xFFFF).min()) >= EPISODES or rounds_f > 4_000_000:
                    break
            dt = time.perf_counter() - t0
            aux["full_training_run"] = {"wall_s": dt, "rounds": rounds_f, "agent_steps": sf["agent_steps"],
                                        "episodes": sf["episodes_done"], "value": sf["agent_steps"] / dt, "unit": UNIT,
                                        "note": "2^20 envs x 5000 episodes per env, every epsilon segment, host wall clock"}
            del lrf, envf
        # the learner on larger tables ("max agents per layout": 3, 4, 5 cars; 126 540 / 1.65 M / 16.5 M states): scan mode
        # -- global touched bitmaps + a mirror table, one scanning pull per sub-step, all cars of several pure-exploration
        # rounds in one env pass with the pulls beside the next pass.  Rolled to the stationary mix first.
        for kc, (bx, ix, ch) in (
                (3, (["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]])),
                (4, (["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]])),
                (5, (["D1", "D2", "D3", "D5", "D7"], [0, 1, 0, 1, 0], [[0, 1, 0], [2, 1, 0], [1, 0, 0], [2, 2, 0], [3, 4, 0]]))):
            sck = m.Scenario(m.layout_3[BLOCK], bx, ix, ch, bad_states_text=bad_text(), step_cap=STEP_CAP)
            envk = BatchedEnv(sck, n, device=dev, seed=3)
            lrk = BatchedQLearner(envk, episodes=EPISODES, seed=3)
            lrk.train_rounds(args.preroll // 2)
            entry = {"states": sck.n_states, "algorithmic_bytes_per_step": FUSED_Q_BYTES_PER_STEP(kc)}
            for tag, rounds_k in (("pure_exploration", 64), ("epsilon_0.5", 32)):
                if tag == "epsilon_0.5":
                    for i in range(6):
                        lrk.params.eps_q16[i] = 32768
                    lrk.train_rounds(8)
                torch.cuda.synchronize()
                envk.stats(reset=True)
                a.record()
                for _ in range(rounds_k // 32):   # 32 rounds per call (RL._run makes calls of 256): within a call the pulls of
                    lrk.train_rounds(32)          # a group of pure-exploration rounds run beside the next group's env pass
                b.record()
                torch.cuda.synchronize()
                sk = envk.stats(reset=True)
                ms = a.elapsed_time(b)
                v = sk["agent_steps"] / (ms * 1e-3)
                entry[tag] = {"value": v, "unit": UNIT, "us_per_round": ms / rounds_k * 1e3, "rounds_per_call": 32,
                              "active_fraction": sk["agent_steps"] / float(rounds_k * n * kc),
                              "touched_keys_per_substep": sk["touched_keys"] / float(rounds_k * kc),
                              "frac_of_hbm_roofline": FUSED_Q_BYTES_PER_STEP(kc) * v / 1e9 / measured_peak()[0]}
            assert envk.stats()["internal_errors"] == 0
            aux["qlearning_%dcar_1Mi_envs" % kc] = entry
            del lrk, envk
        # random-policy rollouts (BASELINE config 3).  Like the learner, every batch is first rolled (untimed) to its
        # stationary mix of active and finished cars -- a round costs the same whether or not a car still moves --
        # and the active fraction of the timed rounds is printed next to the rate.
        rounds = 16
        env2 = BatchedEnv(sc, n, device=dev, env_id_base=rank * n, seed=4)
        for w in range(args.preroll // rounds):
            env2.rollout_random(rounds, w * rounds)
        r0 = (args.preroll // rounds) * rounds
        torch.cuda.synchronize()
        env2.stats(reset=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        flush.zero_()
        a.record()
        for w in range(reps):
            env2.rollout_random(rounds, r0 + w * rounds)
        b.record()
        torch.cuda.synchronize()
        s = env2.stats(reset=True)
        aux["random_rollout_no_outputs"] = {"value": s["agent_steps"] / (a.elapsed_time(b) * 1e-3), "unit": UNIT,
                                            "rounds_per_launch": rounds,
                                            "active_fraction": s["agent_steps"] / float(reps * rounds * n * k)}
        r0 += reps * rounds
        outs = env2.alloc_rollout_outputs(1)
        a.record()
        for w in range(reps * 4):
            env2.rollout_random(1, r0 + w, outputs=outs)
        b.record()
        torch.cuda.synchronize()
        s = env2.stats(reset=True)
        ms = a.elapsed_time(b)
        aux["random_rollout_outputs_materialised"] = {
            "value": s["agent_steps"] / (ms * 1e-3), "unit": UNIT, "rounds_per_launch": 1,
            "active_fraction": s["agent_steps"] / float(reps * 4 * n * k),
            "algorithmic_GBps": (10.0 + 16.0 / k) * s["agent_steps"] / (ms * 1e-3) / 1e9}
        # gym-style e2e: host actions in, host (state, next, reward, legal, status, done) out, per car slot
        acts = torch.zeros(n, dtype=torch.uint8).pin_memory()
        out = {"state": torch.empty(n, dtype=torch.int32).pin_memory(),
               "next_state": torch.empty(n, dtype=torch.int32).pin_memory(),
               "reward": torch.empty(n, dtype=torch.float32).pin_memory(),
               "legal": torch.empty(n, dtype=torch.uint8).pin_memory(),
               "status": torch.empty(n, dtype=torch.uint8).pin_memory(),
               "done": torch.empty(n, dtype=torch.uint8).pin_memory()}
        env2.reset()
        env2.stats(reset=True)
        t0 = time.perf_counter()
        for w in range(10):
            for c in range(k):
                env2.step_car_host(c, acts, out)
        dt = time.perf_counter() - t0
        s = env2.stats(reset=True)
        aux["gym_step_host_buffers"] = {"value": s["agent_steps"] / dt, "unit": UNIT,
                                        "h2d_bytes_per_call": n, "d2h_bytes_per_call": 15 * n}
        del env2
        # config 3 point: random-policy sweep, max agents (4 cars), 2^24 envs, outputs materialised
        b4, i4 = ["D2", "D6", "D4", "D8"], [0, 0, 0, 0]
        c4 = [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]]
        sc4 = m.Scenario(m.layout_3[BLOCK], b4, i4, c4, bad_states_text=bad_text(), step_cap=STEP_CAP)
        n4 = 1 << 24
        env4 = BatchedEnv(sc4, n4, device=dev, seed=5)
        outs4 = env4.alloc_rollout_outputs(1)
        for w in range(args.preroll // 16):            # untimed: to the stationary mix (see above)
            env4.rollout_random(16, w * 16)
        r4 = (args.preroll // 16) * 16
        for w in range(3):
            env4.rollout_random(1, r4 + w, outputs=outs4)
        torch.cuda.synchronize()
        env4.stats(reset=True)
        a.record()
        for w in range(10):
            env4.rollout_random(1, r4 + 3 + w, outputs=outs4)
        b.record()
        torch.cuda.synchronize()
        s = env4.stats(reset=True)
        ms = a.elapsed_time(b)
        aux["random_rollout_4car_16Mi_envs_outputs"] = {
            "value": s["agent_steps"] / (ms * 1e-3), "unit": UNIT, "rounds_per_launch": 1,
            "active_fraction": s["agent_steps"] / float(10 * n4 * 4),
            "car_slots_per_s": 10 * n4 * 4 / (ms * 1e-3),
            "algorithmic_GBps": (10.0 + 16.0 / 4) * s["agent_steps"] / (ms * 1e-3) / 1e9,
            "slot_GBps": (10.0 + 16.0 / 4) * 10 * n4 * 4 / (ms * 1e-3) / 1e9,
            "note": "records (256 MiB) + outputs (640 MiB) exceed L2: HBM-streaming regime; every car slot of every env "
                    "writes its 10 output bytes whether or not the car still moves (slot_GBps counts those)"}
        del env4, outs4
        # -dd precompute of the headline scenario (2 cars) and of the 4-car scenario: reachability bitmap
        for tag, scn in (("2car", sc), ("4car", sc4)):
            envd = BatchedEnv(scn, 1024, device=dev, seed=7)
            joint = (2 * scn.n_local_nodes + 1) ** scn.n_cars
            entry = {"joint_states": joint}
            for method in ("bfs", "sweep"):
                envd.deadlock_reachability(method=method)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                bm, count = envd.deadlock_reachability(method=method)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                entry[method] = {"ms": dt * 1e3, "levels" if method == "bfs" else "sweeps": count}
            live = int(sum(int((bm[i:i + (1 << 22)].view(torch.uint8).to(torch.int16).unsqueeze(1)
                                >> torch.arange(8, device=dev, dtype=torch.int16) & 1).sum())
                           for i in range(0, bm.numel(), 1 << 22)))
            entry["live_states"] = live
            # roofline of the search: per live state one bit set in the result and one in a frontier, per examined
            # predecessor one result-bit probe (4-byte sector reads in L2); reported as states and probes per second
            entry["bfs_live_states_per_s"] = live / (entry["bfs"]["ms"] * 1e-3)
            nck = scn.n_cars
            entry["roofline"] = {
                "unit": "bit operations/s",
                "bits_set_per_s": 2.0 * live / (entry["bfs"]["ms"] * 1e-3),
                "probes_per_s_upper_bound": 5.0 * nck * live / (entry["bfs"]["ms"] * 1e-3),
                "note": "algorithmic work of the backward search (SURVEY 8d): per live state one bit set in the result and "
                        "one in a frontier, and at most 5 x n_cars candidate predecessors examined (one result-bit probe "
                        "each); the search is bound by its " + str(entry["bfs"].get("levels")) + " grid barriers and "
                        "scattered L2 probes, not by HBM (the 4-car bitmaps are 3.5 MB)"}
            entry["note"] = ("bfs = backward breadth-first search, one cooperative launch, host wall clock incl. the "
                             "scratch memset and the final 8-byte read-back; sweep = forward sweeps to the fixpoint")
            aux["dd_reachability_" + tag] = entry
            del envd, bm
        # config 5 point: DQN feed, 2^24 envs, (convert(s), action, convert(s'), reward, mask(s')) per agent-step
        n5 = 1 << 24
        env5 = BatchedEnv(sc, n5, device=dev, seed=6)
        f_s = torch.empty((n5, 80), dtype=torch.uint8, device=dev)
        f_n = torch.empty((n5, 80), dtype=torch.uint8, device=dev)
        lg = torch.empty(n5, dtype=torch.uint8, device=dev)
        sid = torch.empty(n5, dtype=torch.int32, device=dev)
        o5 = {"reward": torch.empty(n5, dtype=torch.float32, device=dev),
              "next_legal": torch.empty(n5, dtype=torch.uint8, device=dev),
              "done": torch.empty(n5, dtype=torch.uint8, device=dev),
              "status": torch.empty(n5, dtype=torch.uint8, device=dev)}
        kernel_ms, reps5 = 0.0, 6
        for w in range(2 + reps5):
            if w == 2:
                env5.stats(reset=True)
            for c in range(k):
                e0, e1, e2, e3 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                e0.record()
                env5.dqn_observe(c, features=f_s, legal=lg, state=sid)
                e1.record()
                acts = torch.where(lg != 0, torch.log2((lg & (0 - lg)).float()).to(torch.uint8),
                                   torch.full_like(lg, 0xFF))   # lowest legal action (stand-in policy)
                e2.record()
                env5.dqn_act(c, acts, next_features=f_n, out=o5, auto_reset=True, round=w)
                e3.record()
                torch.cuda.synchronize()
                if w >= 2:
                    kernel_ms += e0.elapsed_time(e1) + e2.elapsed_time(e3)
        s = env5.stats(reset=True)
        aux["dqn_feed_16Mi_envs"] = {
            "value": s["agent_steps"] / (kernel_ms * 1e-3), "unit": UNIT,
            "algorithmic_GBps": (166.0 + 16.0 / k) * s["agent_steps"] / (kernel_ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": (166.0 + 16.0 / k) * s["agent_steps"] / (kernel_ms * 1e-3) / 1e9 / measured_peak()[0],
            "note": "dqn_observe + dqn_act kernels only (the policy network is PyTorch and not timed)"}
        del env5, f_s, f_n

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(), n_cars=k, envs_per_gpu=n, **{
                       "rounds_per_step": RPS,
                       "step": ("one train_rounds(%d) call = %d round-robin rounds over every env = one q_rounds_pure_kernel "
                                "launch (all cars, all rounds) + q_marks_reduce_kernel + q_multi_pull_kernel (the %d sub-steps' "
                                "updates); aux.one_round_per_step holds the one-round-per-step figures" % (RPS, RPS, RPS * k))
                               if RPS > 1 else
                               "one round-robin round over every env = q_round_kernel (all cars) + one q_pull_dense_kernel per car slot",
                       "l2": "flushed between timed steps (256 MiB device write outside the event pair)",
                       "preroll_rounds": args.preroll, "active_fraction": active_fraction,
                       "sync_interval_rounds": interval if world > 1 else None,
                       "allreduces_in_timed_region": syncs_timed if world > 1 else None,
                       "allreduce_us": sync_us, "allreduce": sync_kind, "nccl_allreduce_avg_us_same_table": nccl_us,
                       "parallelism": "dp%d: env shards + per-GPU Q replicas averaged every sync_interval_rounds" % world}),
            "q_updates_per_s": total_upd / (dev_ms * 1e-3),
            "wall_s_timed_region": wall,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(lr.params),
                    "d2h_bytes_per_step": 80,
                    "call": "BatchedQLearner.train_rounds_host(1) -> mdpp_q_train_rounds_host: learner parameters from "
                            "host memory in, the counters of the step back in host memory (stored there by the "
                            "device, the host waits for them) before the call returns, every step"},
            "gpu_launches": gpu_launches * world,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clk, "aux": aux,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
