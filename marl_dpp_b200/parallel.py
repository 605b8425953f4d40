"""Data-parallel plumbing: env shards + Q-replica averaging.

Envs are independent, so a batch shards by contiguous env-id ranges with no data-path
collective; the RNG is keyed by GLOBAL env id, so a shard computes exactly what the same envs
would compute inside one big batch.  The only exchange step is the periodic averaging of the
per-GPU Q replicas.  On one node that is ONE hand-written kernel reading every rank's table
through NVLink peer memory (csrc/q_sync.cuh, PeerExchange below: CUDA IPC is the plumbing, the
kernel is the product) -- a member of the learner's kernel chain, so a pure-exploration round runs
while the replicas are exchanged.  Tables too large for the one-shot exchange, and the CPU tests
(gloo), use a library all-reduce instead (average_).  The reference has no counterpart (single
process).  Averaging semantics: plain mean over ALL ranks per key -- a key only some ranks have
updated since the last exchange is pulled towards the others' value; at the batch sizes this is
built for every rank touches every reachable key between two exchanges.
"""
import ctypes as C

import torch
import torch.distributed as dist


def shard_env_ids(envs_per_rank, rank, world):
    """(env_id_base, n_envs) of `rank`: weak scaling, contiguous global ids."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return rank * envs_per_rank, envs_per_rank


def split_env_ids(total_envs, rank, world):
    """Strong-scaling split of `total_envs` ids: first (total % world) ranks get one extra."""
    base, extra = divmod(total_envs, world)
    n = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, n


def average_(tensor, group=None):
    """In-place mean over the ranks of `group`; a no-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()):
        return tensor
    world = dist.get_world_size(group)
    if world == 1:
        return tensor
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(tensor, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
        tensor.div_(world)
    return tensor


class ReplicaSync:
    """Calls average_ on the Q replica every `interval` rounds (0 / None = never)."""

    def __init__(self, interval, group=None):
        self.interval = int(interval or 0)
        self.group = group
        self.syncs = 0

    def due(self, rounds_done):
        return self.interval > 0 and rounds_done > 0 and rounds_done % self.interval == 0

    def step(self, q, rounds_done):
        if self.due(rounds_done):
            average_(q, self.group)
            self.syncs += 1
            return True
        return False


def merge_visited_(visited, group=None):
    """Bitwise OR of the per-rank `visited` words (which keys have ever been updated: what the qvalue
    file lists).  A key updated only on other ranks arrives here through the averaging with its bit
    clear; without this merge save_qvalues would drop it.  all-gather + OR (no backend has a BOR)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return visited
    parts = [torch.empty_like(visited) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, visited, group=group)
    for p in parts:
        visited |= p
    return visited


class _DevView:
    """A raw device pointer as something torch.as_tensor understands (no ownership)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerExchange:
    """This rank's exchange buffer (the Q table + the handshake flags of csrc/q_sync.cuh) and every
    peer's buffer mapped into this process.  `q` is the table as a torch tensor [n_states, 8] float32
    -- pass it to BatchedQLearner(q=...) and call attach(learner, interval).

    PeerExchange(scenario, group)    one process per GPU: buffers exchanged as CUDA IPC handles
                                     over the process group (all_gather_object);
    PeerExchange.local(scenario, n)  n buffers in THIS process on the current device (tests with fewer
                                     GPUs than ranks: emulate() runs the exchange of all n ranks as ONE
                                     cooperative launch; saturate_flags() turns a buffer into a peer that
                                     never has to be waited for)."""

    def __init__(self, scenario, group=None, device=None, _local=None):
        from . import _lib
        self._L = L = _lib.lib()
        self.scenario = scenario
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        nbytes = L.mdpp_sync_buffer_bytes(C.byref(scenario.cfg))
        if nbytes <= 0:
            raise _lib.MdppError("invalid scenario: " + _lib.last_error())
        if scenario.n_states * 32 > self.max_table_bytes():
            raise _lib.MdppError("table too large for the one-shot exchange: use a library all-reduce (average_)")
        self._own, self._opened = [], []
        with torch.cuda.device(self.device):
            if _local:
                self.rank, self.world = 0, int(_local)
                for _ in range(self.world):
                    p = C.c_void_p()
                    _lib.check(L.mdpp_sync_alloc(nbytes, C.byref(p), None))
                    self._own.append(p.value)
                self.ptrs = list(self._own)
            else:
                self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
                p, hd = C.c_void_p(), (C.c_uint8 * 64)()
                _lib.check(L.mdpp_sync_alloc(nbytes, C.byref(p), hd))
                self._own.append(p.value)
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(hd), group=group)
                self.ptrs = []
                for r in range(self.world):
                    if r == self.rank:
                        self.ptrs.append(p.value)
                        continue
                    q = C.c_void_p()
                    _lib.check(L.mdpp_sync_open((C.c_uint8 * 64).from_buffer_copy(handles[r]), C.byref(q)))
                    self._opened.append(q.value)
                    self.ptrs.append(q.value)
                dist.barrier(group=group)  # every rank has mapped every buffer before anyone uses or frees one
        self.tables = [self._table(p) for p in (self.ptrs if _local else [self.ptrs[self.rank]])]
        self.q = self.tables[0]

    @staticmethod
    def max_table_bytes():
        return 148 * 128 * 4 * 16   # MDPP_SYNC_MAX_TABLE_BYTES

    @classmethod
    def local(cls, scenario, world, device=None):
        return cls(scenario, device=device, _local=world)

    def _table(self, ptr):
        return torch.as_tensor(_DevView(ptr, (self.scenario.n_states, 8)), device=self.device)

    def attach(self, learner, interval, rank=None):
        """Bind a learner's handle: replicas are averaged before every round R > 0 with R % interval == 0
        inside train_rounds (0 = only when sync_replicas() is called)."""
        from . import _lib
        rank = self.rank if rank is None else int(rank)
        if learner.q.data_ptr() != self.ptrs[rank]:
            raise ValueError("the learner's q must be this exchange's table (BatchedQLearner(..., q=exchange.q))")
        arr = (C.c_void_p * self.world)(*self.ptrs)
        _lib.check(self._L.mdpp_sync_attach(learner.env._h, rank, self.world, arr, int(interval)))
        learner.exchange = self

    def flags(self, r):
        """The handshake flag words of local buffer r as an int32 tensor (tests)."""
        off = ((self.scenario.n_states * 32 + 255) // 256) * 256
        return torch.as_tensor(_DevView(self.ptrs[r] + off, (2 * 148 * 8,), "<i4"), device=self.device)

    def emulate(self, learners):
        """One exchange of ALL ranks of a local group as a single cooperative launch on the current
        stream (learners[r] attached as rank r)."""
        from . import _lib
        arr = (C.c_void_p * self.world)(*[lr.env._h.value for lr in learners])
        with torch.cuda.device(self.device):
            _lib.check(self._L.mdpp_sync_emulate(arr, self.world, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def close(self):
        for p in self._opened:
            self._L.mdpp_sync_close(C.c_void_p(p))
        for p in self._own:
            self._L.mdpp_sync_free(C.c_void_p(p))
        self._opened, self._own = [], []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- -dd reachability sharded over ranks (last sentence) ---------------------------------
def shard_words(n_words, rank, world):
    """Equal word ranges (the all-gather wants equal pieces): (first word, words per rank); the last ranks' pieces
    may reach past n_words -- the bitmap is padded to world * words_per_rank."""
    per = (n_words + world - 1) // world
    return rank * per, per


def sharded_fixpoint(bitmap, sweep_own, rank, world, group=None, max_sweeps=100000):
    """Monotone fixpoint over a REPLICATED bitmap whose word range [rank * per, (rank + 1) * per) this rank owns.
    sweep_own(bitmap) sets bits of own states only (reading any) and returns whether it set one.  After every
    sweep the ranks exchange their ranges (all-gather: a word belongs to exactly one rank, so no OR is needed) and
    the "anything changed anywhere" flag (all-reduce MAX).  Returns the number of sweeps.  bitmap: int32
    [world * per], identical on every rank on entry."""
    per = bitmap.numel() // world
    assert per * world == bitmap.numel()
    pieces = list(bitmap.view(world, per).unbind(0))
    sweeps = 0
    while True:
        changed = bool(sweep_own(bitmap))
        sweeps += 1
        if world > 1:
            own = pieces[rank].clone()
            dist.all_gather(pieces, own, group=group)
            flag = torch.tensor([1 if changed else 0], dtype=torch.int32, device=bitmap.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
            changed = bool(int(flag.item()))
        if not changed:
            return sweeps
        if sweeps >= max_sweeps:
            raise RuntimeError("sharded reachability did not converge")


def sharded_reachability(env, group=None):
    """The -dd bitmap of env's scenario with the joint-state range split over the ranks of `group` (one process per
    GPU): every rank sweeps its own word range (mdpp_deadlock_sweep_range), the ranks all-gather their ranges.
    Returns (bitmap int32 [ceil(n / 32)], sweeps); every rank ends with the whole bitmap, bit-identical to
    BatchedEnv.deadlock_reachability.  Without a process group: one rank, the plain sweeps."""
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    n = env.joint_state_count()
    words = (n + 31) // 32
    first_word, per = shard_words(words, rank, world)
    bitmap = torch.zeros(per * world, dtype=torch.int32, device=env.device)
    changed = torch.zeros(1, dtype=torch.int32, device=env.device)

    def sweep_own(bm):
        changed.zero_()
        env.deadlock_sweep_range(bm, first_word * 32, per * 32, changed)
        return int(changed.item())

    sweeps = sharded_fixpoint(bitmap, sweep_own, rank, world, group)
    return bitmap[:words], sweeps
