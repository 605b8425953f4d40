"""Batched env + tabular learner: torch tensors in, the C ABI of include/mdpp.h underneath.

PyTorch is plumbing here: it owns device memory (one workspace blob per handle, the Q table,
the I/O tensors), the stream, and the NCCL process group used to average Q replicas.  Every
compute step is a hand-written sm_100a kernel in csrc/.  No CPU fallback: constructing a
BatchedEnv without a CUDA device raises.
"""
import ctypes as C

import torch

from . import _lib
from .tables import ACTIONS5, MdppStats, Scenario, learner_params

ACT_SKIP = 0xFF
ACT_GREEDY = 0xFE
REWARD_KINDS = {"q_reward": 0, "dqn_reward": 1}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class BatchedEnv:
    """n_envs independent copies of one (block, initial state) scenario, stepped in lockstep.

    Mirrors, per env, reference src/env_gym.py Environment: reset / initialize_cars / initialize
    (-> reset), encodestate + getlegalactions + step (-> step_car), is_game_ended (-> all_done).
    """

    def __init__(self, scenario, n_envs, device=None, env_id_base=0, seed=3):
        if not torch.cuda.is_available():
            raise RuntimeError("marl_dpp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        assert isinstance(scenario, Scenario)
        self.scenario = scenario
        self.n_envs = int(n_envs)
        self.n_cars = scenario.n_cars
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.seed = int(seed)
        self.env_id_base = int(env_id_base)
        L = _lib.lib()
        with torch.cuda.device(self.device):
            nbytes = L.mdpp_workspace_bytes(C.byref(scenario.cfg), self.n_envs)
            if nbytes < 0:
                raise _lib.MdppError("invalid scenario: " + _lib.last_error())
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            h = C.c_void_p()
            bad = scenario.bad_bitmap.ctypes.data_as(C.c_void_p) if scenario.bad_bitmap is not None else None
            _lib.check(L.mdpp_create(C.byref(scenario.cfg), self.n_envs, self.env_id_base, _ptr(self.workspace),
                                     nbytes, bad, _stream(), C.byref(h)))
        self._h = h
        self._L = L
        off = L.mdpp_env_records(h) - self.workspace.data_ptr()
        self.records = self.workspace[off:off + 16 * self.n_envs].view(torch.int32).view(self.n_envs, 4)
        self.reset()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.mdpp_destroy(h)

    # -- reset / initialize ---------------------------------------------------------------------
    def reset(self, headings=None, mask=None, zero_counters=True, seed=None):
        """headings: uint8 [n_envs, n_cars] in 0..3 = N,E,S,W (what the reference draws with
        random.shuffle, src/env_gym.py:345-349) or None for Philox; mask: uint8 [n_envs] or None."""
        if headings is not None:
            headings = headings.to(self.device, torch.uint8).contiguous()
            assert headings.shape == (self.n_envs, self.n_cars)
        if mask is not None:
            mask = mask.to(self.device, torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self._L.mdpp_reset(self._h, _ptr(mask), _ptr(headings),
                                          self.seed if seed is None else int(seed), int(zero_counters), _stream()))

    # -- one agent-step of one car slot ------------------------------------------------------------
    def step_car(self, car, actions, deadlock_id=1, reward="q_reward", outputs=True):
        """actions: uint8 [n_envs], 0..4 = S,L,R,F,T, 0xFF = leave this env alone."""
        actions = actions.to(self.device, torch.uint8).contiguous()
        n, dev = self.n_envs, self.device
        out = {}
        if outputs:
            out = {"state": torch.empty(n, dtype=torch.int32, device=dev),
                   "next_state": torch.empty(n, dtype=torch.int32, device=dev),
                   "reward": torch.empty(n, dtype=torch.float32, device=dev),
                   "legal": torch.empty(n, dtype=torch.uint8, device=dev),
                   "status": torch.empty(n, dtype=torch.uint8, device=dev),
                   "done": torch.empty(n, dtype=torch.uint8, device=dev)}
        with torch.cuda.device(dev):
            _lib.check(self._L.mdpp_step_car(self._h, car, deadlock_id, REWARD_KINDS[reward], _ptr(actions),
                                             _ptr(out.get("state")), _ptr(out.get("next_state")),
                                             _ptr(out.get("reward")), _ptr(out.get("legal")),
                                             _ptr(out.get("status")), _ptr(out.get("done")), _stream()))
        return out

    def step_car_host(self, car, actions, out, deadlock_id=1, reward="q_reward"):
        """End-to-end variant: `actions` and every tensor in `out` are (pinned) HOST tensors; the
        H2D copy, the kernel, the D2H copies and the synchronisation all happen inside the call."""
        with torch.cuda.device(self.device):
            _lib.check(self._L.mdpp_step_car_host(self._h, car, deadlock_id, REWARD_KINDS[reward], _ptr(actions),
                                                  _ptr(out.get("state")), _ptr(out.get("next_state")),
                                                  _ptr(out.get("reward")), _ptr(out.get("legal")),
                                                  _ptr(out.get("status")), _ptr(out.get("done")), _stream()))
        return out

    # -- string-API support ---------------------------------------------------------------------
    def query(self, car, state_idx, actions=None, reward="q_reward"):
        """Evaluate transitions from ARBITRARY state ids (entry i reads env i's cars for the clash
        term only): legal mask of the state, successor id, reward, legal mask of the successor."""
        dev = self.device
        state_idx = torch.as_tensor(state_idx, dtype=torch.int32, device=dev).contiguous()
        n = state_idx.numel()
        if actions is not None:
            actions = torch.as_tensor(actions, dtype=torch.uint8, device=dev).contiguous()
        out = {"legal": torch.empty(n, dtype=torch.uint8, device=dev),
               "next_state": torch.empty(n, dtype=torch.int32, device=dev),
               "reward": torch.empty(n, dtype=torch.float32, device=dev),
               "next_legal": torch.empty(n, dtype=torch.uint8, device=dev)}
        with torch.cuda.device(dev):
            _lib.check(self._L.mdpp_query_transitions(self._h, car, REWARD_KINDS[reward], _ptr(state_idx),
                                                      _ptr(actions), n, _ptr(out["legal"]), _ptr(out["next_state"]),
                                                      _ptr(out["reward"]), _ptr(out["next_legal"]), _stream()))
        return out

    def update_car(self, car, nodes):
        """Bare update_car (reference src/env_gym.py:386-441): uint8 [n_envs] node ids, 255 = skip."""
        nodes = torch.as_tensor(nodes, dtype=torch.uint8, device=self.device).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self._L.mdpp_update_car(self._h, car, _ptr(nodes), _stream()))

    def deadlock_reachability(self, method="auto"):
        """-dd precompute: bitmap (int32 words, bit j = joint state j can still reach "all cars
        finished") over the joint state space, and the number of sweeps / search levels it took.
        Joint index = sum_c v_c * V^c, v_c = node_local*2 + destination_index, V-1 = finished.
        method: "bfs" = backward breadth-first search in one cooperative launch (csrc/reach.cuh),
        "sweep" = forward sweeps to a fixpoint (the direct restatement of the definition),
        "auto" = bfs wherever joint indices fit 32 bits."""
        n = self._L.mdpp_joint_state_count(C.byref(self.scenario.cfg))
        if n <= 0:
            raise _lib.MdppError("joint state space too large: " + _lib.last_error())
        bitmap = torch.empty((n + 31) // 32, dtype=torch.int32, device=self.device)
        count = C.c_int32()
        nscratch = self._L.mdpp_deadlock_scratch_bytes(C.byref(self.scenario.cfg))
        if method == "auto":
            method = "bfs" if nscratch > 0 else "sweep"
        with torch.cuda.device(self.device):
            if method == "bfs":
                scratch = torch.empty(nscratch, dtype=torch.uint8, device=self.device)
                _lib.check(self._L.mdpp_deadlock_reachability_bfs(self._h, _ptr(bitmap), _ptr(scratch), nscratch,
                                                                  C.byref(count), _stream()))
            elif method == "sweep":
                _lib.check(self._L.mdpp_deadlock_reachability(self._h, _ptr(bitmap), C.byref(count), _stream()))
            else:
                raise ValueError("unknown method %r" % (method,))
        return bitmap, int(count.value)

    def joint_state_count(self):
        n = self._L.mdpp_joint_state_count(C.byref(self.scenario.cfg))
        if n <= 0:
            raise _lib.MdppError("joint state space too large: " + _lib.last_error())
        return int(n)

    def deadlock_sweep_range(self, bitmap, first_state, n_range, changed):
        """One forward sweep of the reachability fixpoint over joint states [first_state, first_state + n_range)
        of `bitmap` (int32 words on this device); `changed` (int32[1]) is set to 1 when a bit was set."""
        with torch.cuda.device(self.device):
            _lib.check(self._L.mdpp_deadlock_sweep_range(self._h, _ptr(bitmap), int(first_state), int(n_range),
                                                         _ptr(changed), _stream()))

    def joint_index(self, cars):
        """cars: [(node name or None for finished, destination_index)] -> joint state index."""
        from .tables import node_id
        sc = self.scenario
        V = 2 * sc.n_local_nodes + 1
        idx = 0
        for c, (node, di) in enumerate(cars):
            v = V - 1 if node is None else (node_id(node) - sc.node_base) * 2 + di
            idx += v * V ** c
        return idx

    def deadlock_states(self):
        """int32 [n_envs]: state id flagged by the stall heuristic, -1 = none."""
        off = self._L.mdpp_deadlock_states(self._h) - self.workspace.data_ptr()
        return self.workspace[off:off + 4 * self.n_envs].view(torch.int32)

    # -- DQN feed ---------------------------------------------------------------------------------
    def dqn_observe(self, car, features=None, legal=None, state=None):
        """encodestate(car, 1) for every env -> (features uint8 [n_envs, 80] = RL.convert(state),
        legal uint8 [n_envs] = give_mask as a bit mask, state id int32 [n_envs]).  Preallocated
        outputs (e.g. slices of a replay ring) may be passed in."""
        n, dev = self.n_envs, self.device
        features = torch.empty((n, 80), dtype=torch.uint8, device=dev) if features is None else features
        legal = torch.empty(n, dtype=torch.uint8, device=dev) if legal is None else legal
        state = torch.empty(n, dtype=torch.int32, device=dev) if state is None else state
        with torch.cuda.device(dev):
            _lib.check(self._L.mdpp_dqn_observe(self._h, car, _ptr(features), _ptr(legal), _ptr(state), _stream()))
        return features, legal, state

    def dqn_act(self, car, actions, reward="dqn_reward", next_features=None, out=None, auto_reset=False, round=0):
        """env.step(state, action, car) on the observed state -> dict(next_features, reward, next_legal,
        done, status)."""
        n, dev = self.n_envs, self.device
        actions = actions.to(dev, torch.uint8).contiguous()
        o = out or {}
        o.setdefault("next_features", torch.empty((n, 80), dtype=torch.uint8, device=dev)
                     if next_features is None else next_features)
        o.setdefault("reward", torch.empty(n, dtype=torch.float32, device=dev))
        o.setdefault("next_legal", torch.empty(n, dtype=torch.uint8, device=dev))
        o.setdefault("done", torch.empty(n, dtype=torch.uint8, device=dev))
        o.setdefault("status", torch.empty(n, dtype=torch.uint8, device=dev))
        with torch.cuda.device(dev):
            _lib.check(self._L.mdpp_dqn_act(self._h, car, REWARD_KINDS[reward], _ptr(actions), _ptr(o["next_features"]),
                                            _ptr(o["reward"]), _ptr(o["next_legal"]), _ptr(o["done"]),
                                            _ptr(o["status"]), int(auto_reset), int(round), self.seed, _stream()))
        return o

    # -- random-policy rollout -------------------------------------------------------------------
    def rollout_random(self, rounds, first_round=0, reward="q_reward", outputs=None):
        """`rounds` round-robin rounds under the uniform-random legal policy with auto-reset.
        outputs: None or dict of preallocated [rounds, n_cars, n_envs] tensors
        (action u8, reward f32, done u8, next_state i32)."""
        o = outputs or {}
        with torch.cuda.device(self.device):
            _lib.check(self._L.mdpp_rollout_random(self._h, int(rounds), int(first_round), self.seed,
                                                   REWARD_KINDS[reward], _ptr(o.get("action")), _ptr(o.get("reward")),
                                                   _ptr(o.get("done")), _ptr(o.get("next_state")), _stream()))
        return outputs

    def alloc_rollout_outputs(self, rounds):
        shape, dev = (rounds, self.n_cars, self.n_envs), self.device
        return {"action": torch.empty(shape, dtype=torch.uint8, device=dev),
                "reward": torch.empty(shape, dtype=torch.float32, device=dev),
                "done": torch.empty(shape, dtype=torch.uint8, device=dev),
                "next_state": torch.empty(shape, dtype=torch.int32, device=dev)}

    # -- inspection ----------------------------------------------------------------------------
    def launch_count(self):
        """Learner kernels launched through this handle so far."""
        return int(self._L.mdpp_launch_count(self._h))

    def stats(self, reset=False):
        s = MdppStats()
        with torch.cuda.device(self.device):
            _lib.check(self._L.mdpp_get_stats_host(self._h, C.byref(s), int(reset), _stream()))
        return s.as_dict()

    def cars(self):
        """Decoded car records: int64 tensors [n_envs, n_cars] node, dest_index, done, ghost."""
        r = self.records.to(torch.int64) & 0xFFFFFFFF
        packed = r[:, 0] | (r[:, 1] << 32)
        f = torch.stack([(packed >> (12 * i)) & 0xFFF for i in range(self.n_cars)], dim=1)
        return {"node": f & 127, "dest_index": (f >> 7) & 1, "done": (f >> 8) & 1, "ghost": (f >> 9) & 7,
                "episode": r[:, 2] & 0xFFFF, "steps": r[:, 2] >> 16}


class BatchedQLearner:
    """Tabular Q-learning over a BatchedEnv (reference src/algorithms/qlearning.py RL, batched).

    Semantics of one sub-step (car slot c of every env): all envs read the Q table as it was when
    the sub-step began, pick epsilon-greedily (per-env episode counter drives the reference's
    epsilon schedule) and step; each touched key then receives the reference update, which
    (state, action) and the snapshot determine completely, so it is computed once per key.  With
    one env this IS the reference's sequential update, bit for bit in float32.
    """

    def __init__(self, env, episodes=5000, alpha=0.9, discount=0.15, probability_model=3, seed=None, q=None,
                 flags=0):
        self.env = env
        self.params = learner_params(episodes, alpha, discount, probability_model,
                                     env.seed if seed is None else seed, flags)
        n_states = env.scenario.n_states
        self._q = q if q is not None else torch.zeros((n_states, 8), dtype=torch.float32, device=env.device)
        assert self._q.shape == (n_states, 8) and self._q.dtype == torch.float32 and self._q.is_contiguous()
        self.round = 0
        L, h = env._L, env._h
        off = L.mdpp_visited_bits(h) - env.workspace.data_ptr()
        self.visited = env.workspace[off:off + 4 * n_states].view(torch.int32)

    @property
    def q(self):
        """The Q table [n_states, 8] (slots 0..4 = S, L, R, F, T), current after every learner call.  Whoever takes
        the tensor may write it (load a file, average replicas): large tables keep a mirror copy in the workspace,
        so the library is told (mdpp_q_table_written) and copies the table once at its next learner call."""
        self.env._L.mdpp_q_table_written(self.env._h)
        return self._q

    @q.setter
    def q(self, t):
        assert t.shape == self._q.shape and t.dtype == torch.float32 and t.is_contiguous()
        self._q = t
        self.env._L.mdpp_q_table_written(self.env._h)

    def substep(self, car, forced=None, outputs=False):
        env = self.env
        n, dev = env.n_envs, env.device
        if forced is not None:
            forced = forced.to(dev, torch.uint8).contiguous()
        out = {}
        if outputs:
            out = {"action": torch.empty(n, dtype=torch.uint8, device=dev),
                   "state": torch.empty(n, dtype=torch.int32, device=dev),
                   "reward": torch.empty(n, dtype=torch.float32, device=dev)}
        with torch.cuda.device(dev):
            _lib.check(env._L.mdpp_q_substep(env._h, _ptr(self._q), C.byref(self.params), car, self.round,
                                             _ptr(forced), _ptr(out.get("action")), _ptr(out.get("state")),
                                             _ptr(out.get("reward")), _stream()))
        return out

    def finalize(self):
        with torch.cuda.device(self.env.device):
            _lib.check(self.env._L.mdpp_q_finalize(self.env._h, _ptr(self._q), C.byref(self.params), _stream()))

    def train_rounds(self, rounds):
        """`rounds` full round-robin rounds (n_cars x (sub-step + finalize)); asynchronous.
        On 1- and 2-car tables the library runs up to 8 rounds per launch while every env is certainly in an
        epsilon = 1 segment; it knows that from a host-side bound on the episode counters that grows by one per
        round and is tightened whenever `env.stats()` is fetched -- fetch the counters every few hundred rounds."""
        env = self.env
        with torch.cuda.device(env.device):
            _lib.check(env._L.mdpp_q_train_rounds(env._h, _ptr(self._q), C.byref(self.params), self.round,
                                                  int(rounds), _stream()))
        self.round += int(rounds)

    def train_rounds_host(self, rounds):
        """End-to-end call: learner parameters from host memory in, stats back, synchronised."""
        env = self.env
        s = MdppStats()
        if torch.cuda.current_device() == env.device.index:  # the common case: no device switch around the call
            _lib.check(env._L.mdpp_q_train_rounds_host(env._h, _ptr(self._q), C.byref(self.params), self.round,
                                                       int(rounds), C.byref(s), _stream()))
        else:
            with torch.cuda.device(env.device):
                _lib.check(env._L.mdpp_q_train_rounds_host(env._h, _ptr(self._q), C.byref(self.params), self.round,
                                                           int(rounds), C.byref(s), _stream()))
        self.round += int(rounds)
        return s.as_dict()

    def profile_env_kernel(self, rounds=1):
        """Profiling hook: only the env-side kernel of `rounds` fused rounds (1: the round kernel; 2..8: the
        multi-round pure-exploration kernel); its marks are never pulled."""
        env = self.env
        with torch.cuda.device(env.device):
            _lib.check(env._L.mdpp_q_profile_env_rounds(env._h, _ptr(self._q), C.byref(self.params), self.round,
                                                        int(rounds), _stream()))
        self.round += int(rounds)

    def sync_replicas(self, group=None):
        """Average the per-GPU Q replicas now.  With a PeerExchange attached (parallel.PeerExchange.attach):
        the library's own one-kernel exchange over NVLink peer memory; otherwise one library all-reduce
        (NCCL AVG, or SUM + scale on gloo)."""
        if getattr(self, "exchange", None) is not None:
            env = self.env
            with torch.cuda.device(env.device):
                _lib.check(env._L.mdpp_q_sync_replicas(env._h, _ptr(self._q), _stream()))
            return
        from . import parallel
        parallel.average_(self.q, group)

    def sync_count(self):
        """Replica averagings launched by the library through this handle (explicit + inside train_rounds)."""
        return int(self.env._L.mdpp_sync_count(self.env._h))

    def merge_visited(self, group=None):
        """OR the ranks' updated-key bits (call before qvalue_items / save on a multi-GPU run)."""
        from . import parallel
        parallel.merge_visited_(self.visited, group)

    # -- qvalue .txt format (reference src/algorithms/qlearning.py:159-185) --------------------------
    def qvalue_items(self):
        """[(state string, action char, float)] for every key updated so far, in state-id order."""
        sc = self.env.scenario
        vis = self.visited.cpu()
        q = self.q.cpu()
        out = []
        for s in torch.nonzero(vis).flatten().tolist():
            name = sc.decode(s)
            for a in range(5):
                if (int(vis[s]) >> a) & 1:
                    out.append((name, ACTIONS5[a], float(q[s, a])))
        return out
