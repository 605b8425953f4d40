"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (oracle/mdpp_oracle.c).

The oracle is the checker, never the product.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module.
"""
import ctypes as C
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmdpp_oracle.so")

ORC_NONE = 255
ACTIONS = "SLRFBT"  # reference src/env_gym.py:6-21
HEADS = "NESW"      # reference src/layout.py:10


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("mdpp_oracle.c", "mdpp_oracle_batch.c", "mdpp_oracle.h", "Makefile")]
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    subprocess.check_call(["make", "-s", "-C", HERE, "-B", "libmdpp_oracle.so"])
    return LIB_PATH


# ---- node / box strings <-> ids --------------------------------------------------------------
def node_id(s):
    m = re.fullmatch(r"([UD])(\d+)([NESW])", s)
    return (m.group(1) == "D") * 36 + (int(m.group(2)) - 1) * 4 + HEADS.index(m.group(3))


def box_id(s):
    m = re.fullmatch(r"([UD])(\d+)", s)
    return (m.group(1) == "D") * 9 + int(m.group(2)) - 1


def node_str(n):
    if n == ORC_NONE:
        return "?YY"
    return "UD"[n // 36] + str((n % 36) // 4 + 1) + HEADS[n & 3]


class Layout(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32),
                ("n_no_box", C.c_int32), ("no_box", C.c_int32 * 18),
                ("n_disc", C.c_int32), ("disc_nodes", C.c_int32 * 72),
                ("n_station", C.c_int32), ("station_box", C.c_int32 * 8), ("station_entrance", C.c_int32 * 8),
                ("n_dest", C.c_int32), ("dest_nodes", (C.c_int32 * 8) * 2),
                ("tcdi", C.c_int32),
                ("n_sel", C.c_int32), ("sel_key", C.c_int32 * 4), ("sel_n", C.c_int32 * 4),
                ("sel_boxes", (C.c_int32 * 18) * 4)]


class State(C.Structure):
    _fields_ = [("pnode", C.c_uint8), ("dnode", C.c_uint8), ("n_others", C.c_uint8),
                ("others", (C.c_uint8 * 2) * 6)]

    def key(self):
        return (self.pnode, self.dnode, tuple((self.others[i][0], self.others[i][1]) for i in range(self.n_others)))


def layout_struct(d):
    """Layout dict (reference src/layout.py schema) -> orc_layout."""
    lay = Layout()
    lay.rows, lay.cols = d["number_of_rows"], d["number_of_cols"]
    lay.n_no_box = len(d["no_box_list"])
    for i, b in enumerate(d["no_box_list"]):
        lay.no_box[i] = box_id(b)
    lay.n_disc = len(d["disconnected_nodes_list"])
    for i, n in enumerate(d["disconnected_nodes_list"]):
        lay.disc_nodes[i] = node_id(n)
    lay.n_station = len(d["station_box_list"])
    for i, b in enumerate(d["station_box_list"]):
        lay.station_box[i] = box_id(b)
        lay.station_entrance[i] = sum(1 << HEADS.index(h) for h in d["station_entrance_list"][i])
    lists = d["station_destination_list"]
    assert len(lists) == 2
    lay.n_dest = len(lists[0])
    for li in range(2):
        for i, n in enumerate(lists[li]):
            lay.dest_nodes[li][i] = node_id(n)
    lay.tcdi = d["training_complete_destination_index"]
    sel = d["selective_blocked_boxes_dictionary"]
    lay.n_sel = len(sel)
    for k, (key, boxes) in enumerate(sel.items()):
        lay.sel_key[k] = box_id(key)
        lay.sel_n[k] = len(boxes)
        for j, b in enumerate(boxes):
            lay.sel_boxes[k][j] = box_id(b)
    return lay


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    P = C.c_void_p
    SP = C.POINTER(State)
    sig = {
        "orc_cfg_create": (P, [C.POINTER(Layout)]),
        "orc_cfg_destroy": (None, [P]),
        "orc_cfg_successor": (C.c_int, [P, C.c_int, C.c_int]),
        "orc_cfg_in_no_node_list": (C.c_int, [P, C.c_int]),
        "orc_cfg_base_legal": (C.c_int, [P, C.c_int]),
        "orc_cfg_bfs_len": (C.c_int, [P, C.c_int, C.c_int]),
        "orc_env_create": (P, [P]),
        "orc_env_destroy": (None, [P]),
        "orc_env_set_bad_states": (None, [P, C.c_char_p]),
        "orc_env_reset": (None, [P]),
        "orc_env_initialize_cars": (None, [P, C.c_int]),
        "orc_env_initialize": (None, [P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
        "orc_env_encodestate": (None, [P, C.c_int, C.c_int, SP]),
        "orc_env_legal": (C.c_int, [P, SP]),
        "orc_env_successor_state": (None, [P, SP, C.c_int, SP]),
        "orc_env_reward": (C.c_double, [P, SP, SP, C.c_int, C.c_int]),
        "orc_env_update_car": (None, [P, C.c_int, C.c_int]),
        "orc_env_is_game_ended": (C.c_int, [P]),
        "orc_env_car_done": (C.c_int, [P, C.c_int]),
        "orc_env_car_node": (C.c_int, [P, C.c_int]),
        "orc_env_car_dest_index": (C.c_int, [P, C.c_int]),
        "orc_env_car_done_count": (C.c_int, [P, C.c_int]),
        "orc_env_set_car": (None, [P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "orc_state_to_string": (C.c_int, [SP, C.c_char_p, C.c_int]),
        "orc_state_from_string": (C.c_int, [C.c_char_p, SP]),
        "orc_q_create": (P, [C.c_int]),
        "orc_q_destroy": (None, [P]),
        "orc_q_set_params": (None, [P, C.c_double, C.c_double]),
        "orc_q_get": (C.c_double, [P, SP, C.c_int]),
        "orc_q_peek": (C.c_double, [P, SP, C.c_int, C.POINTER(C.c_int)]),
        "orc_q_set": (None, [P, SP, C.c_int, C.c_double]),
        "orc_q_greedy": (C.c_int, [P, P, SP]),
        "orc_q_value": (C.c_double, [P, P, SP]),
        "orc_q_update": (None, [P, P, SP, C.c_int, SP, C.c_double]),
        "orc_q_size": (C.c_int64, [P]),
        "orc_q_entry": (None, [P, C.c_int64, SP, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
        "orc_epsilon": (C.c_double, [C.c_int, C.c_int, C.c_int]),
        "orc_philox4x32": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int]),
        "orc_philox_rounds": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def state_str(st):
    buf = C.create_string_buffer(128)
    lib().orc_state_to_string(C.byref(st), buf, 128)
    return buf.value.decode()


def state_from_str(s):
    st = State()
    if lib().orc_state_from_string(s.encode(), C.byref(st)) != 0:
        raise ValueError("bad state string %r" % s)
    return st


def philox(ctr, key, rounds=10):
    """Philox4x32 with `rounds` rounds; philox_rounds() = the count the batch semantics are built with."""
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32(c, k, o, rounds)
    return list(o)


def philox_rounds():
    return lib().orc_philox_rounds()


class Config:
    """Static graph, reference src/configuration.py Configuration."""

    def __init__(self, layout_dict):
        self.layout = layout_dict
        self._lay = layout_struct(layout_dict)
        self.h = lib().orc_cfg_create(C.byref(self._lay))
        if not self.h:
            raise ValueError("unsupported layout")

    def successor(self, node, a):
        return lib().orc_cfg_successor(self.h, node, a)

    def base_legal(self, node):
        return lib().orc_cfg_base_legal(self.h, node)

    def bfs_len(self, src, dst):
        return lib().orc_cfg_bfs_len(self.h, src, dst)

    def in_no_node_list(self, node):
        return bool(lib().orc_cfg_in_no_node_list(self.h, node))


def mask_str(m):
    return "".join(ACTIONS[a] for a in range(6) if (m >> a) & 1)


class Env:
    """One scalar env, reference src/env_gym.py Environment (string API kept for the tests)."""

    def __init__(self, cfg, bad_states_text=None):
        self.cfg = cfg
        self.h = lib().orc_env_create(cfg.h)
        self._bad = bad_states_text.encode() if bad_states_text is not None else None
        lib().orc_env_set_bad_states(self.h, self._bad)
        self.n = 0

    def reset(self):
        lib().orc_env_reset(self.h)

    def initialize_cars(self, n):
        self.n = n
        lib().orc_env_initialize_cars(self.h, n)

    def initialize(self, dest_index, start_nodes, choose):
        n = self.n
        di = (C.c_int32 * n)(*dest_index)
        sn = (C.c_int32 * n)(*[node_id(x) if isinstance(x, str) else x for x in start_nodes])
        ch = (C.c_int32 * (3 * n))(*[(list(c) + [0, 0, 0])[:3][k] for c in choose for k in range(3)])
        lib().orc_env_initialize(self.h, di, sn, ch)

    def encodestate(self, car, deadlock_id):
        st = State()
        lib().orc_env_encodestate(self.h, car, deadlock_id, C.byref(st))
        return st

    def legal(self, st):
        return lib().orc_env_legal(self.h, C.byref(st))

    def successor_state(self, st, a):
        out = State()
        lib().orc_env_successor_state(self.h, C.byref(st), a, C.byref(out))
        return out

    def reward(self, prev, nxt, car, dqn=False):
        return lib().orc_env_reward(self.h, C.byref(prev), C.byref(nxt), car, int(dqn))

    def update_car(self, car, node):
        lib().orc_env_update_car(self.h, car, node)

    def is_game_ended(self):
        return bool(lib().orc_env_is_game_ended(self.h))

    def car_done(self, car):
        return bool(lib().orc_env_car_done(self.h, car))

    def car(self, i):
        L = lib()
        return (L.orc_env_car_node(self.h, i), L.orc_env_car_dest_index(self.h, i),
                L.orc_env_car_done(self.h, i), L.orc_env_car_done_count(self.h, i))

    def set_car(self, i, node, dest_index, done=0, done_count=5):
        lib().orc_env_set_car(self.h, i, node_id(node) if isinstance(node, str) else node, dest_index, done, done_count)


class QTable:
    """reference src/algorithms/qlearning.py RL.qvals (a Counter) + update rule."""

    def __init__(self, f32=False, alpha=0.9, discount=0.15):
        self.h = lib().orc_q_create(int(f32))
        lib().orc_q_set_params(self.h, alpha, discount)

    def get(self, st, a):
        return lib().orc_q_get(self.h, C.byref(st), a)

    def set(self, st, a, v):
        lib().orc_q_set(self.h, C.byref(st), a, v)

    def greedy(self, env, st):
        return lib().orc_q_greedy(self.h, env.h, C.byref(st))

    def value(self, env, st):
        return lib().orc_q_value(self.h, env.h, C.byref(st))

    def update(self, env, st, a, nst, r):
        lib().orc_q_update(self.h, env.h, C.byref(st), a, C.byref(nst), r)

    def items(self):
        """[(state_string, action_char, value)] in insertion order (the qvalue .txt line order)."""
        out = []
        st, a, v = State(), C.c_int(), C.c_double()
        for i in range(lib().orc_q_size(self.h)):
            lib().orc_q_entry(self.h, i, C.byref(st), C.byref(a), C.byref(v))
            out.append((state_str(st), ACTIONS[a.value], v.value))
        return out


def epsilon(episodes, episode, model=3):
    return lib().orc_epsilon(episodes, episode, model)


# ---- batched driver (oracle/mdpp_oracle_batch.c) -------------------------------------------------
class Learner(C.Structure):
    _fields_ = [("alpha", C.c_float), ("discount", C.c_float), ("episodes", C.c_int32),
                ("eps_threshold", C.c_int32 * 5), ("eps_q16", C.c_uint32 * 6), ("seed", C.c_uint32),
                ("flags", C.c_uint32)]


def learner(episodes, alpha=0.9, discount=0.15, model=3, seed=3, flags=0):
    """Schedule derived from the oracle's own epsilon function (reference src/utils/utils.py:12-52)."""
    lp = Learner()
    lp.alpha, lp.discount, lp.episodes, lp.seed, lp.flags = alpha, discount, episodes, seed, flags
    eps = [epsilon(episodes, e, model) for e in range(episodes)]
    values = [1.0, 0.8, 0.5, 0.3, 0.15, 0.0] if model == 3 else None
    if model != 3:
        raise NotImplementedError
    for i in range(5):  # threshold i = first episode whose epsilon is below values[i]
        lp.eps_threshold[i] = next((e for e in range(episodes) if eps[e] < values[i]), episodes)
    for i in range(6):
        lp.eps_q16[i] = int(round(values[i] * 65536))
    return lp


_batch_sigs_done = False


def _batch_lib():
    global _batch_sigs_done
    L = lib()
    if not _batch_sigs_done:
        P = C.c_void_p
        I32P = C.POINTER(C.c_int32)
        L.orc_batch_create.restype = P
        L.orc_batch_create.argtypes = [P, C.c_int64, C.c_int, I32P, I32P, I32P, C.c_char_p, C.c_int, C.c_uint32]
        L.orc_batch_destroy.argtypes = [P]
        L.orc_batch_destroy.restype = None
        L.orc_batch_env.restype = P
        L.orc_batch_env.argtypes = [P, C.c_int64]
        L.orc_batch_episode.restype = C.c_int32
        L.orc_batch_episode.argtypes = [P, C.c_int64]
        L.orc_batch_steps.restype = C.c_int32
        L.orc_batch_steps.argtypes = [P, C.c_int64]
        L.orc_batch_stats.restype = None
        L.orc_batch_stats.argtypes = [P, C.POINTER(C.c_uint64)]
        L.orc_batch_reset.restype = None
        L.orc_batch_reset.argtypes = [P, P, C.c_uint32, C.c_int]
        L.orc_batch_rollout_random.restype = None
        L.orc_batch_rollout_random.argtypes = [P, C.c_int, C.c_uint64, C.c_uint32, C.c_int, C.c_int, P, P, P, P]
        L.orc_batch_q_substep.restype = None
        L.orc_batch_q_substep.argtypes = [P, P, C.POINTER(Learner), C.c_int, C.c_uint64, P, P, P, P]
        L.orc_batch_q_round.restype = None
        L.orc_batch_q_round.argtypes = [P, P, C.POINTER(Learner), C.c_uint64]
        L.orc_batch_disagreements.restype = C.c_uint64
        L.orc_batch_disagreements.argtypes = [P]
        L.orc_batch_touched_keys.restype = C.c_uint64
        L.orc_batch_touched_keys.argtypes = [P]
        _batch_sigs_done = True
    return L


def set_threads(n=None):
    """Host threads of the batched learner's per-env phase and of the reachability enumeration (the
    full-size parity tests use every core; results do not depend on the thread count)."""
    L = _batch_lib()
    L.orc_batch_set_threads.restype = None
    L.orc_batch_set_threads.argtypes = [C.c_int]
    L.orc_batch_set_threads(int(n or os.cpu_count() or 1))


class Batch:
    """N scalar oracle envs driven with the batch semantics of DESIGN.md (CPU, for checking)."""

    def __init__(self, cfg, n_envs, boxes, dest_index, choose, bad_states_text=None, step_cap=0, env_id_base=0):
        import numpy as np
        self.np = np
        self.cfg = cfg
        self.n_envs, self.n_cars = n_envs, len(boxes)
        n = self.n_cars
        sb = (C.c_int32 * n)(*[box_id(b) for b in boxes])
        di = (C.c_int32 * n)(*dest_index)
        ch = (C.c_int32 * (3 * n))(*[(list(c) + [0, 0, 0])[:3][k] for c in choose for k in range(3)])
        self._bad = bad_states_text.encode() if bad_states_text is not None else None
        self.h = _batch_lib().orc_batch_create(cfg.h, n_envs, n, sb, di, ch, self._bad, step_cap, env_id_base)

    def __del__(self):
        if getattr(self, "h", None):
            _batch_lib().orc_batch_destroy(self.h)
            self.h = None

    def reset(self, headings=None, seed=3, zero_counters=True):
        np = self.np
        hp = None
        if headings is not None:
            headings = np.ascontiguousarray(headings, dtype=np.uint8)
            hp = headings.ctypes.data_as(C.c_void_p)
        _batch_lib().orc_batch_reset(self.h, hp, seed, int(zero_counters))

    def rollout_random(self, rounds, first_round=0, seed=3, dqn=False, threads=1, outputs=True):
        np = self.np
        shape = (rounds, self.n_cars, self.n_envs)
        out = {}
        if outputs:
            out = {"action": np.empty(shape, np.uint8), "reward": np.empty(shape, np.float32),
                   "done": np.empty(shape, np.uint8), "next": (State * (rounds * self.n_cars * self.n_envs))()}
        p = lambda k: out[k].ctypes.data_as(C.c_void_p) if k in out else None
        _batch_lib().orc_batch_rollout_random(
            self.h, rounds, first_round, seed, int(dqn), threads, p("action"), p("reward"), p("done"),
            C.cast(out["next"], C.c_void_p) if outputs else None)
        return out

    def q_substep(self, q, lp, car, rnd, forced=None, outputs=True):
        np = self.np
        n = self.n_envs
        out = {}
        if outputs:
            out = {"action": np.empty(n, np.uint8), "reward": np.empty(n, np.float32), "state": (State * n)()}
        fp = None
        if forced is not None:
            forced = np.ascontiguousarray(forced, dtype=np.uint8)
            fp = forced.ctypes.data_as(C.c_void_p)
        _batch_lib().orc_batch_q_substep(
            self.h, q.h, C.byref(lp), car, rnd, fp,
            out["action"].ctypes.data_as(C.c_void_p) if outputs else None,
            C.cast(out["state"], C.c_void_p) if outputs else None,
            out["reward"].ctypes.data_as(C.c_void_p) if outputs else None)
        return out

    def q_round(self, q, lp, rnd):
        _batch_lib().orc_batch_q_round(self.h, q.h, C.byref(lp), rnd)

    def disagreements(self):
        """(sub-step, key) groups whose contributing envs computed different values; the batch
        semantics rest on this being 0 (see the header of mdpp_oracle_batch.c)."""
        return int(_batch_lib().orc_batch_disagreements(self.h))

    def touched_keys(self):
        return int(_batch_lib().orc_batch_touched_keys(self.h))

    def env_cars(self, e):
        L = lib()
        h = _batch_lib().orc_batch_env(self.h, e)
        return [(L.orc_env_car_node(h, i), L.orc_env_car_dest_index(h, i), L.orc_env_car_done(h, i),
                 L.orc_env_car_done_count(h, i)) for i in range(self.n_cars)]

    def export(self):
        """Every env at once: cars int16 [n_envs, n_cars, 4] = (node, destination_index, done,
        done_count clamped to -1), episode int32 [n_envs], steps int32 [n_envs]."""
        np = self.np
        L = _batch_lib()
        L.orc_batch_export.restype = None
        L.orc_batch_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        cars = np.empty((self.n_envs, self.n_cars, 4), np.int16)
        episode, steps = np.empty(self.n_envs, np.int32), np.empty(self.n_envs, np.int32)
        L.orc_batch_export(self.h, cars.ctypes.data_as(C.c_void_p), episode.ctypes.data_as(C.c_void_p),
                           steps.ctypes.data_as(C.c_void_p))
        return cars, episode, steps

    def counters(self, e):
        L = _batch_lib()
        return L.orc_batch_episode(self.h, e), L.orc_batch_steps(self.h, e)

    def stats(self):
        o = (C.c_uint64 * 5)()
        _batch_lib().orc_batch_stats(self.h, o)
        rs = o[4] if o[4] < (1 << 63) else o[4] - (1 << 64)
        return {"agent_steps": o[0], "episodes_done": o[1], "episodes_capped": o[2], "explore_steps": o[3],
                "reward_sum": rs}


class CpuBench:
    """Independent oracle learner replicas, one per host thread (CPU baseline of bench.py)."""

    def __init__(self, cfg, threads, envs_per_thread, boxes, dest_index, choose, bad_states_text, step_cap, lp):
        L = _batch_lib()
        L.orc_bench_create.restype = C.c_void_p
        L.orc_bench_create.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_char_p, C.c_int,
                                       C.POINTER(Learner)]
        L.orc_bench_run.restype = C.c_uint64
        L.orc_bench_run.argtypes = [C.c_void_p, C.c_int]
        L.orc_bench_destroy.restype = None
        L.orc_bench_destroy.argtypes = [C.c_void_p]
        n = len(boxes)
        sb = (C.c_int32 * n)(*[box_id(b) for b in boxes])
        di = (C.c_int32 * n)(*dest_index)
        ch = (C.c_int32 * (3 * n))(*[(list(c) + [0, 0, 0])[:3][k] for c in choose for k in range(3)])
        self._bad = bad_states_text.encode() if bad_states_text is not None else None
        self.cfg = cfg
        self.threads = threads
        self.h = L.orc_bench_create(cfg.h, threads, envs_per_thread, n, sb, di, ch, self._bad, step_cap, C.byref(lp))

    def run(self, rounds):
        return int(_batch_lib().orc_bench_run(self.h, rounds))

    def close(self):
        if self.h:
            _batch_lib().orc_bench_destroy(self.h)
            self.h = None


def reachability(cfg, n_cars, choose, n_local_nodes=36, node_base=36):
    """CPU brute-force deadlock reachability bitmap (numpy uint32 words) + number of live states."""
    import numpy as np
    L = _batch_lib()
    L.orc_reachability.restype = C.c_int64
    L.orc_reachability.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_void_p]
    V = 2 * n_local_nodes + 1
    n = V ** n_cars
    ch = (C.c_int32 * (3 * n_cars))(*[(list(c) + [0, 0, 0])[:3][k] for c in choose for k in range(3)])
    out = np.zeros((n + 31) // 32, dtype=np.uint32)
    live = L.orc_reachability(cfg.h, n_cars, ch, n_local_nodes, node_base, out.ctypes.data_as(C.c_void_p))
    return out, int(live)


def convert(st, total_cars=5):
    """reference dqn_open_source.RL.convert for one state -> list of 16*total_cars ints."""
    L = lib()
    L.orc_convert.restype = None
    L.orc_convert.argtypes = [C.POINTER(State), C.c_int, C.c_void_p]
    buf = (C.c_uint8 * (16 * total_cars))()
    L.orc_convert(C.byref(st), total_cars, C.cast(buf, C.c_void_p))
    return list(buf)
