"""Static tables: layout dict -> the constant tables the CUDA kernels index.

Replaces, on the host and once per (block, initial state):
  * Configuration.__init__ and give_*_action_node  (reference src/configuration.py:44-266)
    -> StaticGraph.succ[node][S,L,R,F,B,T]
  * get_legal_actions                              (reference src/env_gym.py:114-130) -> base_legal
  * breadth_first_search + src_dest_length_dict    (reference src/env_gym.py:443-483,585-596)
    -> bfs_len, a dense table instead of a per-episode memo
  * the bad-states file read                       (reference src/env_gym.py:71-83)
    -> one bit per heading-stripped state id
The graph is built by grid arithmetic on (floor, row, col, heading), not by string surgery.

Node id = floor*36 + (box-1)*4 + heading; floor U=0, D=1; heading N,E,S,W = 0..3; box18 = node>>2.
"""
import ctypes as C
import re
from collections import deque

import numpy as np

from . import _lib

NONE = 255
HEADS = "NESW"
ACTIONS6 = "SLRFBT"   # reference action order, src/env_gym.py:6-21
ACTIONS5 = "SLRFT"    # what survives the legal filter ('B' is always dropped, :158-159)
# (d_row, d_col) of one step forward per heading: N is row-1, S row+1, E col+1, W col-1
_STEP = ((-1, 0), (0, 1), (1, 0), (0, -1))

_NODE_RE = re.compile(r"([UD])(\d+)([NESW])")
_BOX_RE = re.compile(r"([UD])(\d+)")


def node_id(s):
    m = _NODE_RE.fullmatch(s)
    if not m:
        raise ValueError("not a node name: %r" % (s,))
    return (m.group(1) == "D") * 36 + (int(m.group(2)) - 1) * 4 + HEADS.index(m.group(3))


def box_id(s):
    m = _BOX_RE.fullmatch(s)
    if not m:
        raise ValueError("not a box name: %r" % (s,))
    return (m.group(1) == "D") * 9 + int(m.group(2)) - 1


def node_name(n):
    return "UD"[n // 36] + str((n % 36) // 4 + 1) + HEADS[n & 3]


def box_name(b):
    return "UD"[b // 9] + str(b % 9 + 1)


class StaticGraph:
    """Successor table of one layout block."""

    def __init__(self, layout):
        rows, cols = layout["number_of_rows"], layout["number_of_cols"]
        if (rows, cols) != (3, 3):
            raise ValueError("only 3x3 blocks are supported (as in the reference's layout_3)")
        self.layout = layout
        self.rows, self.cols = rows, cols
        if len(layout["station_destination_list"]) != 2:
            raise ValueError("station_destination_list must hold exactly two lists")
        # station box -> bitmask of headings listed as its entrances; first listing wins
        self.entrance = {}
        for box, sides in zip(layout["station_box_list"], layout["station_entrance_list"]):
            self.entrance.setdefault(box_id(box), sum(1 << HEADS.index(h) for h in sides))
        self.disconnected = {node_id(n) for n in layout["disconnected_nodes_list"]}
        self.no_boxes = [box_id(b) for b in layout["no_box_list"]]
        self.no_node = np.zeros(72, dtype=bool)
        for n in layout.get("no_node_list", []):
            self.no_node[node_id(n)] = True
        for b in self.no_boxes:
            self.no_node[b * 4:b * 4 + 4] = True
        self.succ = np.full((72, 6), NONE, dtype=np.uint8)
        for n in range(72):
            h = n & 3
            self.succ[n, 0] = n
            self.succ[n, 1] = (n & ~3) | ((h - 1) & 3)
            self.succ[n, 2] = (n & ~3) | ((h + 1) & 3)
            self.succ[n, 3] = self._move(n, +1)
            self.succ[n, 4] = self._move(n, -1)
            self.succ[n, 5] = (n & ~3) | ((h + 2) & 3)
        # base legal set: successor exists and is not a no_node
        self.base_legal6 = np.zeros(72, dtype=np.uint8)
        for n in range(72):
            for a in range(6):
                s = self.succ[n, a]
                if s != NONE and not self.no_node[s]:
                    self.base_legal6[n] |= 1 << a
        self._dist = {}

    def _move(self, n, sign):
        """One box forward (sign=+1) or backward (-1) keeping the heading, or NONE.

        Station rule: a car drives through a station box only along a listed entrance axis --
        forward it must face a listed side when leaving and face away from a listed side when
        entering; backward the two tests swap.  Only the forward move honours
        disconnected_nodes_list."""
        floor, box, h = n // 36, (n % 36) // 4, n & 3
        opp = (h + 2) & 3
        leave_side, enter_side = (h, opp) if sign > 0 else (opp, h)
        ent = self.entrance.get(floor * 9 + box)
        if ent is not None and not (ent >> leave_side) & 1:
            return NONE
        row, col = box % self.rows, box // self.rows
        row += sign * _STEP[h][0]
        col += sign * _STEP[h][1]
        if not (0 <= row < self.rows and 0 <= col < self.cols):
            return NONE
        out = floor * 36 + (col * self.rows + row) * 4 + h
        if sign > 0 and out in self.disconnected:
            return NONE
        ent = self.entrance.get(out >> 2)
        if ent is not None and not (ent >> enter_side) & 1:
            return NONE
        return out

    def bfs_len(self, src, dst):
        """len(actions) the reference's BFS returns: shortest path length over S,L,R,F,T, except
        that src==dst and "unreachable" both give 1."""
        d = self._dist.get(src)
        if d is None:
            d = np.full(72, -1, dtype=np.int32)
            d[src] = 0
            dq = deque([src])
            while dq:
                u = dq.popleft()
                for a in (0, 1, 2, 3, 5):
                    if not (self.base_legal6[u] >> a) & 1:
                        continue
                    v = int(self.succ[u, a])
                    if d[v] < 0:
                        d[v] = d[u] + 1
                        dq.append(v)
            self._dist[src] = d
        return int(d[dst]) if d[dst] > 0 else 1


# ---- ctypes mirrors of include/mdpp.h ------------------------------------------------------------
class MdppConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_cars", C.c_int32), ("n_local_nodes", C.c_int32),
                ("n_dest_nodes", C.c_int32), ("n_dest_boxes", C.c_int32), ("n_src_boxes", C.c_int32),
                ("parking_node", C.c_int32), ("tcdi", C.c_int32), ("bad_states_on", C.c_int32),
                ("step_cap", C.c_int32),
                ("node_local", C.c_uint8 * 72), ("fwd", C.c_uint8 * 72), ("base_legal", C.c_uint8 * 72),
                ("bfs_len", (C.c_uint8 * 72) * 8), ("dest_node", C.c_uint8 * 8),
                ("dest_node_box", C.c_uint8 * 8), ("dest_box", C.c_uint8 * 8), ("src_box_idx", C.c_uint8 * 18),
                ("sel_mask", C.c_uint32 * 8), ("car_dest", (C.c_uint8 * 2) * 5),
                ("start_box", C.c_uint8 * 5), ("start_di", C.c_uint8 * 5),
                ("n_states", C.c_int64), ("n_box_states", C.c_int64)]


class MdppLearner(C.Structure):
    _fields_ = [("alpha", C.c_float), ("discount", C.c_float), ("episodes", C.c_int32),
                ("eps_threshold", C.c_int32 * 5), ("eps_q16", C.c_uint32 * 6), ("seed", C.c_uint32), ("flags", C.c_uint32)]


class MdppStats(C.Structure):
    _fields_ = [("agent_steps", C.c_uint64), ("q_updates", C.c_uint64), ("episodes_done", C.c_uint64),
                ("episodes_capped", C.c_uint64), ("reward_sum", C.c_int64), ("explore_steps", C.c_uint64),
                ("touched_keys", C.c_uint64), ("internal_errors", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def _comb(n, k):
    r = 1
    for i in range(1, k + 1):
        r = r * (n - k + i) // i
    return r


class Scenario:
    """One (layout block, initial state) pair compiled to an mdpp_config.

    boxes / dest_index / choose are an entry of the layout's `initial_states`
    (reference src/layout.py:960-1000): start boxes, starting destination_index per car, and per
    car the indices into station_destination_list[0] and [1]."""

    def __init__(self, layout, boxes, dest_index, choose, bad_states_text=None, step_cap=0):
        self.layout = layout
        self.graph = g = StaticGraph(layout)
        n = len(boxes)
        if not 1 <= n <= 5:
            raise ValueError("1..5 cars supported, got %d" % n)
        if len(dest_index) != n or len(choose) != n:
            raise ValueError("dest_index / choose must have one entry per car")
        self.n_cars = n
        self.boxes, self.dest_index, self.choose = list(boxes), list(dest_index), [list(c) for c in choose]
        lists = [[node_id(x) for x in lst] for lst in layout["station_destination_list"]]
        self.dest_nodes = []
        for lst in lists:
            for d in lst:
                if d not in self.dest_nodes:
                    self.dest_nodes.append(d)
        self.dest_boxes = []
        for d in self.dest_nodes:
            if d >> 2 not in self.dest_boxes:
                self.dest_boxes.append(d >> 2)
        if len(self.dest_nodes) > 8:
            raise ValueError("more than 8 distinct destination nodes")
        self.parking = g.no_boxes[0] * 4 + HEADS.index("W")
        self.tcdi = layout["training_complete_destination_index"]
        start_boxes = [box_id(b) for b in boxes]
        floors = {d // 36 for d in self.dest_nodes} | {b // 9 for b in start_boxes} | {self.parking // 36}
        one_floor = len(floors) == 1
        floor = next(iter(floors)) if one_floor else 0
        self.n_local_nodes = 36 if one_floor else 72
        self.n_src_boxes = 9 if one_floor else 18
        self.node_base = floor * 36 if one_floor else 0
        self.box_base = floor * 9 if one_floor else 0
        k = n - 1
        self.n_codes = 1 + self.n_src_boxes * len(self.dest_boxes)   # 0 = "not listed"
        self.n_multisets = _comb(self.n_codes + k - 1, k)
        self.n_states = self.n_multisets * len(self.dest_nodes) * self.n_local_nodes
        self.n_box_states = self.n_multisets * len(self.dest_boxes) * self.n_src_boxes
        if self.n_states * 8 >= 1 << 32:
            raise ValueError("state space too large for 32-bit keys")

        cfg = MdppConfig()
        cfg.abi_version = 1
        cfg.n_cars = n
        cfg.n_local_nodes = self.n_local_nodes
        cfg.n_dest_nodes = len(self.dest_nodes)
        cfg.n_dest_boxes = len(self.dest_boxes)
        cfg.n_src_boxes = self.n_src_boxes
        cfg.parking_node = self.parking
        cfg.tcdi = self.tcdi
        cfg.bad_states_on = int(bad_states_text is not None)
        cfg.step_cap = step_cap
        for nd in range(72):
            inside = self.node_base <= nd < self.node_base + self.n_local_nodes
            cfg.node_local[nd] = nd - self.node_base if inside else NONE
            f = int(g.succ[nd, 3])
            legal6 = int(g.base_legal6[nd])
            legal5 = (legal6 & 0xF) | ((legal6 >> 5) & 1) << 4          # drop 'B'
            cfg.fwd[nd] = f if (legal6 >> 3) & 1 else NONE
            cfg.base_legal[nd] = legal5
        for i, d in enumerate(self.dest_nodes):
            cfg.dest_node[i] = d
            cfg.dest_node_box[i] = self.dest_boxes.index(d >> 2)
            for nd in range(72):
                cfg.bfs_len[i][nd] = g.bfs_len(nd, d)
        for i, b in enumerate(self.dest_boxes):
            cfg.dest_box[i] = b
            mask = 0
            for key, blocked in layout["selective_blocked_boxes_dictionary"].items():
                if box_id(key) == b:
                    for x in blocked:
                        mask |= 1 << box_id(x)
            cfg.sel_mask[i] = mask
        for b in range(18):
            inside = self.box_base <= b < self.box_base + self.n_src_boxes
            cfg.src_box_idx[b] = b - self.box_base if inside else NONE
        for i in range(n):
            for di in range(2):
                cfg.car_dest[i][di] = self.dest_nodes.index(lists[di][self.choose[i][di]])
            cfg.start_box[i] = start_boxes[i]
            cfg.start_di[i] = dest_index[i]
        cfg.n_states = self.n_states
        cfg.n_box_states = self.n_box_states
        self.cfg = cfg
        self.bad_bitmap = self._bad_bitmap(bad_states_text) if bad_states_text is not None else None

    # ---- state ids <-> the reference's state strings (host side, via the library's codec) -----
    def encode(self, state):
        """'D2NxD6NzD3xD6' -> state id (row of the Q table)."""
        parts = [p.split("x") for p in state.split("z")]
        others = []
        for s, d in parts[1:]:
            others += [box_id(s), box_id(d)]
        arr = (C.c_int32 * max(len(others), 1))(*others)
        idx = _lib.lib().mdpp_encode_state(C.byref(self.cfg), node_id(parts[0][0]), node_id(parts[0][1]),
                                           len(parts) - 1, arr)
        if idx < 0:
            raise ValueError("cannot encode %r: %s" % (state, _lib.last_error()))
        return idx

    def decode(self, idx):
        pn, dn = C.c_int32(), C.c_int32()
        others = (C.c_int32 * 8)()
        k = _lib.lib().mdpp_decode_state(C.byref(self.cfg), int(idx), C.byref(pn), C.byref(dn), others)
        if k < 0:
            raise ValueError("cannot decode %d: %s" % (idx, _lib.last_error()))
        s = node_name(pn.value) + "x" + node_name(dn.value)
        for i in range(k):
            s += "z" + box_name(others[2 * i]) + "x" + box_name(others[2 * i + 1])
        return s

    def _box_state_id(self, entries):
        """Heading-stripped state [(src box, dst box), ...] -> bit index, or None if no encoded
        state can ever produce this string (unsorted others, unknown boxes, too many cars)."""
        (ps, pd), rest = entries[0], entries[1:]
        if len(rest) > self.n_cars - 1 or pd not in self.dest_boxes:
            return None
        if not self.box_base <= ps < self.box_base + self.n_src_boxes:
            return None
        # ascend_array order (reference src/env_gym.py:267-279): box number, then the strings
        key = lambda e: (e[0] % 9, box_name(e[0]), box_name(e[1]))
        if [key(e) for e in rest] != sorted(key(e) for e in rest):
            return None
        codes = []
        for s, d in rest:
            if d not in self.dest_boxes or not self.box_base <= s < self.box_base + self.n_src_boxes:
                return None
            codes.append(1 + (s - self.box_base) * len(self.dest_boxes) + self.dest_boxes.index(d))
        codes = sorted([0] * (self.n_cars - 1 - len(codes)) + codes)
        rank = sum(_comb(c + j, j + 1) for j, c in enumerate(codes))
        return (rank * len(self.dest_boxes) + self.dest_boxes.index(pd)) * self.n_src_boxes + (ps - self.box_base)

    def _bad_bitmap(self, text):
        bits = np.zeros((self.n_box_states + 31) // 32 + 1, dtype=np.uint32)
        for line in text.split("\n"):
            if line == "":
                continue
            try:
                entries = [tuple(box_id(x) for x in part.split("x")) for part in line.split("z")]
                if any(len(e) != 2 for e in entries):
                    continue
            except ValueError:
                continue
            i = self._box_state_id(entries)
            if i is not None:
                bits[i >> 5] |= np.uint32(1 << (i & 31))
        return bits


def epsilon_schedule(episodes, model=3):
    """Thresholds and epsilon values of give_action_choose_probability
    (reference src/utils/utils.py:12-52); `int(episodes * 0.7)` etc. in float64 like the reference."""
    if model == 0:
        return [0] * 5, [0.0] * 6
    if model == 1:
        return [episodes + 1] * 5, [1.0] * 6
    if model == 2:
        return [int(episodes * 0.4)] * 5, [1.0, 0.0, 0.0, 0.0, 0.0, 0.0]
    if model == 3:
        return ([int(episodes * 0.4), int(episodes * 0.6), int(episodes * 0.7), int(episodes * 0.85),
                 int(episodes * 0.99)], [1.0, 0.8, 0.5, 0.3, 0.15, 0.0])
    raise ValueError("unknown probability model %r" % (model,))


LEARN_FROZEN, LEARN_MODE0, LEARN_DETECT, LEARN_OBSERVED, LEARN_RANDOM_ORDER = 1, 2, 4, 8, 16


def learner_params(episodes, alpha=0.9, discount=0.15, model=3, seed=3, flags=0):
    if not 0 <= int(episodes) <= 65535:
        raise ValueError("episodes must be in 0..65535 (the per-env episode counter is 16 bits), got %r" % (episodes,))
    th, eps = epsilon_schedule(episodes, model)
    lp = MdppLearner()
    lp.alpha, lp.discount, lp.episodes, lp.seed, lp.flags = alpha, discount, episodes, seed, flags
    for i in range(5):
        lp.eps_threshold[i] = th[i]
    for i in range(6):
        lp.eps_q16[i] = int(round(eps[i] * 65536))
    return lp
