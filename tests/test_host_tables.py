"""Host logic (no GPU): layout dicts, static-table builder, state-id codec, bad-state bitmap and the
C-ABI surface, checked against the reference-recorded goldens and the CPU oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, bad_states_text, load_golden

import marl_dpp_b200 as m
from marl_dpp_b200 import _lib, tables as T
from oracle import oracle as orc


@pytest.fixture(scope="module", autouse=True)
def built_library():
    _lib.build()


@pytest.mark.parametrize("block", [0, 2])
def test_layout_dict_equals_reference(block):
    import json
    g = load_golden("tables_b%d" % block)["layout"]
    assert json.loads(json.dumps(m.layout_3[block])) == g


@pytest.mark.parametrize("block", [0, 2])
def test_static_graph_equals_reference_and_oracle(block):
    g = load_golden("tables_b%d" % block)
    sg = T.StaticGraph(m.layout_3[block])
    ocfg = orc.Config(m.layout_3[block])
    for n, succ in g["action_dictionary"].items():
        nid = T.node_id(n)
        for a, s in enumerate(succ):
            want = T.NONE if s.endswith("YY") else T.node_id(s)
            assert sg.succ[nid, a] == want == ocfg.successor(nid, a), (n, a)
        assert "".join(T.ACTIONS6[a] for a in range(6) if (sg.base_legal6[nid] >> a) & 1) == g["base_legal"][n]
        assert bool(sg.no_node[nid]) == (n in g["no_node_list"])
    for d, row in g["bfs_len"].items():
        for n, length in row.items():
            assert sg.bfs_len(T.node_id(n), T.node_id(d)) == length
    for src in range(72):
        for dst in range(72):
            assert sg.bfs_len(src, dst) == ocfg.bfs_len(src, dst)


SCEN = [
    (0, ["U4"], [1], [[0, 0, 0]]),
    (2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]]),
    (2, ["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]]),
    (2, ["D2", "D6", "D4", "D8"], [0, 0, 0, 0], [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]]),
    (2, ["D1", "D2", "D3", "D5", "D7"], [0, 1, 0, 1, 0], [[0, 1, 0], [2, 1, 0], [1, 0, 0], [2, 2, 0], [3, 4, 0]]),
]


@pytest.mark.parametrize("block,boxes,idx,choose", SCEN)
def test_state_codec_roundtrip_dense(block, boxes, idx, choose):
    """decode is a bijection from [0, n_states) onto canonical strings (sampled), encode inverts it,
    and the strings parse into the oracle's canonical (reference-sorted) form unchanged."""
    sc = T.Scenario(m.layout_3[block], boxes, idx, choose)
    rng = np.random.default_rng(1)
    ids = set(int(x) for x in rng.integers(0, sc.n_states, size=1500)) | {0, sc.n_states - 1}
    seen = set()
    for i in ids:
        s = sc.decode(i)
        assert sc.encode(s) == i
        assert orc.state_str(orc.state_from_str(s)) == s
        seen.add(s)
    assert len(seen) == len(ids)
    with pytest.raises(ValueError):
        sc.decode(sc.n_states)


def test_state_codec_on_reference_strings():
    for name in ("traj_b2_4car_bad", "traj_b0_3car_bad", "traj_b2_5car", "traj_b2_3car_mode0_bad"):
        g = load_golden(name)
        sc = T.Scenario(m.layout_3[g["block"]], g["boxes"], g["idx"], g["choose"])
        strings = {s[1] for ep in g["episodes"] for s in ep["steps"]} | {s[4] for ep in g["episodes"] for s in ep["steps"]}
        ids = {}
        for s in strings:
            i = sc.encode(s)
            assert 0 <= i < sc.n_states
            assert sc.decode(i) == s
            ids[i] = s
        assert len(ids) == len(strings)


@pytest.mark.parametrize("block,boxes,idx,choose", SCEN[1:4])
def test_bad_state_bitmap_equals_string_membership(block, boxes, idx, choose):
    """bit(state) == (heading-stripped state string in bad_states), the reference's test
    (src/env_gym.py:568-577), over a sample of the whole state space."""
    text = bad_states_text(block)
    lines = set(x for x in text.split("\n") if x)
    sc = T.Scenario(m.layout_3[block], boxes, idx, choose, bad_states_text=text)
    rng = np.random.default_rng(2)
    hits = 0
    db = sc.dest_boxes
    for i in rng.integers(0, sc.n_states, size=6000):
        s = sc.decode(int(i))
        stripped = re.sub("[NESW]", "", s)
        ent = [tuple(T.box_id(x) for x in p.split("x")) for p in stripped.split("z")]
        bid = sc._box_state_id(ent)
        assert bid is not None
        bit = (int(sc.bad_bitmap[bid >> 5]) >> (bid & 31)) & 1
        assert bool(bit) == (stripped in lines), s
        hits += bit
    # every matchable line of the file is representable
    total = int(sum(bin(int(w)).count("1") for w in sc.bad_bitmap))
    assert total > 0 and hits >= 0


def test_epsilon_schedule_equals_reference():
    g = load_golden("known_answers")["eps_model3"]
    for E in (100, 60, 7, 5000):
        th, eps = T.epsilon_schedule(E, 3)

        def eps_of(e):
            for t, v in zip(th, eps):
                if e < t:
                    return v
            return eps[5]
        got = [eps_of(e) for e in range(E)]
        assert got == [orc.epsilon(E, e, 3) for e in range(E)]
        want = g[str(E)]
        if want and isinstance(want[0], list):
            assert [[i, got[i]] for i in range(E) if i == 0 or got[i] != got[i - 1]] == want
        else:
            assert got == want


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mdpp.h")).read()
    declared = set(re.findall(r"\b(mdpp_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"mdpp_handle"}
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTS)
    assert _lib.lib().mdpp_abi_version() == 1


def test_python_constants_match_the_header():
    """The learner flags, forced-action codes and reward kinds the Python layer passes are the header's."""
    from marl_dpp_b200 import batched
    hdr = open(os.path.join(ROOT, "include", "mdpp.h")).read()
    defs = {k: int(v.rstrip("u"), 0) for k, v in re.findall(r"#define\s+(MDPP_[A-Z_0-9]+)\s+(0x[0-9A-Fa-f]+u?|\d+u?)\b", hdr)}
    assert (defs["MDPP_LEARN_FROZEN"], defs["MDPP_LEARN_MODE0"], defs["MDPP_LEARN_DETECT"], defs["MDPP_LEARN_OBSERVED"],
            defs["MDPP_LEARN_RANDOM_ORDER"]) == (T.LEARN_FROZEN, T.LEARN_MODE0, T.LEARN_DETECT, T.LEARN_OBSERVED,
                                                 T.LEARN_RANDOM_ORDER)
    assert defs["MDPP_ACT_SKIP"] == batched.ACT_SKIP and defs["MDPP_ACT_GREEDY"] == batched.ACT_GREEDY
    enums = dict(re.findall(r"\b(MDPP_REWARD_[A-Z]+)\s*=\s*(\d+)", hdr))
    assert int(enums["MDPP_REWARD_Q"]) == batched.REWARD_KINDS["q_reward"]
    assert int(enums["MDPP_REWARD_DQN"]) == batched.REWARD_KINDS["dqn_reward"]
    assert defs["MDPP_ABI_VERSION"] == _lib.lib().mdpp_abi_version()


def test_workspace_size_and_no_cpu_fallback():
    import torch
    sc = T.Scenario(m.layout_3[2], ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]])
    L = _lib.lib()
    b1 = L.mdpp_workspace_bytes(C.byref(sc.cfg), 1 << 20)
    b2 = L.mdpp_workspace_bytes(C.byref(sc.cfg), 1 << 21)
    assert 16 * (1 << 20) < b1 < b2 < 200 * (1 << 20)
    bad = T.MdppConfig()
    assert L.mdpp_workspace_bytes(C.byref(bad), 10) == -1
    if not torch.cuda.is_available():
        from marl_dpp_b200.batched import BatchedEnv
        with pytest.raises(RuntimeError):
            BatchedEnv(sc, 16)
        h = C.c_void_p()
        rc = L.mdpp_create(C.byref(sc.cfg), 16, 0, C.c_void_p(256), 1 << 30, None, None, C.byref(h))
        assert rc in (2, 3) and not h.value  # MDPP_ENODEVICE / MDPP_ECUDA: never a silent CPU path


def _fast_tables(sc):
    """Host-built table image of the table-driven kernels, decoded into numpy arrays."""
    L = _lib.lib()
    size = L.mdpp_fast_tables(C.byref(sc.cfg), None, 0)
    buf = np.zeros(size, dtype=np.uint8)
    assert L.mdpp_fast_tables(C.byref(sc.cfg), buf.ctypes.data_as(C.c_void_p), size) == size
    off = 0
    own = buf[off:off + 5 * 256 * 16].view(np.uint32).reshape(5, 256, 4); off += 5 * 256 * 16
    oth = buf[off:off + 5 * 256 * 8].view(np.uint32).reshape(5, 256, 2); off += 5 * 256 * 8
    nxt = buf[off:off + 5 * 256 * 8 * 2].view(np.uint16).reshape(5, 256, 8); off += 5 * 256 * 8 * 2
    nth = buf[off:off + 32 * 4].view(np.uint32); off += 32 * 4
    ghost = buf[off:off + 5 * 8].view(np.uint32).reshape(5, 2)
    return own, oth, ghost, nth, nxt


@pytest.mark.parametrize("block,boxes,idx,choose", SCEN)
def test_fast_tables_replay_the_oracle_env(block, boxes, idx, choose):
    """The round / rollout kernels read everything the reference derives from a car's (node,
    destination_index) from host-built tables.  Here the kernel's step (csrc/q_round.cuh round_car_step /
    rollout_car_step) is restated in Python on those tables and replayed against the oracle env from random
    joint states: state string, legal set, and for every legal action the successor, the car after
    update_car and both rewards must agree -- ghosts of finished cars included."""
    rng = np.random.default_rng(7)
    sc = m.Scenario(m.layout_3[block], boxes, idx, choose, bad_states_text=None)
    own, oth, ghost, nth, nxt = _fast_tables(sc)
    n = sc.n_cars
    ocfg = orc.Config(m.layout_3[block])
    env = orc.Env(ocfg)
    env.reset()
    env.initialize_cars(n)
    env.initialize(idx, [b + "N" for b in boxes], choose)
    nodes = [nd for nd in range(sc.node_base, sc.node_base + sc.n_local_nodes) if not ocfg.in_no_node_list(nd)]
    A5TO6 = [0, 1, 2, 3, 5]
    stride = len(sc.dest_nodes) * sc.n_local_nodes
    checked = 0
    for trial in range(300):
        # a random joint state: distinct boxes for the cars still driving, some cars finished with a ghost count
        fields = []
        used = set()
        for i in range(n):
            done = trial % 3 == 0 and rng.random() < 0.3
            if done:
                cnt = int(rng.integers(-2, 6))
                env.set_car(i, sc.parking, int(rng.integers(0, 2)), 1, cnt)
                fields.append((sc.parking, 0, 1, cnt))
                continue
            while True:
                nd = int(rng.choice(nodes))
                if (nd >> 2) % 9 not in used:
                    break
            used.add((nd >> 2) % 9)
            di = int(rng.integers(0, 2))
            env.set_car(i, nd, di, 0, 5)
            fields.append((nd, di, 0, 5))
        c = int(rng.integers(0, n))
        if fields[c][2]:
            continue
        node, di = fields[c][0], fields[c][1]
        o = own[c][node | (di << 7)]
        codes, occ9, occ18 = [], 0, 0
        for i in range(n):
            if i == c:
                continue
            nd_i, di_i, done_i, cnt_i = fields[i]
            e = (0, 0)
            if not done_i:
                e = oth[i][nd_i | (di_i << 7)]
            elif cnt_i > 0:          # training_done_count - 1 >= 0: the ghost is still listed (src/env_gym.py:307-318)
                e = ghost[i]
            codes.append(int(e[0]) >> 18)
            occ18 |= int(e[0]) & 0x3FFFF
            occ9 |= int(e[1])
        codes.sort()
        rank = sum(_comb_ref(cd + j, j + 1) for j, cd in enumerate(codes))
        sidx = rank * stride + (int(o[0]) & 0x3FF)
        st = env.encodestate(c, 1)
        assert sc.decode(sidx) == orc.state_str(st), trial
        lm = (int(o[0]) >> 24) & 31
        hit = (occ9 * 0x201) & int(o[1])
        if hit & 0x1FF:
            lm &= 8
        if (hit >> 9) or (((int(o[1]) >> 18) & 1) and (int(o[2]) & occ18)):
            lm &= ~8
        legal6 = env.legal(st)
        assert lm == (legal6 & 0xF) | (((legal6 >> 5) & 1) << 4), (trial, orc.state_str(st))
        dnode = (int(o[0]) >> 10) & 127
        for a in range(5):
            if not (lm >> a) & 1:
                continue
            mv = int(nxt[c][node | (di << 7)][a])
            reached = (mv >> 9) & 1
            nn = dnode if reached else mv & 127
            ns = env.successor_state(st, A5TO6[a])
            assert ns.pnode == nn
            for dqn in (False, True):   # no bad-states file here: the arithmetic branch of q_reward / dqn_reward
                r = -10 + 50 * ((int(o[3]) & 0xFF) - (mv >> 10)) + ((110 if dqn else 10010) if reached else 0)
                want = env.reward(st, ns, c, dqn)
                assert (r / 100.0 if dqn else float(r)) == want, (trial, a, dqn)
            # update_car on a scratch copy of the car
            env.update_car(c, nn)
            got = env.car(c)
            assert (got[0], got[1], got[2]) == (mv & 127, (mv >> 7) & 1, (mv >> 8) & 1), (trial, a)
            env.set_car(c, node, di, 0, 5)
            checked += 1
        # encodestate(., 1) moved the ghost counters: put the finished cars back for the next trial
    assert checked > 500
    for mask in range(32):
        bits = [k for k in range(5) if (mask >> k) & 1]
        for j, k in enumerate(bits):
            assert (int(nth[mask]) >> (3 * j)) & 7 == k


def _comb_ref(n, k):
    from math import comb
    return comb(n, k) if n >= k else 0


def test_dqn_reward_division_needs_no_float64():
    """dqn_reward returns reward/100 in float64 (reference src/env_gym.py:553); the kernels store float32 and
    divide in single precision.  Exhaustively: for every integer reward the kernels can produce (and far beyond)
    the correctly rounded float32 quotient equals the float64 quotient rounded to float32."""
    r = np.arange(-2_000_000, 2_000_001, dtype=np.int64)
    via_f64 = (r.astype(np.float64) / 100.0).astype(np.float32)
    via_f32 = r.astype(np.float32) / np.float32(100.0)
    assert via_f32.dtype == np.float32 and np.array_equal(via_f64, via_f32)


def test_scan_slot_division_trick():
    """csrc/q_learner.cuh scan_slot: bit = m / NW by one multiply-high with magic = floor(2^32 / NW) + 1 and a
    one-step correction (m < 32 * NW, NW < 2^27).  The same arithmetic in Python, over the bitmap sizes of the real
    tables (1-5 cars, layout 3) and the edges of every bit plane."""
    import random
    rng = random.Random(7)
    for n_states in (108, 6660, 126540, 1645020, 16450200, 3, 4_000_000_000 // 5 // 32):
        marks = n_states * 5
        nw = ((marks + 31) // 32 + 3) & ~3
        assert nw < 1 << 27
        magic = (1 << 32) // nw + 1
        assert magic < 1 << 32
        probes = [0, marks - 1] + [b * nw + d for b in range(32) for d in (-1, 0, 1)] + [rng.randrange(marks) for _ in range(2000)]
        for m in probes:
            if not 0 <= m < marks:
                continue
            b = (m * magic) >> 32
            w = (m - b * nw) & 0xFFFFFFFF
            if w >= 1 << 31:                     # (int32_t)w < 0
                b -= 1
                w = (w + nw) & 0xFFFFFFFF
            assert (b, w) == divmod(m, nw) and b < 32, (n_states, m)
