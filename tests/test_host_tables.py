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
