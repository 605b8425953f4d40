"""Host-side formats and the Configuration object against reference-recorded goldens (no GPU)."""
import pytest

from conftest import bad_states_text, load_golden

from marl_dpp_b200 import formats, layout_3
from marl_dpp_b200.configuration import Configuration


@pytest.mark.parametrize("block", [0, 2])
def test_configuration_matches_reference(block):
    t = load_golden("tables_b%d" % block)
    c = Configuration(layout_3[block])
    assert c.action_dictionary == t["action_dictionary"]
    assert list(c.action_dictionary) == t["node_order"]          # dict order quirk U1-3, D1-9, U4-9
    assert c.no_node_list == t["no_node_list"]
    assert c.destination_nodes_list == t["destination_nodes_list"]
    assert c.station_box_list == t["station_box_list"] and c.no_box_list == t["no_box_list"]
    for n in t["node_order"]:
        assert "".join(c.get_legal_actions(n)) == t["base_legal"][n]
    for d, row in t["bfs_len"].items():
        for n, length in row.items():
            acts, boxes = c.breadth_first_search(n, d)
            assert len(acts) == length
    assert layout_3[block]["no_node_list"] == []                  # the layout dict is not mutated


def test_deadlock_expansion_matches_reference():
    g = load_golden("known_answers")["deadlock_expand_max4"]
    for state, want in g.items():
        got = formats.expand_deadlock_state(state, ["D6", "D4", "D8", "D7"], state.count("z") + 1, 4)
        assert sorted(got) == want


def test_save_deadlock_state_unions_with_file(tmp_path):
    p = tmp_path / "bad.txt"
    p.write_text("D1xD4zD2xD6\nD5xD6zD6xD8\n")
    final = formats.save_deadlock_state("D5NxD6NzD6xD8", ["D6", "D4", "D8", "D7"], str(p), 2, 3)
    lines = formats.read_bad_states(str(p))
    assert set(lines) == final and "D1xD4zD2xD6" in final and "D6xD8zD5xD6" in final
    assert "D5xD6zD6xD8zD4xD4" in final and len(lines) == len(set(lines))
    # the shipped fixtures are fixed points of their own canonical 2-car states
    text = bad_states_text(2)
    known = set(x for x in text.split("\n") if x)
    two = [s for s in known if s.count("z") == 1]
    assert two and all(formats.strip_headings(s) == s for s in known)


def test_qvalue_file_roundtrip(tmp_path):
    items = [("D2NxD6NzD3xD6", "F", -5.312068462371826), ("D5ExD6N", "S", 0.0), ("U4WxU2W", "T", 9990.000000000002)]
    path = formats.qvalue_path("7", str(tmp_path))
    formats.write_qvalues(path, items)
    assert open(path).read().split("\n")[0] == "D2NxD6NzD3xD6|F|-5.312068462371826"
    assert formats.read_qvalues(path) == items
    # a reference-recorded Counter written in the reference's format reloads exactly
    g = load_golden("qrun_b0_1car")["qvals"]
    formats.write_qvalues(path, [(s, a, float(v)) for s, a, v in g])
    assert [(s, a, repr(v)) for s, a, v in formats.read_qvalues(path)] == [tuple(x) for x in g]
