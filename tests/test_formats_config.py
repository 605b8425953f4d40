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


def test_configuration_accepts_the_references_object():
    """A stand-in with the reference Configuration's attribute names (src/configuration.py:44-145)."""
    from types import SimpleNamespace
    d = layout_3[2]
    ref_like = SimpleNamespace(
        environment_rows=3, environment_columns=3, environment_bad_states_file=d["bad_states_file"],
        no_box_list=d["no_box_list"], no_node_list=["D9N", "D9E", "D9S", "D9W"], no_floor=1,
        environment_node_name_list=d["environment_node_name_list"], is_destination_node=True,
        no_of_stations=3, disconnected_nodes_list=d["disconnected_nodes_list"],
        station_box_list=d["station_box_list"], destination_nodes_list=d["station_destination_list"],
        station_processing_count_list=d["station_processing_count_list"],
        station_entrance_list=d["station_entrance_list"], no_of_cars=3,
        training_complete_destination_index=1,
        selective_blocked_nodes_dictionary=d["selective_blocked_nodes_dictionary"],
        selective_blocked_boxes_dictionary=d["selective_blocked_boxes_dictionary"],
        cars_info=d["cars_info"], initial_states=d["initial_states"])
    c = Configuration.coerce(ref_like)
    assert c.action_dictionary == Configuration(d).action_dictionary
    assert c.no_node_list == ["D9N", "D9E", "D9S", "D9W"] and c.initial_states == d["initial_states"]
    assert Configuration.coerce(c) is c and Configuration.coerce(d).layout["no_box_list"] == ["D9"]


def test_block_renaming_matches_reference(tmp_path):
    """convert_textfile / reverse_convert_textfile (reference src/algorithms/qlearning.py:457-597):
    byte-identical files on every block number, incl. blank lines and a missing final newline."""
    g = load_golden("textfile_convert")
    assert sorted({c["block"] for c in g["cases"]}) == [0, 1, 2, 3]
    for c in g["cases"]:
        (tmp_path / "in.txt").write_text(c["input"])
        formats.convert_textfile("in", "out", c["block"], directory=str(tmp_path))
        assert (tmp_path / "out.txt").read_text() == c["converted"]
        formats.convert_textfile("out", "back", c["block"], directory=str(tmp_path), reverse=True)
        assert (tmp_path / "back.txt").read_text() == c["reversed"]
        # the renaming is a bijection on box names: undoing it restores every line
        assert c["reversed"].split("\n")[:len(c["input"].split("\n"))] == c["input"].split("\n")
    with pytest.raises(KeyError):
        formats.block_box_map(4)


def test_plugin_file_conversions_match_reference(tmp_path):
    """The same three methods through the RL plugin (they need no device)."""
    from marl_dpp_b200.algorithms.qlearning import RL
    g = load_golden("textfile_convert")
    rl = RL(input_dim=80, output_dim=5, cars=5, sess=None, env_config=layout_3[2], qvalue_dir=str(tmp_path),
            bad_states="")
    a = g["ascending"]
    (tmp_path / "asc_in.txt").write_text(a["input"])
    rl.convert_textfile_in_ascending_order("asc_in", "asc_out")
    assert (tmp_path / "asc_out.txt").read_text() == a["output"]
    assert [[s, act, repr(v)] for (s, act), v in rl.qvals_new.items()] == a["qvals_new"]
    c = g["cases"][2]
    (tmp_path / "cv_in.txt").write_text(c["input"])
    rl.convert_textfile("cv_in", "cv_out", c["block"])
    rl.reverse_convert_textfile("cv_out", "cv_back", c["block"])
    assert (tmp_path / "cv_out.txt").read_text() == c["converted"]
    assert (tmp_path / "cv_back.txt").read_text() == c["reversed"]
    assert formats.ascend_state("D5ExD6NzD8xD8zD1xD6") == load_golden("known_answers")["ascend"]
