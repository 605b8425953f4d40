"""On-disk formats of the hot path, host side.

* qvalue files: one `state|action|repr(float)` line per key
  (reference src/algorithms/qlearning.py:159-185); name '0' means "no file".
* bad-states files: one heading-stripped state string per line, read by the env
  (reference src/env_gym.py:71-83) and grown by save_deadlock_state
  (reference src/utils/utils.py:141-199).
* the block -> whole-floor ("MSV") renaming of qvalue files, and its inverse
  (reference src/algorithms/qlearning.py:457-597), plus the re-keying of a table into
  ascend_array order (reference :439-455).
"""
import itertools
import os

QVALUE_DIR = "./src/data/qvalue_files/layout-3/"  # hard-coded in the reference (qlearning.py:160,170)
QVALUE_ROOT = "./src/data/qvalue_files/"           # convert_textfile works one level up (qlearning.py:463-465)


def qvalue_path(number, directory=QVALUE_DIR):
    return os.path.join(directory, str(number) + ".txt")


def write_qvalues(path, items):
    """items: iterable of (state string, action char, float)."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w+") as f:
        for state, action, value in items:
            f.write(str(state) + "|" + action + "|" + repr(float(value)) + "\n")


def read_qvalues(path):
    """-> list of (state, action, float); later duplicates win, like the reference's dict."""
    out = []
    with open(path, "r") as f:
        for line in f.read().split("\n"):
            if line == "":
                continue
            state, action, value = line.split("|")[:3]
            out.append((state, action, float(value)))
    return out


def read_bad_states(path):
    with open(path, "r") as f:
        return [x for x in f.read().split("\n") if x != ""]


def strip_headings(state):
    for h in "NEWS":
        state = state.replace(h, "")
    return state


def expand_deadlock_state(state, station_boxes, no_of_cars, max_cars):
    """All strings save_deadlock_state derives from one detected state: the heading-stripped state,
    plus up to (max_cars - no_of_cars) rounds of appending a parked car `zBxB` for every station
    box that is not already a source, each in every ordering whose first entry has src != dst."""
    base = strip_headings(state)
    generated, expanded = [base], set()
    for _ in range(1, max_cars - no_of_cars + 1):
        fresh = []
        for s in generated:
            if s in expanded:
                continue
            expanded.add(s)
            sources = [part[:2] for part in s.split("z")]
            fresh += [s + "z" + b + "x" + b for b in station_boxes if b not in sources]
        generated += fresh
    out = set()
    for s in set(generated):
        parts = s.split("z")
        for perm in itertools.permutations(parts, len(parts)):
            src, dst = perm[0].split("x")
            if src != dst:
                out.add("z".join(perm))
    return out


def save_deadlock_state(state, station_boxes, bad_states_file, no_of_cars, max_cars):
    """reference src/utils/utils.py:159-199: union with the file's content, file rewritten."""
    with open(bad_states_file, "r") as f:
        stored = f.read().split()
    final = set(stored) | expand_deadlock_state(state, station_boxes, no_of_cars, max_cars)
    with open(bad_states_file, "w+") as f:
        for s in sorted(final):
            f.write(s + "\n")
    return final


# ---- block -> whole-floor renaming (reference src/algorithms/qlearning.py:457-597) -----------------
# A block is one 3x3 window of the 3x5 two-floor grid (README.md:6-7): its boxes 1..9 become boxes of
# floor 'A' (the reference's D) or 'B' (its U) of the whole grid -- the left windows keep their numbers,
# the right windows are shifted by two columns (6 boxes).
_BLOCK_FLOOR_SHIFT = {0: ("U", "B", 0), 1: ("D", "A", 0), 2: ("D", "A", 6), 3: ("U", "B", 6)}
_BOX_TERMINATORS = "xzNEWS|"  # what may follow a box name in a qvalue line (qlearning.py:535-541)


def block_box_map(block_number, reverse=False):
    """{block box: whole-floor box} of convert_textfile, or its inverse for reverse_convert_textfile."""
    try:
        local, whole, shift = _BLOCK_FLOOR_SHIFT[block_number]
    except KeyError:
        raise KeyError(block_number)  # the reference's box_maps_dict[block_number] raises the same
    pairs = [(local + str(k), whole + str(k + shift)) for k in range(1, 10)]
    return {b: a for a, b in pairs} if reverse else dict(pairs)


def rename_boxes(text, box_map):
    """The reference's line-by-line textual substitution: a box name is replaced wherever one of
    x z N E W S | follows it.  The file's lines are re-joined with a newline after EVERY element of
    text.split('\n'), so a file that ends in a newline gains one more (qlearning.py:533-543)."""
    out = []
    for line in text.split("\n"):
        for key, new in box_map.items():
            for t in _BOX_TERMINATORS:
                line = line.replace(key + t, new + t)
        out.append(line + "\n")
    return "".join(out)


def convert_textfile(input_textfile, output_textfile, block_number, directory=QVALUE_ROOT, reverse=False):
    """convert_textfile / reverse_convert_textfile: names are relative to ./src/data/qvalue_files/
    (NOT the layout-3 sub-directory the other methods use) and get '.txt' appended."""
    box_map = block_box_map(block_number, reverse)
    with open(os.path.join(directory, input_textfile + ".txt"), "r") as f:
        text = f.read()
    with open(os.path.join(directory, output_textfile + ".txt"), "w+") as f:
        f.write(rename_boxes(text, box_map))


def ascend_state(state):
    """array_to_state(ascend_array(state_to_array(state))) (reference src/env_gym.py:226-251,267-279):
    the primary car stays first, the others are sorted by (box number, [source, destination])."""
    parts = [p.split("x") for p in state.split("z")]
    rest = sorted(((int("".join(ch for ch in p[0] if ch.isdigit())), p) for p in parts[1:]))
    return "z".join("x".join(p) for p in [parts[0]] + [p for _, p in rest])
