"""On-disk formats of the hot path, host side.

* qvalue files: one `state|action|repr(float)` line per key
  (reference src/algorithms/qlearning.py:159-185); name '0' means "no file".
* bad-states files: one heading-stripped state string per line, read by the env
  (reference src/env_gym.py:71-83) and grown by save_deadlock_state
  (reference src/utils/utils.py:141-199).
"""
import itertools
import os

QVALUE_DIR = "./src/data/qvalue_files/layout-3/"  # hard-coded in the reference (qlearning.py:160,170)


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
