#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.json.gz from the LIVE, UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
The outputs are committed; /root/reference is never read by tests, smoke() or bench.py.

What is recorded (all produced by reference code, none by this repo's implementation):
  tables_b{0,2}.json.gz   Configuration.action_dictionary, no_node_list, base legal sets
                          (Environment.get_legal_actions) and breadth_first_search lengths
                          (reference src/configuration.py:44-145, src/env_gym.py:114-130,443-483)
  traj_*.json.gz          round-robin rollouts under a recorded (heading, action) stream: per
                          agent-step state string, legal set, action, next state, reward
                          (reference src/algorithms/qlearning.py:211-233 loop shape,
                           src/env_gym.py:281-329,188-224,256-265,555-620,386-441)
  qrun_*.json.gz          RL.train_given_state / detect-style runs with the coin / choice /
                          heading decisions recorded and the final Counter (insertion order kept)
                          (reference src/algorithms/qlearning.py:74-116,187-239)
  qrand_*.json.gz         RL.train_given_state_random_order runs, the drawn car recorded as well
                          (reference src/algorithms/qlearning.py:241-295)
  known_answers.json.gz   the SURVEY 8c known-answer probes re-derived live.
  textfile_convert.json.gz  RL.convert_textfile / reverse_convert_textfile on qvalue files of all four
                          block numbers and RL.convert_textfile_in_ascending_order (file written and
                          the qvals_new it builds) (reference src/algorithms/qlearning.py:439-597)
"""
import gzip
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def dump(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".json.gz")
    raw = json.dumps(obj, separators=(",", ":"), sort_keys=True).encode()
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(raw)
    print("%-28s %8d B raw %8d B gz" % (name, len(raw), os.path.getsize(path)))


# ------------------------------------------------------------------------------------------
def gen_tables(ref, block):
    env = ref.make_env(block)
    cfg = env.env_conf
    nodes = list(cfg.action_dictionary.keys())
    dests = []
    for lst in cfg.destination_nodes_list:
        for d in lst:
            if d not in dests:
                dests.append(d)
    bfs = {}
    for d in dests:
        row = {}
        for n in nodes:
            acts, _ = env.breadth_first_search([[n, d]])
            row[n] = len(acts)
        bfs[d] = row
    return {
        "block": block,
        "node_order": nodes,
        "action_dictionary": {n: cfg.action_dictionary[n] for n in nodes},
        "no_node_list": list(cfg.no_node_list),
        "base_legal": {n: "".join(env.get_legal_actions(n)) for n in nodes},
        "bfs_len": bfs,
        "destination_nodes_list": cfg.destination_nodes_list,
        "station_box_list": cfg.station_box_list,
        "no_box_list": cfg.no_box_list,
        # the layout dict itself (reference src/layout.py:2-36,244-275); no_node_list is grown in
        # place by Configuration.__init__ (src/configuration.py:52,138-141) so it is reset here
        "layout": dict(ref.layout.layout_3[block], no_node_list=[]),
    }


# ------------------------------------------------------------------------------------------
def gen_traj(ref, name, block, boxes, idx, choose, deadlock_id, reward, bad, episodes, cap, seed,
             use_step=False):
    """Round-robin rollout, uniform-random legal actions (the reference's eps=1 behaviour)."""
    env = ref.make_env(block, bad_states=bad)
    assert env.give_bad_state_reward == bool(bad)
    rng = random.Random(seed)
    random.seed(seed + 1000)  # the env's heading shuffle draws from the global generator
    eps_out = []
    total = 0
    for ep in range(episodes):
        with ref_harness.quiet():
            env.reset()
        env.initialize_cars(len(boxes))
        env.initialize(idx, boxes, choose)
        env.select_reward_function(reward)
        steps = []
        capped = False
        done = False
        while not done and not capped:
            for num in range(len(boxes)):
                if env.check_car_training(num):
                    continue
                state = env.encodestate(num, deadlock_id)
                legal = env.getlegalactions(env.state_to_array(state))
                if not legal:
                    raise RuntimeError("no legal action in %s" % state)
                action = rng.choice(legal)
                if use_step:
                    _, _, nxt, r, _ = env.step(state, action, num)
                else:
                    nxt = env.get_successor_state(state, action)
                    r = env.reward_function(state, nxt, num, action)
                    arr = env.state_to_array(nxt)
                    env.update_car(num, arr[0][0], arr[0][1], 1)
                steps.append([num, state, "".join(legal), action, nxt, r])
                if len(steps) >= cap:
                    capped = True
                    break
            done = env.is_game_ended()
        total += len(steps)
        eps_out.append({"headings": list(env.given_nodes), "steps": steps, "done": bool(done)})
    dump("traj_" + name, {
        "block": block, "boxes": boxes, "idx": idx, "choose": choose, "deadlock_id": deadlock_id,
        "reward": reward, "bad_states": bool(bad), "cap": cap, "episodes": eps_out,
        "n_steps": total})


# ------------------------------------------------------------------------------------------
class _Abort(Exception):
    pass


def gen_qrun(ref, name, block, boxes, idx, choose, bad, episodes, max_steps, seed):
    """Real RL.train_given_state with every random decision recorded."""
    rl, env = ref.make_rl(block, bad_states=bad)
    Q = ref.qlearning
    random.seed(seed)
    log = {"episodes": []}
    cur = {}
    state = {"n": 0}

    real_flip = ref.utils.flipcoin
    real_choice = random.choice
    real_init = env.initialize
    real_update = rl.update

    def flip(p):
        r = real_flip(p)
        cur["pending_coin"] = bool(r)
        return r

    class RandomProxy:
        def __getattr__(self, k):
            return getattr(random, k)

        @staticmethod
        def choice(seq):
            c = real_choice(seq)
            cur["pending_choice"] = c
            return c

    def init(a, b, c):
        real_init(a, b, c)
        cur.clear()
        ep = {"headings": list(env.given_nodes), "eps": rl.epsilon, "steps": []}
        log["episodes"].append(ep)
        cur["ep"] = ep

    def update(s, a, ns, r):
        real_update(s, a, ns, r)
        coin = cur.pop("pending_coin")
        ch = cur.pop("pending_choice", None)
        # [state, explore?, action, next_state, reward, Q(s,a) after update]
        cur["ep"]["steps"].append([s, int(coin), a, ns, r, repr(rl.qvals[(s, a)])])
        if coin:
            assert ch == a
        state["n"] += 1
        if state["n"] >= max_steps:
            raise _Abort()

    Q.flipcoin = flip
    Q.random = RandomProxy()
    env.initialize = init
    rl.update = update
    aborted = False
    try:
        with ref_harness.quiet():
            rl.train_given_state(episodes, "0", "0", boxes, len(boxes), idx, choose)
    except _Abort:
        aborted = True
    finally:
        Q.flipcoin = real_flip
        Q.random = random
        env.initialize = real_init
    qv = [[s, a, repr(v)] for (s, a), v in rl.qvals.items()]
    dump("qrun_" + name, {
        "block": block, "boxes": boxes, "idx": idx, "choose": choose, "bad_states": bool(bad),
        "episodes_total": episodes, "aborted": aborted, "n_steps": state["n"],
        "alpha": 0.9, "discount": 0.15, "log": log["episodes"], "qvals": qv})


def gen_qrand(ref, name, block, boxes, idx, choose, bad, episodes, max_steps, seed):
    """Real RL.train_given_state_random_order (reference src/algorithms/qlearning.py:241-295) with every
    random decision recorded: the car drawn by random.randint before each agent-step (draws that hit a
    finished car are logged too -- they consume the stream and nothing else), coin, choice, headings."""
    rl, env = ref.make_rl(block, bad_states=bad)
    Q = ref.qlearning
    random.seed(seed)
    log = {"episodes": []}
    cur = {}
    state = {"n": 0}
    real_flip = ref.utils.flipcoin
    real_choice = random.choice
    real_randint = random.randint
    real_init = env.initialize
    real_update = rl.update

    def flip(p):
        r = real_flip(p)
        cur["pending_coin"] = bool(r)
        return r

    class RandomProxy:
        def __getattr__(self, k):
            return getattr(random, k)

        @staticmethod
        def choice(seq):
            c = real_choice(seq)
            cur["pending_choice"] = c
            return c

        @staticmethod
        def randint(a, b):
            c = real_randint(a, b)
            cur["ep"]["draws"].append(c)
            cur["pending_car"] = c
            return c

    def init(a, b, c):
        real_init(a, b, c)
        cur.clear()
        ep = {"headings": list(env.given_nodes), "eps": rl.epsilon, "steps": [], "draws": []}
        log["episodes"].append(ep)
        cur["ep"] = ep

    def update(s, a, ns, r):
        real_update(s, a, ns, r)
        coin = cur.pop("pending_coin")
        ch = cur.pop("pending_choice", None)
        # [car, state, explore?, action, next_state, reward, Q(s,a) after update]
        cur["ep"]["steps"].append([cur["pending_car"], s, int(coin), a, ns, r, repr(rl.qvals[(s, a)])])
        if coin:
            assert ch == a
        state["n"] += 1
        if state["n"] >= max_steps:
            raise _Abort()

    Q.flipcoin = flip
    Q.random = RandomProxy()
    env.initialize = init
    rl.update = update
    aborted = False
    result = None
    try:
        with ref_harness.quiet():
            result = rl.train_given_state_random_order(episodes, "0", "0", boxes, len(boxes), idx, choose)
    except _Abort:
        aborted = True
    finally:
        Q.flipcoin = real_flip
        Q.random = random
        env.initialize = real_init
    qv = [[s, a, repr(v)] for (s, a), v in rl.qvals.items()]
    dump("qrand_" + name, {
        "block": block, "boxes": boxes, "idx": idx, "choose": choose, "bad_states": bool(bad),
        "episodes_total": episodes, "aborted": aborted, "n_steps": state["n"], "seed": seed, "result": result,
        "alpha": 0.9, "discount": 0.15, "log": log["episodes"], "qvals": qv})


# ------------------------------------------------------------------------------------------
def gen_known_answers(ref):
    out = {}
    env = ref.make_env(2)
    # example transition and goal step, block 2
    env.initialize_cars(2)
    env.select_reward_function("q_reward")
    env.cars[0].source_node, env.cars[0].destination_index, env.cars[0].destination_choose_array = "D2N", 0, [0, 1, 0]
    env.cars[1].source_node, env.cars[1].destination_index, env.cars[1].destination_choose_array = "D3E", 0, [0, 1, 0]
    s = env.encodestate(0, 1)
    ns = env.get_successor_state(s, "F")
    out["transition"] = {"state": s, "legal": "".join(env.getlegalactions(env.state_to_array(s))),
                         "action": "F", "next": ns, "reward": env.reward_function(s, ns, 0, "F")}
    env.cars[0].source_node = "D6S"
    env.cars[1].source_node = "D4E"
    s = env.encodestate(0, 1)
    ns = env.get_successor_state(s, "T")
    out["goal"] = {"state": s, "action": "T", "next": ns, "reward": env.reward_function(s, ns, 0, "T")}
    # ghost known-answer: car 1 finishes at D8E, car 0 at D5E heading for D6N
    env = ref.make_env(2)
    env.initialize_cars(2)
    env.cars[0].source_node, env.cars[0].destination_index, env.cars[0].destination_choose_array = "D5E", 0, [0, 1, 0]
    env.cars[1].source_node, env.cars[1].destination_index, env.cars[1].destination_choose_array = "D8N", 1, [0, 1, 0]
    env.update_car(1, "D8E", "D8E", 1)
    ghost = []
    for _ in range(8):
        s = env.encodestate(0, 1)
        ghost.append([s, "".join(env.getlegalactions(env.state_to_array(s))), env.cars[1].training_done_count])
    out["ghost"] = {"encodes": ghost, "parked": env.cars[1].source_node,
                    "parked_di": env.cars[1].destination_index, "mode0": env.encodestate(0, 0)}
    arr = env.state_to_array("D5ExD6NzD8xD8zD1xD6")
    out["ascend"] = env.array_to_state(env.ascend_array(arr))
    # epsilon schedule, model 3 (reference src/utils/utils.py:33-52)
    sched = {}
    for E in (100, 60, 5000, 7):
        sched[str(E)] = [ref.utils.give_action_choose_probability(E, e, 3) for e in range(E)]
    out["eps_model3"] = {k: v if len(v) <= 100 else
                         [[i, v[i]] for i in range(len(v)) if i == 0 or v[i] != v[i - 1]]
                         for k, v in sched.items()}
    # deadlock-state expansion (reference src/utils/utils.py:141-199) on a scratch file
    path = os.path.join(ref.scratch, "dl_scratch.txt")
    exp = {}
    for st, n in (("D5NxD6NzD6xD8", 2), ("D5ExD6NzD6xD8zD4xD4", 3)):
        open(path, "w").close()
        ref.utils.save_deadlock_state(st, ["D6", "D4", "D8", "D7"], path, n, 4)
        exp[st] = sorted(open(path).read().split())
    out["deadlock_expand_max4"] = exp
    dump("known_answers", out)


def gen_convert(ref):
    """RL.convert (the 80-bit DQN featurizer) and give_mask of the noisy double DQN
    (reference src/algorithms/dqn_open_source.py:200-209,245-313), called unbound on recorded states."""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        import src.algorithms.dqn_open_source as D
    out = {}
    for name in ("traj_b2_4car_bad", "traj_b2_5car", "traj_b0_3car_bad", "traj_b2_2car_dqn_step", "traj_b0_1car"):
        with gzip.open(os.path.join(OUT, name + ".json.gz")) as f:
            g = json.load(f)
        states = []
        for ep in g["episodes"]:
            for st in ep["steps"]:
                for s in (st[1], st[4]):
                    if s not in states:
                        states.append(s)
        rng = random.Random(7)
        rng.shuffle(states)
        states = states[:120]
        feats = D.RL.convert(None, states, 5)
        env = ref.make_env(g["block"], bad_states=False)
        masks = []
        for s in states:
            valid = env.getlegalactions(env.state_to_array(s))
            masks.append([i for i, a in enumerate(['S', 'L', 'R', 'F', 'T']) if a in valid])
        out[name] = {"block": g["block"], "boxes": g["boxes"], "idx": g["idx"], "choose": g["choose"],
                     "states": states, "features": ["".join(str(int(x)) for x in row) for row in feats],
                     "masks": masks}
    dump("dqn_convert", out)


def gen_textfile_conversions(ref):
    """The three file-conversion methods of RL (reference src/algorithms/qlearning.py:439-597), run
    unmodified in the harness' scratch cwd on qvalue files written in the reference's own format."""
    root = os.path.join(ref.scratch, "src/data/qvalue_files")
    rl, _env = ref.make_rl(2)
    out = {"cases": []}

    def lines_of(golden, limit):
        with gzip.open(os.path.join(OUT, golden + ".json.gz")) as f:
            q = json.load(f)["qvals"]
        return "".join("%s|%s|%s\n" % (s, a, v) for s, a, v in q[:limit])

    d_text = lines_of("qrun_b2_3car_bad", 70) + lines_of("qrun_b2_2car", 50)
    u_text = lines_of("qrun_b0_1car", 60) + "U4NxU2WzU8xU1zU5xU4|F|1.5\nU9WxU1N|T|-2.25\n"
    odd = "D1NxD6N|S|0.5\n\nD10xD2|F|1.0\nXD3xD4zD5S|L|2.0"   # blank line, no trailing newline, near misses
    for block, text in ((0, u_text), (1, d_text), (2, d_text), (3, u_text), (2, odd), (1, "")):
        with open(os.path.join(root, "cv_in.txt"), "w") as f:
            f.write(text)
        with ref_harness.quiet():
            rl.convert_textfile("cv_in", "cv_out", block)
            rl.reverse_convert_textfile("cv_out", "cv_back", block)
        out["cases"].append({"block": block, "input": text,
                             "converted": open(os.path.join(root, "cv_out.txt")).read(),
                             "reversed": open(os.path.join(root, "cv_back.txt")).read()})
    # convert_textfile_in_ascending_order: others of some keys deliberately out of order, one duplicate key
    asc_in = ("D5ExD6NzD8xD8zD1xD6|F|1.25\nD5ExD6NzD1xD6zD8xD8|F|-3.5\nD2NxD6NzD3xD6|S|0.1\n"
              "D2NxD6NzD7xD4zD3xD6zD5xD8|L|7.0\nD2NxD6NzD3xD8zD3xD6|R|-0.75\nD2NxD6NzD3xD6|S|0.30000000000000004\n")
    with open(os.path.join(root, "layout-3", "asc_in.txt"), "w") as f:
        f.write(asc_in)
    with ref_harness.quiet():
        rl.convert_textfile_in_ascending_order("asc_in", "asc_out")
    out["ascending"] = {"input": asc_in, "output": open(os.path.join(root, "layout-3", "asc_out.txt")).read(),
                        "qvals_new": [[s, a, repr(v)] for (s, a), v in rl.qvals_new.items()]}
    dump("textfile_convert", out)


def gen_all_qrand(ref):
    R = gen_qrand
    R(ref, "b2_2car", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], False, 60, 40000, 6)
    R(ref, "b2_3car_bad", 2, ["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]], True, 40, 30000, 7)


def main():
    ref = ref_harness.load()
    if "--only-qrand" in sys.argv:  # the other fixtures are unchanged
        gen_all_qrand(ref)
        ref.close()
        return
    if "--only-textfiles" in sys.argv:
        gen_textfile_conversions(ref)
        ref.close()
        return
    for b in (0, 2):
        dump("tables_b%d" % b, gen_tables(ref, b))
    gen_known_answers(ref)
    # the two bad-states data files, verbatim (reference src/data/bad_states_files/layout-3/)
    bdir = os.path.join(ref_harness.REFERENCE_ROOT, "src/data/bad_states_files/layout-3")
    dump("bad_states", {k: open(os.path.join(bdir, k + ".txt")).read() for k in ("up_left", "down_right")})
    T = gen_traj
    T(ref, "b0_1car", 0, ["U4"], [1], [[0, 0, 0]], 1, "q_reward", False, 40, 4000, 11)
    T(ref, "b0_1car_2dest", 0, ["U8"], [0], [[0, 0, 0]], 1, "q_reward", False, 30, 4000, 12)
    T(ref, "b0_3car_bad", 0, ["U4", "U8", "U7"], [0, 1, 1], [[0, 0, 0]] * 3, 1, "q_reward", True, 12, 1500, 13)
    T(ref, "b0_4car", 0, ["U4", "U8", "U5", "U7"], [0, 1, 0, 0], [[0, 0, 0]] * 4, 1, "q_reward", False, 8, 1500, 14)
    T(ref, "b2_2car", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], 1, "q_reward", False, 25, 3000, 15)
    T(ref, "b2_2car_bad", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], 1, "q_reward", True, 25, 3000, 16)
    T(ref, "b2_3car_mode0_bad", 2, ["D5", "D4", "D8"], [0, 0, 0], [[2, 1, 0], [1, 0, 0], [2, 1, 0]], 0,
      "q_reward", True, 10, 1500, 17)
    T(ref, "b2_4car_bad", 2, ["D2", "D6", "D4", "D8"], [0, 0, 0, 0],
      [[0, 1, 0], [0, 1, 0], [1, 0, 0], [2, 1, 0]], 1, "q_reward", True, 8, 1500, 18)
    T(ref, "b2_2car_dqn_step", 2, ["D3", "D5"], [1, 1], [[0, 2, 0], [0, 1, 0]], 1, "dqn_reward", False, 15,
      2000, 19, use_step=True)
    T(ref, "b2_5car", 2, ["D1", "D2", "D3", "D5", "D7"], [0, 1, 0, 1, 0],
      [[0, 1, 0], [2, 1, 0], [1, 0, 0], [2, 2, 0], [3, 4, 0]], 1, "q_reward", True, 4, 1500, 20)
    G = gen_qrun
    G(ref, "b0_1car", 0, ["U4"], [1], [[0, 0, 0]], False, 200, 10 ** 9, 3)
    G(ref, "b2_2car", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], False, 60, 40000, 3)
    G(ref, "b2_2car_bad", 2, ["D3", "D2"], [0, 0], [[0, 1, 0], [0, 1, 0]], True, 60, 40000, 4)
    G(ref, "b2_3car_bad", 2, ["D6", "D4", "D8"], [0, 0, 0], [[0, 1, 0], [1, 0, 0], [2, 1, 0]], True, 40, 30000, 5)
    gen_all_qrand(ref)
    gen_convert(ref)
    gen_textfile_conversions(ref)
    ref.close()


if __name__ == "__main__":
    main()
