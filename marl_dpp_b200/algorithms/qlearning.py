"""Tabular Q-learning plugin with the reference's algorithm contract
(reference src/algorithms/qlearning.py:13-437; constructed by app.py:91-96 as
`RL(**settings.params, sess=, env_config=)` and driven through train_given_state /
detect_deadlocks_v0 / validate_given_state, app.py:42-72).

The per-episode Python loop of the reference (qlearning.py:187-239) is replaced by the fused
CUDA learner over `n_envs` lockstep copies of the initial state (marl_dpp_b200.batched); with
n_envs = 1 it performs the reference's update sequence.  File arguments keep the reference's
meaning: strings, '0' = none, files under ./src/data/qvalue_files/layout-3/.
"""
import os

import numpy as np
import torch

from .. import formats
from ..batched import BatchedEnv, BatchedQLearner
from ..configuration import Configuration
import random

from ..tables import (ACTIONS5, HEADS, LEARN_DETECT, LEARN_FROZEN, LEARN_MODE0, LEARN_OBSERVED, LEARN_RANDOM_ORDER,
                      Scenario, epsilon_schedule)


class RL:
    def __init__(self, input_dim=None, output_dim=None, cars=None, sess=None, env_config=None,
                 learning_rate=0.0001, batch_size=32, n_envs=1, device=None, step_cap=10000, seed=3,
                 bad_states=None, qvalue_dir=formats.QVALUE_DIR, reference_rng=False):
        """input_dim / output_dim / cars / sess / learning_rate / batch_size are accepted and ignored,
        exactly like the reference (qlearning.py:14-15).  Extra keywords: n_envs lockstep copies,
        step_cap agent-steps per episode before a forced reset (the reference's train_given_state has
        none and can spin forever at epsilon 0 -- SURVEY fact 5; 0 restores that), bad_states: None =
        try the layout's file like the reference (env_gym.py:71-83), or the file's text.
        reference_rng (n_envs == 1 only): draw headings, coins and exploratory actions from Python's
        global `random` in the reference's exact call order (random.shuffle per car at initialize,
        one random.random() per getaction, random.choice(legal) when exploring), so a run seeded like
        the reference (`random.seed(settings.SEED)`) visits the same states and learns the same table."""
        self.env_config = Configuration.coerce(env_config)  # ours, a layout dict, or the reference's object
        self.alpha = 0.9
        self.discount = 0.15
        self.probability_model = 3
        self.epsilon = 0
        self.n_envs = int(n_envs)
        self.device = device
        self.step_cap = int(step_cap)
        self.seed = int(seed)
        self.qvalue_dir = qvalue_dir
        self.reference_rng = bool(reference_rng)
        if self.reference_rng and self.n_envs != 1:
            raise ValueError("reference_rng replays the reference's single-env random stream: n_envs must be 1")
        self.training_complete_destination_index = self.env_config.training_complete_destination_index
        if bad_states is None:
            try:
                with open(self.env_config.environment_bad_states_file, "r") as f:
                    bad_states = f.read()
            except OSError:
                bad_states = None  # "as run" from the repo root: the file is not found
        self.bad_states_text = bad_states
        self.qvals = {}          # (state string, action char) -> float, the reference's Counter
        self.last_stats = None
        self.learner = None

    # ---- Q values <-> files (reference :159-185) -------------------------------------------------
    def load_qvalues(self, inputfile_number):
        self.qvals = {}
        if str(inputfile_number) == '0':
            return None
        for state, action, value in formats.read_qvalues(formats.qvalue_path(inputfile_number, self.qvalue_dir)):
            self.qvals[(state, action)] = value

    def save_qvalues(self, outputfile_number):
        if str(outputfile_number) == '0':
            return None
        self._pull()
        formats.write_qvalues(formats.qvalue_path(outputfile_number, self.qvalue_dir),
                              [(s, a, v) for (s, a), v in self.qvals.items()])

    # ---- file conversions (reference :439-597) ----------------------------------------------------------
    def convert_textfile_in_ascending_order(self, inputfile_number, outputfile_number):
        """Re-key the loaded table with every state's others in ascend_array order into `qvals_new`,
        then save_qvalues(outputfile_number).  Like the reference, save_qvalues writes `self.qvals`
        -- the table as loaded -- so the file on disk is a re-serialisation of the input and the
        ascended table only lives in `qvals_new` (reference :446-455)."""
        self.learner = None          # the file, not a device table from an earlier run, is the source
        self.load_qvalues(inputfile_number)
        self.qvals_new = {}
        for (state, action), value in self.qvals.items():
            self.qvals_new[(formats.ascend_state(state), action)] = value
        self.save_qvalues(outputfile_number)

    def convert_textfile(self, input_textfile, output_textfile, block_number):
        """Rename a block's boxes to whole-floor ("MSV") boxes in a qvalue file (reference :457-542)."""
        formats.convert_textfile(input_textfile, output_textfile, block_number,
                                 directory=self._qvalue_root())

    def reverse_convert_textfile(self, input_textfile, output_textfile, block_number):
        """The inverse renaming (reference :544-597)."""
        formats.convert_textfile(input_textfile, output_textfile, block_number,
                                 directory=self._qvalue_root(), reverse=True)

    def _qvalue_root(self):
        # the reference hard-codes ./src/data/qvalue_files/ here and .../layout-3/ for load/save
        if self.qvalue_dir == formats.QVALUE_DIR:
            return formats.QVALUE_ROOT
        return self.qvalue_dir

    def _push(self, learner):
        """Loaded Counter -> dense device table (keys outside this scenario's state space are kept
        on the host and written back unchanged)."""
        sc = learner.env.scenario
        q = torch.zeros((sc.n_states, 8), dtype=torch.float32)
        vis = torch.zeros(sc.n_states, dtype=torch.int32)
        self._foreign = {}
        for (state, action), value in self.qvals.items():
            try:
                idx, a = sc.encode(state), ACTIONS5.index(action)
            except ValueError:
                self._foreign[(state, action)] = value
                continue
            q[idx, a] = value
            vis[idx] |= 1 << a
        learner.q[:, :5] = q[:, :5].to(learner.q.device)   # slots 5..7 of a row belong to the library
        learner.visited.copy_(vis)

    def _pull(self):
        if self.learner is None:
            return
        self.qvals = dict(getattr(self, "_foreign", {}))
        for state, action, value in self.learner.qvalue_items():
            self.qvals[(state, action)] = value

    # ---- Counter-style accessors (reference :34-72) -----------------------------------------------
    def getqvalue(self, state, action):
        self._pull()
        return self.qvals.get((state, action), 0.0)

    def computeactionfromqvalues(self, state):
        """First maximum over the legal actions in S,L,R,F,T order (reference :60-72, :659-667)."""
        env = self.learner.env
        sc = env.scenario
        idx = sc.encode(state)
        legal = int(env.query(0, [idx])["legal"][0])
        if legal == 0:
            return None
        row = self.learner.q[idx].tolist()
        best = None
        for a in range(5):
            if (legal >> a) & 1 and (best is None or row[a] > row[best]):
                best = a
        return ACTIONS5[best]

    getpolicy = computeactionfromqvalues

    def _legal_of(self, state):
        if self.learner is None:
            raise RuntimeError("no scenario yet: call prepare(...) or one of the train/validate methods first")
        env = self.learner.env
        idx = env.scenario.encode(state)
        return idx, int(env.query(0, [idx])["legal"][0])

    def computevaluefromqvalues(self, state):
        """max over the legal actions of Q(state, .), 0.0 without legal actions (reference :45-58)."""
        idx, legal = self._legal_of(state)
        row = self.learner.q[idx].tolist()
        vals = [row[a] for a in range(5) if (legal >> a) & 1]
        return max(vals) if vals else 0.0

    getvalue = computevaluefromqvalues

    def getaction(self, state, epsilon):
        """One random.random() draw always; below epsilon -> random.choice(legal), else the greedy action;
        None without legal actions (reference :74-105, flipcoin src/utils/utils.py:255-265)."""
        idx, legal = self._legal_of(state)
        legal_actions = [ACTIONS5[a] for a in range(5) if (legal >> a) & 1]
        if not legal_actions:
            return None
        if random.random() < epsilon:
            return random.choice(legal_actions)
        return self.computeactionfromqvalues(state)

    def update(self, state, action, nextState, reward):
        """Q[s,a] += alpha * (r + discount * max_a' Q[s',a'] - Q[s,a]) on the device table (reference :107-116),
        float32 with the kernels' operation order."""
        idx, _ = self._legal_of(state)
        a = ACTIONS5.index(action)
        q = self.learner.q
        f32 = torch.float32
        nv = torch.tensor(self.computevaluefromqvalues(nextState), dtype=f32)
        qa = q[idx, a].cpu()
        target = torch.tensor(float(reward), dtype=f32) + torch.tensor(self.discount, dtype=f32) * nv
        q[idx, a] = (qa + torch.tensor(self.alpha, dtype=f32) * (target - qa)).to(q.device)
        self.learner.visited[idx] |= 1 << a

    def prepare(self, input_car_states, no_of_cars=1, destination_index_array=[0, 0, 0, 0],
                destination_choose_arrays=[0, 0, 0], episodes=1):
        """Build the device table for a scenario without training (the reference's methods above lean on
        the module-global env having been initialised, qlearning.py:17-18,197-205)."""
        return self._make(episodes, input_car_states, no_of_cars, destination_index_array, destination_choose_arrays)

    # ---- training (reference :187-239) -----------------------------------------------------------
    def _make(self, episodes, boxes, n, idx, choose, flags=0, model=None):
        sc = Scenario(self.env_config.layout, list(boxes[:n]), list(idx[:n]), [list(c) for c in choose[:n]],
                      bad_states_text=self.bad_states_text, step_cap=self.step_cap)
        benv = BatchedEnv(sc, self.n_envs, device=self.device, seed=self.seed)
        self.learner = BatchedQLearner(benv, episodes=episodes, alpha=self.alpha, discount=self.discount,
                                       probability_model=self.probability_model if model is None else model,
                                       seed=self.seed, flags=flags)
        self._push(self.learner)
        return self.learner

    def _run(self, learner, episodes, on_episode=None, max_rounds=None):
        """Rounds until every env used its episode budget.  Returns rounds run."""
        chunk = 1 if self.n_envs == 1 else 256
        rounds = 0
        while True:
            learner.train_rounds(chunk)
            rounds += chunk
            # a counter fetch lets the library tighten its bound on the episode counters (it grows by one per
            # round otherwise) and so keep the cheap pure-exploration kernels for the whole epsilon = 1 segment
            learner.env.stats()
            ep = learner.env.cars()["episode"]
            lo = int(ep.min())
            if on_episode is not None:
                on_episode(lo)
            if lo >= episodes or (max_rounds is not None and rounds >= max_rounds):
                return rounds

    def _train_reference_rng(self, episodes, outputfile_number, boxes, n, idx, choose):
        """The reference's loop (qlearning.py:191-238) on the host, one env, every decision drawn from
        Python's `random` in the reference's order; the env step, the greedy argmax and the TD update
        run on the GPU."""
        learner = self._make(episodes, boxes, n, idx, choose, flags=LEARN_OBSERVED)
        env = learner.env
        th, eps_values = epsilon_schedule(episodes, self.probability_model)
        peek = torch.tensor([0xFE], dtype=torch.uint8)
        self.trace = []
        for episode in range(episodes):
            if episode == 30:
                self.save_qvalues(outputfile_number)
            self.epsilon = next((v for t, v in zip(th, eps_values) if episode < t), eps_values[5])
            heads = []
            for num in range(n):                        # env.initialize: random.shuffle(nodes)[0] per car
                nodes = list("NESW")
                random.shuffle(nodes)
                heads.append(HEADS.index(nodes[0]))
            env.reset(headings=torch.tensor([heads], dtype=torch.uint8), zero_counters=False)
            done = [False] * n
            steps = 0
            ep0 = int(env.cars()["episode"][0])
            while not all(done):
                for num in range(n):
                    if done[num]:
                        continue
                    o = env.step_car(num, peek)          # encodestate(num, 1) + getlegalactions
                    legal = int(o["legal"][0])
                    legal_actions = [a for a in range(5) if (legal >> a) & 1]
                    forced = 0xFF
                    if legal_actions:
                        if random.random() < self.epsilon:      # flipcoin
                            forced = random.choice(legal_actions)
                        else:
                            forced = 0xFE                       # computeactionfromqvalues on the device
                    out = learner.substep(num, forced=torch.tensor([forced], dtype=torch.uint8), outputs=True)
                    learner.finalize()
                    self.trace.append((int(out["state"][0]), int(out["action"][0])))
                    steps += 1
                learner.round += 1
                cars = env.cars()
                # the last car slot's launch starts the next episode by itself when every car is done
                if int(cars["episode"][0]) != ep0:
                    break
                done = [bool(x) for x in cars["done"][0].tolist()]
                if self.step_cap and steps >= self.step_cap:
                    break
        return learner

    def train_given_state(self, episodes, inputfile_number, outputfile_number, input_car_states=None, no_of_cars=1,
                          destination_index_array=[0, 0, 0, 0], destination_choose_arrays=[0, 0, 0]):
        self.load_qvalues(inputfile_number)
        if self.reference_rng:
            learner = self._train_reference_rng(episodes, outputfile_number, input_car_states, no_of_cars,
                                                destination_index_array, destination_choose_arrays)
            self.last_stats = learner.env.stats()
            self.save_qvalues(outputfile_number)
            self._pull()
            return
        learner = self._make(episodes, input_car_states, no_of_cars, destination_index_array,
                             destination_choose_arrays)
        saved = []

        def snapshot(min_episode):  # `if episode == 30: self.save_qvalues(outputfile_number)` (:194-195)
            if not saved and min_episode >= 30:
                saved.append(True)
                self.save_qvalues(outputfile_number)
        self._run(learner, episodes, on_episode=snapshot)
        self.last_stats = learner.env.stats()
        self.save_qvalues(outputfile_number)
        self._pull()

    # ---- training, random car order (reference :241-295) ---------------------------------------------
    def _train_random_order_reference_rng(self, episodes, outputfile_number, boxes, n, idx, choose):
        """The reference's loop (qlearning.py:245-293) on the host, one env: random.randint picks the car
        (a finished car consumes the draw and nothing else), then the decisions of getaction in the
        reference's order; env step, greedy argmax and TD update on the GPU.  Returns (learner, ok): ok is
        False when an episode ran past 10 000 actions (the reference's `return 0`, :274-277)."""
        learner = self._make(episodes, boxes, n, idx, choose, flags=LEARN_OBSERVED | LEARN_RANDOM_ORDER)
        env = learner.env
        th, eps_values = epsilon_schedule(episodes, self.probability_model)
        peek = torch.tensor([0xFE], dtype=torch.uint8)
        self.trace = []
        for episode in range(episodes):
            if episode == 30:
                self.save_qvalues(outputfile_number)
            self.epsilon = next((v for t, v in zip(th, eps_values) if episode < t), eps_values[5])
            heads = []
            for num in range(n):
                nodes = list("NESW")
                random.shuffle(nodes)
                heads.append(HEADS.index(nodes[0]))
            env.reset(headings=torch.tensor([heads], dtype=torch.uint8), zero_counters=False)
            done = [False] * n
            actions = 0
            ep0 = int(env.cars()["episode"][0])
            while True:
                num = random.randint(0, n - 1)
                if done[num]:
                    continue
                o = env.step_car(num, peek)              # encodestate(num, 1) + getlegalactions
                legal = int(o["legal"][0])
                legal_actions = [a for a in range(5) if (legal >> a) & 1]
                low = 0xF
                if legal_actions:
                    if random.random() < self.epsilon:
                        low = random.choice(legal_actions)
                    else:
                        low = 0xE
                actions += 1
                if actions > 10000:
                    return learner, False
                out = learner.substep(0, forced=torch.tensor([(num << 4) | low], dtype=torch.uint8), outputs=True)
                learner.finalize()
                learner.round += 1
                self.trace.append((num, int(out["state"][0]), int(out["action"][0])))
                cars = env.cars()
                if int(cars["episode"][0]) != ep0:       # every car finished: the launch started the next episode
                    break
                done = [bool(x) for x in cars["done"][0].tolist()]
        return learner, True

    def train_given_state_random_order(self, episodes, inputfile_number, outputfile_number, input_car_states=None,
                                       no_of_cars=1, destination_index_array=[0, 0, 0, 0],
                                       destination_choose_arrays=[0, 0, 0]):
        """Like train_given_state, but every agent-step is taken by a randomly drawn car and an episode
        that runs past 10 000 actions aborts the training with 0 (a state without a path, :274-277).
        Batched (n_envs > 1): every env draws its own car in every sub-step (MDPP_LEARN_RANDOM_ORDER); the
        cap is the learner's step_cap, and the call returns 0 after the run if any episode hit it."""
        self.load_qvalues(inputfile_number)
        if self.reference_rng:
            learner, ok = self._train_random_order_reference_rng(episodes, outputfile_number, input_car_states,
                                                                 no_of_cars, destination_index_array,
                                                                 destination_choose_arrays)
            self.last_stats = learner.env.stats()
            self._pull()
            if not ok:
                print('\nSkipping! The state cannot be trained as it has no paths possible.')
                return 0
            self.save_qvalues(outputfile_number)
            return
        cap = self.step_cap
        self.step_cap = 10000 if cap == 0 else min(cap, 10000)  # this variant always has the cap (:274)
        try:
            learner = self._make(episodes, input_car_states, no_of_cars, destination_index_array,
                                 destination_choose_arrays, flags=LEARN_RANDOM_ORDER)
        finally:
            self.step_cap = cap
        saved = []

        def snapshot(min_episode):
            if not saved and min_episode >= 30:
                saved.append(True)
                self.save_qvalues(outputfile_number)
        self._run(learner, episodes, on_episode=snapshot)
        self.last_stats = learner.env.stats()
        self._pull()
        if self.last_stats["episodes_capped"]:
            print('\nSkipping! The state cannot be trained as it has no paths possible.')
            return 0
        self.save_qvalues(outputfile_number)

    # ---- deadlock detection (reference :297-382) -----------------------------------------------------
    def detect_deadlocks_v0(self, episodes=40, inputfile_number=0, outputfile_number=0, input_car_states=None,
                            no_of_cars=1, destination_index_array=[0, 0, 0, 0], destination_choose_arrays=[0, 0, 0],
                            max_cars=6):
        """Stall heuristic: Q-learning episodes with encodestate(., 0); in the epsilon = 0 episodes an
        env whose greedy policy emits 3*n_cars consecutive 'S' flags that state and stops.
        Return value as in the reference (:350-382): the EPISODE at which a not-yet-known state was
        flagged -- the state is expanded and written to the bad-states file (utils.py:159-199) and
        the Q values are NOT saved (the reference returns before save_qvalues); False when a
        flagged state is already in the file (the reference's `skip`); else the last episode index,
        after save_qvalues.  Batched: every env runs the heuristic; the smallest flagging episode
        over the envs is returned and every distinct new state is written."""
        self.env_config.no_of_cars = no_of_cars
        self.load_qvalues(inputfile_number)
        learner = self._make(episodes, input_car_states, no_of_cars, destination_index_array,
                             destination_choose_arrays, flags=LEARN_MODE0 | LEARN_DETECT)
        self._run(learner, episodes)
        sc = learner.env.scenario
        flagged = learner.env.deadlock_states().cpu().numpy()
        hit = flagged >= 0
        states = sorted({sc.decode(int(i)) for i in np.unique(flagged[hit])})
        self.detected_deadlocks = states
        bad_file = self.env_config.environment_bad_states_file
        known = set()
        if self.bad_states_text is not None:
            try:
                known = set(formats.read_bad_states(bad_file))
            except OSError:   # bad_states was passed as text and the layout's file does not exist yet
                known = {x for x in self.bad_states_text.split("\n") if x != ""}
        fresh = [s for s in states if s not in known]  # the reference compares the full string, headings included (:350)
        if len(fresh) < len(states):
            self._pull()
            return False
        if fresh:
            if self.bad_states_text is not None:
                if not os.path.exists(bad_file):
                    os.makedirs(os.path.dirname(bad_file) or ".", exist_ok=True)
                    with open(bad_file, "w") as f:
                        f.write(self.bad_states_text)
                for state in fresh:
                    formats.save_deadlock_state(state, self.env_config.station_box_list, bad_file,
                                                state.count('z') + 1, max_cars)
            # the episode in which each env was flagged rides in the high half of record word 3
            ep = (learner.env.records[:, 3].cpu().numpy().astype(np.int64) >> 16) & 0xFFFF
            self._pull()
            return int(ep[hit].min())
        self.save_qvalues(outputfile_number)
        self._pull()
        return episodes - 1

    # ---- validation (reference :384-437, minus the cv2 key-press UI) -----------------------------------
    def validate_given_state(self, episodes, inputfile_number, outputfile_number, input_car_states=None, no_of_cars=1,
                             destination_index_array=[0, 0, 0, 0], destination_choose_arrays=[0, 0, 0], waitkey=0,
                             max_rounds=4096):
        """Greedy (epsilon 0) rollout of the loaded table, no learning; cars move round-robin instead
        of by key press.  Returns {"success": fraction of envs that finished, "rounds": ...}."""
        self.load_qvalues(inputfile_number)
        learner = self._make(max(int(episodes), 1), input_car_states, no_of_cars, destination_index_array,
                             destination_choose_arrays, flags=LEARN_FROZEN, model=0)
        rounds = self._run(learner, max(int(episodes), 1), max_rounds=max_rounds)
        st = learner.env.stats()
        finished = st["episodes_done"] - st["episodes_capped"]
        return {"success": finished / float(self.n_envs * max(int(episodes), 1)), "rounds": rounds,
                "agent_steps": st["agent_steps"], "capped": st["episodes_capped"]}
