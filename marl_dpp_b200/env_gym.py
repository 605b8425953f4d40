"""Gym-style `Environment` with the reference's string API (reference src/env_gym.py:4-99), for ONE
env, computed by the CUDA kernels through the C ABI -- a drop-in for code written against the
reference env (`reset -> initialize_cars -> initialize -> select_reward_function`, then per
agent-step `encodestate -> getlegalactions -> get_successor_state / reward_function / update_car`
or `step`, reference src/algorithms/qlearning.py:197-229, dqn_open_source.py:490-517).

The dynamic semantics (state encoding with its ghost side effect, legal filter, rewards,
update_car, done) run on the GPU; only pure string plumbing and static-table lookups
(successor node of a node, unfiltered legal actions, BFS paths) are answered on the host from
the tables the Configuration holds.  The batched, fast path is marl_dpp_b200.batched."""
import random
from copy import deepcopy

import torch

from .batched import BatchedEnv
from .configuration import Configuration
from .tables import ACTIONS5, ACTIONS6, HEADS, Scenario, node_id, node_name


class _Car:
    """Read-only view of one car (reference src/configuration.py:19-35 attribute names)."""

    def __init__(self, car_id, source_node, destination_index, choose, done, done_count):
        self.car_id = car_id
        self.source_node = source_node
        self.destination_index = destination_index
        self.destination_choose_array = choose
        self.training_done = done
        self.training_done_count = done_count


class Environment:
    def __init__(self, env_conf, device=None):
        env_conf = Configuration.coerce(env_conf)
        self.action_mapping = {a: i for i, a in enumerate(ACTIONS6)}
        self.reverse_action_mapping = {i: a for i, a in enumerate(ACTIONS6)}
        self.env_conf = env_conf
        self.device = device
        self.box_node_list = env_conf.box_node_list
        self.node_action_list = env_conf.action_dictionary
        self.no_box_list = env_conf.no_box_list
        self.no_node_list = env_conf.no_node_list
        self.no_of_cars = env_conf.no_of_cars
        self.no_of_stations = env_conf.no_of_stations
        self.destination_nodes_list = env_conf.destination_nodes_list
        self.destination_box_list = env_conf.station_box_list
        self.selective_blocked_nodes_dictionary = env_conf.selective_blocked_nodes_dictionary
        self.selective_blocked_boxes_dictionary = env_conf.selective_blocked_boxes_dictionary
        self.training_complete_destination_index = env_conf.training_complete_destination_index
        self.blocked_nodes_list = []
        self.src_dest_length_dict = {}
        self.reward_function = None
        self._reward_name = None
        self._pending = None      # cars_info rows waiting for initialize()
        self._benv = None
        self._key = None
        if not hasattr(self, 'give_bad_state_reward'):  # kept across reset(), like the reference (:71)
            self.environment_bad_states_file = env_conf.environment_bad_states_file
            try:
                with open(self.environment_bad_states_file, 'r') as f:
                    self.bad_states = [x for x in f.read().split('\n') if x != '']
                self.give_bad_state_reward = True
            except OSError:  # the reference swallows every error here (:82); only a missing file is expected
                self.give_bad_state_reward = False

    # ---- episode set-up ---------------------------------------------------------------------------
    def reset(self):
        keep = (self._benv, self._key)
        self.__init__(self.env_conf, self.device)
        self._benv, self._key = keep

    def initialize_cars(self, no_of_cars):
        self.env_conf.no_of_cars = no_of_cars
        self.no_of_cars = no_of_cars
        self._pending = [deepcopy(self.env_conf.cars_info[i]) for i in range(no_of_cars)]

    def initialize(self, destination_index_array, input_car_states, destination_choose_arrays):
        """Headings are drawn with random.shuffle from Python's global generator exactly as the
        reference does (src/env_gym.py:345-349), so a seeded caller sees the same start nodes."""
        if input_car_states is None:
            return
        self.given_nodes = []
        for num in range(self.no_of_cars):
            nodes = deepcopy(self.box_node_list[input_car_states[num]])
            random.shuffle(nodes)
            self.given_nodes.append(nodes[0])
        self._place(destination_index_array, input_car_states, destination_choose_arrays)

    def reinitialize(self, destination_index_array, input_car_states, destination_choose_arrays):
        if input_car_states is not None:
            self._place(destination_index_array, input_car_states, destination_choose_arrays)

    def _place(self, dest_index, boxes, choose):
        n = self.no_of_cars
        key = (tuple(boxes[:n]), tuple(dest_index[:n]), tuple(tuple(c) for c in choose[:n]), self.give_bad_state_reward)
        if key != self._key:
            bad = "\n".join(self.bad_states) if self.give_bad_state_reward else None
            sc = Scenario(self.env_conf.layout, list(boxes[:n]), list(dest_index[:n]), [list(c) for c in choose[:n]],
                          bad_states_text=bad)
            self._benv, self._key = BatchedEnv(sc, 1, device=self.device), key
        self._choose = [list(c) for c in choose[:n]]
        heads = torch.tensor([[HEADS.index(x[-1]) for x in self.given_nodes]], dtype=torch.uint8)
        self._benv.reset(headings=heads)

    def select_reward_function(self, name):
        if name == 'q_reward':
            self.reward_function, self._reward_name = self.q_reward, name
        elif name == 'dqn_reward':
            self.reward_function, self._reward_name = self.dqn_reward, name

    # ---- static lookups (host tables) -----------------------------------------------------------
    def get_legal_actions(self, node):
        return self.env_conf.get_legal_actions(node)

    def get_successor_node(self, node, action):
        return self.node_action_list[node][self.action_mapping[action]]

    def state_to_array(self, state):
        return [part.split('x') for part in state.split('z')]

    def array_to_state(self, array):
        return 'z'.join(pair[0] + 'x' + pair[1] for pair in array)

    def get_successor_state(self, state, action):
        array = self.state_to_array(state)
        array[0][0] = self.get_successor_node(array[0][0], action)
        return self.array_to_state(array)

    def breadth_first_search(self, cars_source_destinations_pairs):
        return self.env_conf.breadth_first_search(cars_source_destinations_pairs[0][0],
                                                  cars_source_destinations_pairs[0][1])

    # ---- dynamic semantics (GPU) --------------------------------------------------------------------
    def _sc(self):
        if self._benv is None:
            raise RuntimeError("call initialize_cars() and initialize() first")
        return self._benv.scenario

    def encodestate(self, car_id, deadlock_id):
        """reference :281-329, including the ghost-counter side effect on the other cars."""
        out = self._benv.step_car(car_id, torch.tensor([0xFE], dtype=torch.uint8), deadlock_id=deadlock_id,
                                  reward=self._reward_name or 'q_reward')
        if int(out["status"][0]) == 1:
            raise ValueError("car %d has finished its training" % car_id)
        return self._sc().decode(int(out["state"][0]))

    def getlegalactions_filtered(self, state):
        if not isinstance(state, str):
            state = self.array_to_state(state)
        m = int(self._benv.query(0, [self._sc().encode(state)])["legal"][0])
        return [ACTIONS5[a] for a in range(5) if (m >> a) & 1]

    def getlegalactions(self, state):
        return self.getlegalactions_filtered(state)

    def _reward(self, previous_state, state, car_id, kind):
        sc = self._sc()
        prev, nxt = self.state_to_array(previous_state)[0][0], self.state_to_array(state)[0][0]
        action = next((a for a in range(5) if self.node_action_list[prev][self.action_mapping[ACTIONS5[a]]] == nxt), None)
        if action is None:
            raise ValueError("%s is not a successor of %s" % (state, previous_state))
        r = float(self._benv.query(car_id, [sc.encode(previous_state)], [action], reward=kind)["reward"][0])
        if kind == 'dqn_reward':
            return int(round(r * 100)) / 100  # the kernel returns float32; the reference `reward/100` in float64
        if not (self.give_bad_state_reward and r == -10000.0):
            return int(r)
        return r

    def q_reward(self, previous_state, state, car_id, action='S'):
        return self._reward(previous_state, state, car_id, 'q_reward')

    def dqn_reward(self, previous_state, state, car_id, action='S'):
        return self._reward(previous_state, state, car_id, 'dqn_reward')

    def update_car(self, car_id, source_node, destination_node, block_destination_nodes=0):
        self._benv.update_car(car_id, [node_id(source_node)])

    def step(self, state, action, num):
        """reference :85-99 -> (state, action, next_state, reward, next_state_array)."""
        next_state = self.get_successor_state(state, action)
        reward = self.reward_function(state, next_state, num, action)
        next_state_array = self.state_to_array(next_state)
        self.update_car(num, next_state_array[0][0], next_state_array[0][1], 0)
        return state, action, next_state, reward, next_state_array

    def get_cars(self):
        c = {k: v[0].tolist() for k, v in self._benv.cars().items() if v.dim() == 2}
        return [_Car(i, node_name(c["node"][i]), c["dest_index"][i], self._choose[i], c["done"][i],
                     {7: -1}.get(c["ghost"][i], c["ghost"][i])) for i in range(self.no_of_cars)]

    @property
    def cars(self):
        return self.get_cars()

    def check_car_training(self, num):
        return int(self._benv.cars()["done"][0, num])

    def is_game_ended(self):
        return bool(self._benv.cars()["done"][0].all())

    def update_simulator_display(self, waitkey):
        return -1  # the cv2 window is out of scope
