"""`Configuration` -- the object the reference hands to its algorithm plugins
(reference src/configuration.py:38-145, built by app.give_configuration_object, app.py:34-39).

Same attribute names, so code written against the reference's `env_config` keeps working; the
successor table behind `action_dictionary` comes from the arithmetic StaticGraph, and nothing is
printed or drawn (the reference's Display / cv2 UI is out of scope)."""
from copy import deepcopy

from .tables import ACTIONS6, NONE, StaticGraph, node_id, node_name

_HEADS = ['N', 'E', 'S', 'W']


class Configuration:
    def __init__(self, env_config):
        c = deepcopy(env_config)  # the reference grows no_node_list in place (configuration.py:52,138-141)
        self.layout = c
        self.graph = StaticGraph(c)
        self.disconnected_nodes_list = c['disconnected_nodes_list']
        self.environment_rows = c["number_of_rows"]
        self.environment_columns = c["number_of_cols"]
        self.environment_bad_states_file = c["bad_states_file"]
        self.initial_states = c["initial_states"]
        self.no_box_list = c["no_box_list"]
        self.no_floor = c["no_floor"]
        self.environment_node_name_list = c["environment_node_name_list"]
        self.training_complete_destination_index = c["training_complete_destination_index"]
        self.selective_blocked_nodes_dictionary = c["selective_blocked_nodes_dictionary"]
        self.selective_blocked_boxes_dictionary = c["selective_blocked_boxes_dictionary"]
        self.is_destination_node = c["is_destination_node"]
        self.no_of_stations = c["no_of_stations"]
        self.station_box_list = c["station_box_list"]
        self.destination_nodes_list = c["station_destination_list"]
        self.station_processing_count_list = c["station_processing_count_list"]
        self.station_entrance_list = c["station_entrance_list"]
        self.no_of_cars = c["no_of_cars"]
        self.cars_info = c["cars_info"]
        self.blocked_nodes_list = []
        # box order of the reference's nested loop (configuration.py:100-107): U1-3, D1-9, U4-9
        self.box_list = ['U%d' % b for b in (1, 2, 3)] + ['D%d' % b for b in range(1, 10)] + \
                        ['U%d' % b for b in range(4, 10)]
        self.box_node_list = {b: [b + h for h in _HEADS] for b in self.box_list}
        self.action_dictionary = {}
        for b in self.box_list:
            for n in self.box_node_list[b]:
                row = []
                for a in range(6):
                    s = int(self.graph.succ[node_id(n), a])
                    row.append(n[0] + 'YY' if s == NONE else node_name(s))
                self.action_dictionary[n] = row
        self.no_node_list = list(c["no_node_list"])
        for b in self.no_box_list:
            self.no_node_list.extend(self.box_node_list[b])
        self.stations_info = [[self.station_box_list[i], 0, self.station_processing_count_list[i]]
                              for i in range(self.no_of_stations)]

    @classmethod
    def from_reference(cls, ref_config):
        """Accept the reference's own Configuration object (reference src/configuration.py:44-145): the
        layout dict is rebuilt from its attributes, so `RL(..., env_config=ed)` works unchanged when
        the plugin is imported from the reference's app.py."""
        r = ref_config
        return cls({
            "number_of_rows": r.environment_rows, "number_of_cols": r.environment_columns,
            "bad_states_file": r.environment_bad_states_file, "no_box_list": list(r.no_box_list),
            "no_node_list": [],  # the reference's list is the derived one (every node of a no_box)
            "no_floor": r.no_floor, "environment_node_name_list": list(r.environment_node_name_list),
            "is_destination_node": r.is_destination_node, "no_of_stations": r.no_of_stations,
            "disconnected_nodes_list": list(r.disconnected_nodes_list), "station_box_list": list(r.station_box_list),
            "station_destination_list": [list(x) for x in r.destination_nodes_list],
            "station_processing_count_list": list(r.station_processing_count_list),
            "station_entrance_list": [list(x) for x in r.station_entrance_list], "no_of_cars": r.no_of_cars,
            "training_complete_destination_index": r.training_complete_destination_index,
            "selective_blocked_nodes_dictionary": dict(r.selective_blocked_nodes_dictionary),
            "selective_blocked_boxes_dictionary": dict(r.selective_blocked_boxes_dictionary),
            "cars_info": deepcopy(r.cars_info), "initial_states": deepcopy(r.initial_states)})

    @classmethod
    def coerce(cls, env_config):
        if isinstance(env_config, cls):
            return env_config
        if isinstance(env_config, dict):
            return cls(env_config)
        return cls.from_reference(env_config)

    def get_legal_actions(self, node):
        """reference src/env_gym.py:114-130 (unfiltered, 'B' included)."""
        m = int(self.graph.base_legal6[node_id(node)])
        return [ACTIONS6[a] for a in range(6) if (m >> a) & 1]

    def breadth_first_search(self, source_node, destination_node):
        """reference src/env_gym.py:443-483: (actions, visited boxes); expansion order S,L,R,F,T, FIFO."""
        from collections import deque
        fringe = deque([(source_node, [], [])])
        closed = {source_node}
        while fringe:
            node, acts, boxes = fringe.popleft()
            for action in self.get_legal_actions(node):
                if action == 'B':
                    continue
                item = self.action_dictionary[node][ACTIONS6.index(action)]
                if item == destination_node:
                    return acts + [action], boxes + [item[:-1]]
                if item in closed:
                    continue
                fringe.append((item, acts + [action], boxes + [item[:-1]]))
                closed.add(item)
        return ['S'], []
