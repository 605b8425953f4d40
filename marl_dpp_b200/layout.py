"""Layout / block selection -- same dict schema and values as the reference's `layout_3`
(reference src/layout.py:1-36 block 0 "up_left", :244-275 block 2 "down_right"; blocks 1 and 3
are named by the reference README but not implemented there either).

Only the live entries are carried; the reference file's several hundred commented-out
`initial_states` are usable as-is through `train_given_state(boxes, n, dest_index, choose)`.
`tests/test_tables.py` checks this module against the dict recorded from the reference.
"""

_CAR_COLOURS = [(0, 255, 0), (0, 0, 255), (255, 215, 0), (187, 255, 20), (0, 0, 0), (0, 0, 0), (0, 0, 0)]
_ALL_SIDES = ['N', 'E', 'W', 'S']


def _cars(starts, moves, dest_idx, choose_len, first=0):
    return [[s, c, m, d, [0] * choose_len]
            for s, c, m, d in zip(starts, _CAR_COLOURS[first:], moves, dest_idx)]


layout_3 = {
    0: {
        "number_of_rows": 3,
        "number_of_cols": 3,
        "bad_states_file": 'data/bad_states_files/layout-3/up_left.txt',
        "no_box_list": ['U6', 'U9', 'U3'],
        "no_node_list": [],
        "no_floor": 1,
        "environment_node_name_list": ['N', 'E', 'S', 'W'],
        "is_destination_node": True,
        "no_of_stations": 3,
        "disconnected_nodes_list": [],
        "station_box_list": ['U4', 'U2', 'U1'],
        "station_destination_list": [['U4N', 'U2W', 'U1N'], ['U2W', 'U1N', 'U4N']],
        "station_processing_count_list": [800, 800, 1000, 500, 1000, 800],
        "station_entrance_list": [list(_ALL_SIDES), list(_ALL_SIDES), list(_ALL_SIDES)],
        "no_of_cars": 3,
        "training_complete_destination_index": 1,
        "selective_blocked_nodes_dictionary": {},
        "selective_blocked_boxes_dictionary": {},
        "cars_info": (_cars(['U2E', 'U5E', 'U3E', 'U8E'], [1500, 100, 100, 100], [1, 1, 0, 0], 2)
                      + _cars(['U8N', 'U2E', 'U12W'], [100, 100, 100], [0, 0, 0], 3, first=4)),
        "initial_states": [
            [['U4'], [1], [[0, 0, 0]]],
        ],
    },
    2: {
        "number_of_rows": 3,
        "number_of_cols": 3,
        "bad_states_file": 'data/bad_states_files/layout-3/down_right.txt',
        "no_box_list": ['D9'],
        "no_node_list": [],
        "disconnected_nodes_list": ['D3S', 'D2N'],
        "no_floor": 1,
        "environment_node_name_list": ['N', 'E', 'S', 'W'],
        "is_destination_node": True,
        "no_of_stations": 3,
        "station_box_list": ['D6', 'D4', 'D8', 'D7'],
        "station_destination_list": [['D6N', 'D4E', 'D8E', 'D4N', 'D7S'], ['D4E', 'D8E', 'D6N', 'D4N', 'D7S']],
        "station_processing_count_list": [800, 800, 1000, 500, 1000, 800],
        "station_entrance_list": [['N', 'W'], list(_ALL_SIDES), ['E', 'W'], ['W', 'S']],
        "no_of_cars": 3,
        "training_complete_destination_index": 1,
        "selective_blocked_nodes_dictionary": {'D6N': ['D5N', 'D5E', 'D5W', 'D5S', 'D6N', 'D6E', 'D6W', 'D6S']},
        "selective_blocked_boxes_dictionary": {'D6': ['D5', 'D6']},
        "cars_info": _cars(['D2E', 'D5E', 'D1E', 'D2E', 'U8N', 'U2E', 'U10W'],
                           [1500, 100, 100, 100, 100, 100, 100], [1, 1, 0, 0, 0, 0, 0], 2),
        "initial_states": [
            [['D3', 'D2'], [0, 0], [[0, 1, 0], [0, 1, 0]]],
        ],
    },
}

BLOCK_NAMES = {0: "up_left", 2: "down_right"}
