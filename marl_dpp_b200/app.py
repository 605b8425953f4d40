"""Driver with the reference's flag surface (reference app.py:15-101):

    python -m marl_dpp_b200.app -i 0 -o 0 -dd 0 -v 0 -l 3 -bn 0 [--envs N] [--episodes E]

Selects `layout_<l>[bn]`, builds the Configuration, imports algorithms/<settings.algorithm>.RL with
settings.params (+ sess=None, env_config=) and runs the same per-initial-state loop."""
import argparse
import datetime
import importlib
import random
from copy import deepcopy

from . import layout as layout_module
from . import settings
from .configuration import Configuration


def parse_arguments(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-i", "--inputfile", required=True)
    ap.add_argument("-o", "--outputfile", required=True)
    ap.add_argument("-dd", "--detect_deadlock", required=True)
    ap.add_argument("-bn", "--block_number", required=True)
    ap.add_argument("-l", "--layout_number", required=True)
    ap.add_argument("-v", "--validate", required=True)
    ap.add_argument("--envs", type=int, default=1, help="lockstep env copies (1 = the reference's single env)")
    ap.add_argument("--episodes", type=int, default=settings.EPISODES)
    ap.add_argument("--reference-rng", action="store_true",
                    help="one env, decisions from Python's random in the reference's call order (same seed -> same run)")
    return vars(ap.parse_args(argv))


def give_configuration_object(layout_number, block_number):
    return Configuration(getattr(layout_module, 'layout_' + str(layout_number))[int(block_number)])


def start_training(alg, input_file_number, validate, ed, dd, output_file_number, episodes=settings.EPISODES):
    if validate:
        return [alg.validate_given_state(1, input_file_number, '0', i[0], len(i[0]), i[1], i[2], 0)
                for i in ed.initial_states]
    start_time = datetime.datetime.now()
    if dd and hasattr(alg, 'detect_deadlocks_v0'):
        for i in ed.initial_states:
            if len(i[0]) == 1:
                continue
            while True:
                episode = alg.detect_deadlocks_v0(100, '0', '0', i[0], len(i[0]), i[1], i[2])
                if not episode or episode == 99:
                    break
    for state in ed.initial_states:
        alg.train_given_state(episodes, input_file_number, output_file_number, state[0], len(state[0]),
                              state[1], state[2])
    return datetime.datetime.now() - start_time


def main(argv=None):
    args = parse_arguments(argv)
    random.seed(settings.SEED)      # the reference seeds at import (app.py:4, src/utils/utils.py:8)
    params = deepcopy(settings.params)
    ed = give_configuration_object(int(args["layout_number"]), int(args["block_number"]))
    RL = importlib.import_module(__package__ + '.algorithms.' + settings.algorithm).RL
    params['sess'] = None
    params['env_config'] = ed
    rl = RL(**params, n_envs=args["envs"], reference_rng=args["reference_rng"])
    out = start_training(rl, args["inputfile"], int(args["validate"]), ed, int(args["detect_deadlock"]),
                         args["outputfile"], args["episodes"])
    print(out, rl.last_stats)
    return rl


if __name__ == '__main__':
    main()
