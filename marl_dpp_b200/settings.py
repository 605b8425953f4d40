"""Algorithm-dict contract of the reference's settings.py (reference settings.py:7-35): the name of
the plugin module under algorithms/ and the keyword arguments its RL class is constructed with."""
TOTAL_CARS = 5
EPISODES = 5000
INPUT_DIM = 8 * TOTAL_CARS * 2
OUTPUT_DIM = 5
LEARNING_RATE = 0.0001
BATCH_SIZE = 32
SEED = 3

# the reference ships 'dqn_open_source' here (settings.py:27); the hot path of this repo is 'qlearning'
algorithm = 'qlearning'

params = {
    'input_dim': INPUT_DIM,
    'output_dim': OUTPUT_DIM,
    'cars': TOTAL_CARS,
    'batch_size': BATCH_SIZE,
}
