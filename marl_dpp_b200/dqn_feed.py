"""DQN data collection feed (BASELINE config 5): the batched env produces, per
agent-step, the tuple the reference's noisy double DQN remembers and replays
(reference src/algorithms/dqn_open_source.py:315-394,490-526):

    (convert(state), action index, convert(next_state), dqn_reward, mask(next_state))

written straight into a preallocated PyTorch ring buffer on the device.  The network (an
80-24-24-24-24-5 MLP in the reference) stays in PyTorch and is supplied by the caller as
`policy(features_uint8, legal_mask_uint8) -> actions uint8`; it is not a kernel target.
"""
import torch


class ReplayRing:
    """Fixed-capacity device-resident replay memory, struct-of-arrays, overwritten round-robin
    (the reference's deque(maxlen=mem_size), dqn_open_source.py:479-484, without its OrderedSet
    de-duplication, which is a host-side set operation on strings)."""

    def __init__(self, capacity, device):
        self.capacity = int(capacity)
        dev = torch.device(device)
        self.state = torch.zeros((self.capacity, 80), dtype=torch.uint8, device=dev)
        self.next_state = torch.zeros((self.capacity, 80), dtype=torch.uint8, device=dev)
        self.action = torch.zeros(self.capacity, dtype=torch.uint8, device=dev)
        self.reward = torch.zeros(self.capacity, dtype=torch.float32, device=dev)
        self.next_mask = torch.zeros(self.capacity, dtype=torch.uint8, device=dev)
        self.valid = torch.zeros(self.capacity, dtype=torch.bool, device=dev)
        self.pos = 0
        self.filled = 0

    def reserve(self, n):
        """A contiguous window of n rows (wraps to 0 when the tail is too short) -> slice."""
        if n > self.capacity:
            raise ValueError("batch of %d rows exceeds the ring capacity %d" % (n, self.capacity))
        if self.pos + n > self.capacity:
            self.valid[self.pos:] = False
            self.pos = 0
        sl = slice(self.pos, self.pos + n)
        self.pos = (self.pos + n) % self.capacity
        self.filled = min(self.capacity, self.filled + n)
        return sl

    def sample(self, batch_size, generator=None):
        """Uniform minibatch over the valid rows: (state, action, next_state, reward, next_mask)."""
        idx = torch.nonzero(self.valid).flatten()
        pick = idx[torch.randint(idx.numel(), (batch_size,), device=idx.device, generator=generator)]
        return (self.state[pick], self.action[pick].long(), self.next_state[pick], self.reward[pick],
                self.next_mask[pick])


def random_legal_policy(generator=None):
    """The reference's exploration branch (random.choice(valid_actions), dqn_open_source.py:231-232)."""
    def policy(features, legal):
        bits = (legal.unsqueeze(1) >> torch.arange(5, device=legal.device, dtype=torch.uint8)) & 1
        w = bits.float() + 1e-9 * (bits.sum(1, keepdim=True) == 0)
        return torch.multinomial(w, 1, generator=generator).flatten().to(torch.uint8)
    return policy


def masked_greedy_policy(model):
    """Masked argmax of a PyTorch Q-network over the legal actions (dqn_open_source.py:234-237)."""
    def policy(features, legal):
        with torch.no_grad():
            q = model(features.float())
        bits = ((legal.unsqueeze(1) >> torch.arange(5, device=legal.device, dtype=torch.uint8)) & 1).bool()
        return q.masked_fill(~bits, float("-inf")).argmax(1).to(torch.uint8)
    return policy


class DQNCollector:
    def __init__(self, env, ring, policy, reward="dqn_reward"):
        self.env, self.ring, self.policy, self.reward = env, ring, policy, reward
        self.round = 0

    def collect_round(self):
        """One round-robin round: for every car slot observe -> policy -> act, samples written in
        place into the ring.  Returns the number of ring rows written (rows of cars that did not
        move are marked invalid)."""
        env, ring = self.env, self.ring
        rows = 0
        for c in range(env.n_cars):
            sl = ring.reserve(env.n_envs)
            feats, legal, _ = env.dqn_observe(c, features=ring.state[sl])
            actions = self.policy(feats, legal)
            moving = legal != 0
            actions = torch.where(moving, actions, torch.full_like(actions, 0xFF))
            out = env.dqn_act(c, actions, reward=self.reward, next_features=ring.next_state[sl],
                              out={"reward": ring.reward[sl], "next_legal": ring.next_mask[sl]},
                              auto_reset=True, round=self.round)
            ring.action[sl] = actions
            ring.valid[sl] = out["status"] == 0
            rows += env.n_envs
        self.round += 1
        return rows
