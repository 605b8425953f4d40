"""DQN data collection feed (BASELINE config 5): the batched env produces, per
agent-step, the tuple the reference's noisy double DQN remembers and replays
(reference src/algorithms/dqn_open_source.py:315-394,490-526):

    (convert(state), action index, convert(next_state), dqn_reward, mask(next_state))

written straight into a preallocated PyTorch ring buffer on the device.  The network (an
80-24-24-24-24-5 MLP in the reference) stays in PyTorch and is supplied by the caller as
`policy(features_uint8, legal_mask_uint8) -> actions uint8`; it is not a kernel target.

Replay semantics of the reference (dqn_open_source.py:315-394), on the device:
  * UniqueReplay -- remember() + the `deque(OrderedSet(self.memory))` de-duplication of replay():
    the memory is a SET of samples; batched it is one presence bit per compact key
    (state id * 5 + action) with a per-key payload (csrc/dqn_replay.cuh);
  * double_dqn_target -- the Double-DQN target with the legal-masked argmax, the reference's
    positional-index line reproduced (index_mode "reference") or corrected ("action").
"""
import ctypes as C

import torch

from . import _lib


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


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class UniqueReplay:
    """The reference's replay memory after its de-duplication (`deque(OrderedSet(self.memory))`,
    dqn_open_source.py:343): the set of distinct samples (state, action, next_state, reward), kept on the
    device as one presence bit per compact key `state id * 5 + action` plus the key's payload (next
    state id, reward, mask of the next state and -- with store_features -- convert(state) and
    convert(next_state)).  insert() takes a whole sub-step of the batched env; sample() draws a
    minibatch without replacement like `np.random.choice(len(memory), batch_size, replace=False)`."""

    def __init__(self, scenario, device, store_features=True):
        self.scenario = scenario
        self.device = dev = torch.device(device)
        self.n_keys = ((scenario.n_states * 5 + 31) // 32) * 32
        self.present = torch.zeros(self.n_keys // 32, dtype=torch.int32, device=dev)
        self.next = torch.zeros(self.n_keys, dtype=torch.int32, device=dev)
        self.reward = torch.zeros(self.n_keys, dtype=torch.float32, device=dev)
        self.mask = torch.zeros(self.n_keys, dtype=torch.uint8, device=dev)
        self.features = torch.zeros((self.n_keys, 160), dtype=torch.uint8, device=dev) if store_features else None
        self.counters = torch.zeros(3, dtype=torch.int64, device=dev)   # offered, new keys, differing duplicates
        self.keys = torch.zeros(self.n_keys, dtype=torch.int32, device=dev)
        self._count = torch.zeros(1, dtype=torch.int32, device=dev)
        self._dirty = True
        self._L = _lib.lib()

    def insert(self, state, action, next_state, reward, next_mask, status=None, features=None, next_features=None):
        """One sub-step of samples: int32 [n] state ids, uint8 [n] actions, int32 [n] next-state ids, float32 [n]
        rewards, uint8 [n] masks of the next states, uint8 [n] status (0 = a sample), uint8 [n, 80] feature rows."""
        n = state.numel()
        with torch.cuda.device(self.device):
            _lib.check(self._L.mdpp_replay_insert(
                n, _p(state), _p(action), _p(next_state), _p(reward), _p(next_mask), _p(status),
                _p(features if self.features is not None else None),
                _p(next_features if self.features is not None else None), self.n_keys, _p(self.present),
                _p(self.next), _p(self.reward), _p(self.mask), _p(self.features), _p(self.counters), _stream()))
        self._dirty = True

    def __len__(self):
        self._compact()
        return int(self._count.item())

    def _compact(self):
        if self._dirty:
            with torch.cuda.device(self.device):
                _lib.check(self._L.mdpp_replay_compact(self.n_keys, _p(self.present), _p(self.keys), _p(self._count),
                                                       _stream()))
            self._dirty = False

    def stats(self):
        c = self.counters.tolist()
        return {"offered": c[0], "unique": c[1], "differing_duplicates": c[2]}

    def sample(self, batch_size, generator=None):
        """-> dict(key, state, action, next_state, reward, next_mask[, features, next_features]) of `batch_size`
        distinct samples, or None while the set is smaller than the batch (reference :344-345)."""
        n = len(self)
        if n < batch_size:
            return None
        pick = torch.randperm(n, device=self.device, generator=generator)[:batch_size]
        key = self.keys[pick].long()
        out = {"key": key, "state": key // 5, "action": key % 5, "next_state": self.next[key].long(),
               "reward": self.reward[key], "next_mask": self.mask[key]}
        if self.features is not None:
            f = self.features[key]
            out["features"], out["next_features"] = f[:, :80], f[:, 80:]
        return out


def double_dqn_target(q_online_next, q_target_next, next_mask, reward, action, prediction, discount,
                      index_mode="reference", counters=None):
    """Double-DQN target of a minibatch (reference dqn_open_source.py:360-379): the online network picks the best
    VALID action of the next state, the target network values it; y = prediction with y[i, action[i]] =
    reward[i] + discount * Q_target(s'_i, a*_i).  index_mode "reference" reproduces the reference's
    `np.argmax(qvalues[masks[i]])` + `np.choose(max_actions, target.T)` pair (the position inside the masked
    sub-array is used as the action index); "action" maps the position back to the action.
    Returns (y float32 [n, 5], target float32 [n])."""
    n = q_online_next.shape[0]
    dev = q_online_next.device
    qo = q_online_next.detach().to(torch.float32).contiguous()
    qt = q_target_next.detach().to(torch.float32).contiguous()
    pr = prediction.detach().to(torch.float32).contiguous()
    y = torch.empty((n, 5), dtype=torch.float32, device=dev)
    t = torch.empty(n, dtype=torch.float32, device=dev)
    mode = {"reference": 0, "action": 1}[index_mode]
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mdpp_dqn_target(n, _p(qo), _p(qt), _p(next_mask.to(torch.uint8).contiguous()),
                                              _p(reward.to(torch.float32).contiguous()),
                                              _p(action.to(torch.int64).contiguous()), _p(pr), float(discount), mode,
                                              _p(y), _p(t), _p(counters), _stream()))
    return y, t


class DQNCollector:
    def __init__(self, env, ring, policy, reward="dqn_reward", memory=None):
        """memory: an optional UniqueReplay that receives every sample as well (de-duplicated on the device)."""
        self.env, self.ring, self.policy, self.reward = env, ring, policy, reward
        self.memory = memory
        self.round = 0

    def collect_round(self):
        """One round-robin round: for every car slot observe -> policy -> act, samples written in
        place into the ring.  Returns the number of ring rows written (rows of cars that did not
        move are marked invalid)."""
        env, ring = self.env, self.ring
        rows = 0
        for c in range(env.n_cars):
            sl = ring.reserve(env.n_envs)
            feats, legal, state = env.dqn_observe(c, features=ring.state[sl])
            actions = self.policy(feats, legal)
            moving = legal != 0
            actions = torch.where(moving, actions, torch.full_like(actions, 0xFF))
            out = env.dqn_act(c, actions, reward=self.reward, next_features=ring.next_state[sl],
                              out={"reward": ring.reward[sl], "next_legal": ring.next_mask[sl]},
                              auto_reset=True, round=self.round)
            ring.action[sl] = actions
            ring.valid[sl] = out["status"] == 0
            if self.memory is not None:  # the same samples into the de-duplicated set (state ids are its keys)
                ok = out["status"] == 0
                nxt = env.query(c, torch.where(ok, state, torch.zeros_like(state)),
                                torch.where(ok, actions, torch.full_like(actions, 0xFF)), reward=self.reward)["next_state"]
                self.memory.insert(state, actions, nxt, out["reward"], out["next_legal"], out["status"],
                                   ring.state[sl], ring.next_state[sl])
            rows += env.n_envs
        self.round += 1
        return rows
