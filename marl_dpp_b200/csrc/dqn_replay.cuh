// dqn_replay.cuh -- the replay side of the noisy double DQN's collection loop, sm_100a.
//
// Reference (src/algorithms/dqn_open_source.py): remember() appends the sample (state, action index, next_state,
// reward) to a deque (:315-323); replay() first de-duplicates the whole memory -- `deque(OrderedSet(self.memory))`
// (:343) -- then draws a minibatch without replacement (:347-350) and forms the Double-DQN target (:360-379).
//
// De-duplication.  A sample is a tuple of strings and numbers; two samples are equal iff (state, action, next
// state, reward) are.  Batched, states are ids, and the key space (state id x 5 actions) is enumerable, so the set
// is a PERFECT hash: one presence bit per compact key  state * 5 + action  (next state and reward are functions of
// the key under legal actions -- the collector only takes legal ones; the rare differing pair would be a second
// tuple in the reference and is counted here).  Millions of samples per sub-step collapse to the few thousand
// distinct keys with one L2 atomic OR each; the first writer of a key stores its payload (next id, reward, mask of
// the next state, the two 80-byte feature rows).  replay_compact_kernel then lists the keys in ascending order --
// the deterministic stand-in for the OrderedSet's first-occurrence order, which only matters to the reference
// through the positions a uniform draw picks -- and the minibatch is gathered from the per-key payload.
//
// Double-DQN target (:360-379), per sample:  a* = argmax over the VALID actions of the online network's Q(s', .),
// target = reward + discount * Q_target(s', a*), written into the online prediction at the taken action.
// The reference computes a* as `np.argmax(qvalues[masks[i]])`, i.e. the POSITION inside the masked sub-array, and
// then uses that position as the action index in `np.choose(max_actions, target.T)`.  index_mode 0 reproduces
// that line exactly (positions == actions whenever no action below the best one is masked); index_mode 1 maps
// the position back to the action.  Arithmetic as in the reference's NumPy expression: float32 product (scalar
// discount x float32 network output), float64 sum with the float64 reward, rounded to float32 when stored into the
// prediction array.
#pragma once
#include "env_core.cuh"

namespace mdpp {

struct ReplayMem {
    uint32_t *present;   // [n_keys / 32] one bit per compact key
    uint32_t *next;      // [n_keys] next-state id
    float *reward;       // [n_keys]
    uint8_t *mask;       // [n_keys] 5-bit legal mask of the next state
    uint8_t *features;   // [n_keys][160] convert(state) | convert(next_state), or NULL
    unsigned long long *counters; // [0] samples offered, [1] new keys, [2] duplicates whose payload differed
    uint32_t n_keys;
};

__global__ void __launch_bounds__(kThreads)
replay_insert_kernel(int64_t n, const uint32_t *__restrict__ state_idx, const uint8_t *__restrict__ action,
                     const uint32_t *__restrict__ next_idx, const float *__restrict__ reward,
                     const uint8_t *__restrict__ next_mask, const uint8_t *__restrict__ status,
                     const uint8_t *__restrict__ features, const uint8_t *__restrict__ next_features, ReplayMem mem) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t offered = 0, fresh = 0;
    if (i < n && (!status || status[i] == 0) && action[i] < MDPP_N_ACT) {
        const uint64_t k64 = (uint64_t)state_idx[i] * MDPP_N_ACT + action[i];
        if (k64 < mem.n_keys) {
            const uint32_t key = (uint32_t)k64, bit = 1u << (key & 31u);
            offered = 1;
            if (!(atomicOr(mem.present + (key >> 5), bit) & bit)) { // first of its key: store the payload
                fresh = 1;
                mem.next[key] = next_idx[i];
                mem.reward[key] = reward[i];
                mem.mask[key] = next_mask[i];
                if (mem.features) {
                    uint4 *dst = reinterpret_cast<uint4 *>(mem.features + (size_t)key * 160u);
                    const uint4 *a = reinterpret_cast<const uint4 *>(features + (size_t)i * 80u);
                    const uint4 *b = reinterpret_cast<const uint4 *>(next_features + (size_t)i * 80u);
#pragma unroll
                    for (int j = 0; j < 5; j++) dst[j] = a[j];
#pragma unroll
                    for (int j = 0; j < 5; j++) dst[5 + j] = b[j];
                }
            }
        }
    }
    offered = __reduce_add_sync(0xFFFFFFFFu, offered);
    fresh = __reduce_add_sync(0xFFFFFFFFu, fresh);
    if ((threadIdx.x & 31u) == 0) {
        if (offered) atomicAdd(mem.counters + 0, (unsigned long long)offered);
        if (fresh) atomicAdd(mem.counters + 1, (unsigned long long)fresh);
    }
}

// duplicates of keys stored by an EARLIER launch whose (next state, reward) differ from the stored payload: in the
// reference those would be distinct tuples.  Counted, never expected (next state and reward follow from the key).
__global__ void __launch_bounds__(kThreads)
replay_check_kernel(int64_t n, const uint32_t *__restrict__ state_idx, const uint8_t *__restrict__ action,
                    const uint32_t *__restrict__ next_idx, const float *__restrict__ reward,
                    const uint8_t *__restrict__ status, ReplayMem mem) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t differ = 0;
    if (i < n && (!status || status[i] == 0) && action[i] < MDPP_N_ACT) {
        const uint64_t k64 = (uint64_t)state_idx[i] * MDPP_N_ACT + action[i];
        if (k64 < mem.n_keys && ((mem.present[k64 >> 5] >> (k64 & 31u)) & 1u)) // only keys an EARLIER call stored
            differ = mem.next[k64] != next_idx[i] || mem.reward[k64] != reward[i];
    }
    differ = __reduce_add_sync(0xFFFFFFFFu, differ);
    if ((threadIdx.x & 31u) == 0 && differ) atomicAdd(mem.counters + 2, (unsigned long long)differ);
}

// keys of the set in ascending order -> list[rank]; *count = size of the set.  One CTA walks the bitmap in chunks of
// 1024 words with a running offset (block scan per chunk): the bitmap of the largest table this library handles is
// a few MB, and the list is rebuilt only when samples were inserted.
__global__ void __launch_bounds__(1024)
replay_compact_kernel(const uint32_t *__restrict__ present, uint32_t n_words, uint32_t n_keys, uint32_t *__restrict__ list,
                      uint32_t *__restrict__ count) {
    __shared__ uint32_t warp_excl[32];
    __shared__ uint32_t running, chunk_total;
    if (threadIdx.x == 0) running = 0u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t w0 = 0; w0 < n_words; w0 += 1024u) {
        const uint32_t w = w0 + threadIdx.x;
        uint32_t bits = w < n_words ? present[w] : 0u;
        const uint32_t cnt = (uint32_t)__popc(bits);
        uint32_t incl = cnt; // inclusive prefix inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        if (lane == 31u) warp_excl[warp] = incl;
        __syncthreads();
        if (warp == 0) { // exclusive prefix of the 32 warp totals
            const uint32_t s = warp_excl[lane];
            uint32_t t = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, t, o);
                if (lane >= (uint32_t)o) t += y;
            }
            warp_excl[lane] = t - s;
            if (lane == 31u) chunk_total = t;
        }
        __syncthreads();
        uint32_t slot = running + warp_excl[warp] + incl - cnt;
        while (bits) {
            const uint32_t b = (uint32_t)__ffs((int)bits) - 1u;
            bits &= bits - 1u;
            const uint32_t key = w * 32u + b;
            if (key < n_keys) list[slot++] = key;
        }
        __syncthreads(); // everybody has read `running`
        if (threadIdx.x == 0) running += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = running;
}

// Double-DQN target (reference :360-379) for a minibatch.  q_online_next / q_target_next / prediction: [n][5].
__global__ void __launch_bounds__(kThreads)
dqn_target_kernel(int64_t n, const float *__restrict__ q_online_next, const float *__restrict__ q_target_next,
                  const uint8_t *__restrict__ next_mask, const float *__restrict__ reward, const int64_t *__restrict__ action,
                  const float *__restrict__ prediction, double discount, int index_mode, float *__restrict__ y,
                  float *__restrict__ target_out, unsigned long long *__restrict__ counters /* [0] invalid_count, [1] empty masks */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t invalid = 0, empty = 0;
    if (i < n) {
        const float *qo = q_online_next + i * MDPP_N_ACT, *qt = q_target_next + i * MDPP_N_ACT;
        const uint32_t m = next_mask[i] & 31u;
        int overall = 0, pos = 0, best_pos = -1, best_act = 0;
        float best = 0.f;
#pragma unroll
        for (int k = 0; k < MDPP_N_ACT; k++) {
            if (qo[k] > qo[overall]) overall = k;                 // np.argmax: first maximum
            if ((m >> k) & 1u) {
                if (best_pos < 0 || qo[k] > best) { best = qo[k]; best_pos = pos; best_act = k; }
                pos++;
            }
        }
        if (best_pos < 0) { empty = 1; best_pos = 0; best_act = 0; } // the reference raises on an empty mask
        invalid = overall != best_pos;                              // `overall_max_action != max_action` (:367)
        const int idx = index_mode == 0 ? best_pos : best_act;
        // NumPy's arithmetic for `rewards + self.discount * max_qvalues`: the Python float scalar times a float32
        // array is a FLOAT32 product (the scalar is cast to the array's type); adding the float64 reward array
        // promotes to float64; storing into the float32 prediction array rounds once more.  The reference's rewards
        // are float64 k / 100 (dqn_reward, src/env_gym.py:553) or integers; the env kernels emit float32(k / 100), so
        // k is recovered and divided again in float64.
        const float prod = __fmul_rn((float)discount, qt[idx]);
        const double r64 = (double)lrintf(__fmul_rn(reward[i], 100.0f)) / 100.0;
        const float t = (float)__dadd_rn(r64, (double)prod);
#pragma unroll
        for (int k = 0; k < MDPP_N_ACT; k++) y[i * MDPP_N_ACT + k] = prediction[i * MDPP_N_ACT + k];
        const int64_t a = action[i];
        if (a >= 0 && a < MDPP_N_ACT) y[i * MDPP_N_ACT + a] = t;
        if (target_out) target_out[i] = t;
    }
    invalid = __reduce_add_sync(0xFFFFFFFFu, invalid);
    empty = __reduce_add_sync(0xFFFFFFFFu, empty);
    if ((threadIdx.x & 31u) == 0 && counters) {
        if (invalid) atomicAdd(counters + 0, (unsigned long long)invalid);
        if (empty) atomicAdd(counters + 1, (unsigned long long)empty);
    }
}

} // namespace mdpp
