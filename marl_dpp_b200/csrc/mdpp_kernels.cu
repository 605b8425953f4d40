// mdpp_kernels.cu -- sm_100a kernels of the batched MARL-DPP hot path.
//
// Mapping: one lane = one env, one warp = a tile of 32 consecutive envs, 16-byte env records read
// and written with one 128-bit access per lane (512 contiguous bytes per warp).  All env work is
// integer ALU + shared-memory table lookups; HBM/L2 traffic is the env records, the optional
// per-step outputs and (learner) two 32-byte Q-row gathers plus the accumulator atomics.
#include "env_core.cuh"

namespace mdpp {

// ------------------------------------------------------------------------------------------------
// workspace-resident control block
// ------------------------------------------------------------------------------------------------
// Counters are privatised over kStatSlots slots (one 64-byte sector each): a CTA reduces its own
// counts in shared memory and adds them to slot (blockIdx & (kStatSlots-1)).  With one slot every
// warp of a 1M-env launch hit the same sector with 5 atomics and the launch waited ~60 us on it.
constexpr int kStatSlots = 128;
struct Control {
    mdpp_stats stats[kStatSlots];
    uint32_t reserved[16]; // (round 1's list mode kept its touched-key counters here)
    // largest per-env episode counter the learner kernels have produced, privatised like the stats; the host
    // uses it (mdpp_get_stats_host) to tighten its bound on "is any env past the epsilon = 1 segment?"
    uint32_t max_episode[kStatSlots];
};

// Index guards: compute-sanitizer is closed on this GPU pool, so data-dependent indices that are not masked by
// construction are range-checked before use.  The sub-step and list kernels skip the access and count the event
// (mdpp_stats.internal_errors, asserted 0 by the tests, which drive the same tables and state encoding as the
// fused kernels); the round kernel and the dense pull, where a counted guard measurably costs time, clamp
// branch-free instead.
__device__ __forceinline__ void guard_tripped(Control *ctl) { atomicAdd((unsigned long long *)&ctl->stats[0].internal_errors, 1ull); }

// End-to-end calls (mdpp_q_train_rounds_host): the counters go to the host WITHOUT a copy command and without a
// stream synchronisation.  One small kernel after the round's kernels sums the privatised slots and stores the
// totals, then a sequence number, straight into pinned host memory (zero-copy over PCIe); the host thread polls that
// word.  Against cudaMemcpyAsync(8 KB) + cudaStreamSynchronize this takes ~15 us out of every host-facing call.
struct HostStats {
    mdpp_stats totals;
    uint32_t max_episode;
    uint32_t pad;
    volatile unsigned long long seq; // written last
};
__global__ void __launch_bounds__(128)
stats_publish_kernel(const Control *__restrict__ ctl, HostStats *host, unsigned long long seq) {
    static_assert(kStatSlots == 128, "one thread per slot");
    __shared__ unsigned long long acc[8];
    __shared__ uint32_t mx;
    if (threadIdx.x < 8) acc[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) mx = 0u;
    __syncthreads();
    const mdpp_stats s = ctl->stats[threadIdx.x];
    const unsigned long long v[8] = {s.agent_steps, s.q_updates, s.episodes_done, s.episodes_capped,
                                     (unsigned long long)s.reward_sum, s.explore_steps, s.touched_keys, s.internal_errors};
#pragma unroll
    for (int k = 0; k < 8; k++) {
        unsigned long long x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
        if ((threadIdx.x & 31u) == 0 && x) atomicAdd(&acc[k], x);
    }
    const uint32_t me = __reduce_max_sync(0xFFFFFFFFu, ctl->max_episode[threadIdx.x]);
    if ((threadIdx.x & 31u) == 0) atomicMax(&mx, me);
    __syncthreads();
    if (threadIdx.x == 0) {
        host->totals.agent_steps = acc[0];
        host->totals.q_updates = acc[1];
        host->totals.episodes_done = acc[2];
        host->totals.episodes_capped = acc[3];
        host->totals.reward_sum = (int64_t)acc[4];
        host->totals.explore_steps = acc[5];
        host->totals.touched_keys = acc[6];
        host->totals.internal_errors = acc[7];
        host->max_episode = mx;
        __threadfence_system();
        host->seq = seq;
    }
}

struct StatAcc {
    uint32_t steps = 0, episodes = 0, capped = 0, explore = 0;
    int32_t reward = 0;
};

struct SmemStats {
    unsigned long long reward;
    unsigned long long packed; // kSmall kernels: steps | explore << 16 | episodes << 32 | capped << 48
    uint32_t steps, episodes, capped, explore;
};

// Call once per thread at kernel end; `ss` must have been zeroed before a __syncthreads().
// kSmall: every per-thread counter is 0/1 and |reward| < 2^26 (one agent-step per thread), so the
// five counters travel in one REDUX and one shared-memory atomic per warp.  kQ: each step is a Q-update.
template <bool kSmall, bool kQ>
__device__ __forceinline__ void flush_stats(const StatAcc &a, int64_t reward64, Control *ctl, SmemStats &ss) {
    const uint32_t lane = threadIdx.x & 31u;
    if constexpr (kSmall) {
        uint32_t packed = a.steps | (a.explore << 8) | (a.episodes << 16) | (a.capped << 24);
        packed = __reduce_add_sync(0xFFFFFFFFu, packed);
        const int rs = (int)__reduce_add_sync(0xFFFFFFFFu, (unsigned)(int)reward64);
        if (lane == 0) {
            const unsigned long long wide = (unsigned long long)(packed & 0xFFu) |
                                            ((unsigned long long)((packed >> 8) & 0xFFu) << 16) |
                                            ((unsigned long long)((packed >> 16) & 0xFFu) << 32) |
                                            ((unsigned long long)(packed >> 24) << 48);
            if (wide) atomicAdd(&ss.packed, wide);
            if (rs) atomicAdd(&ss.reward, (unsigned long long)(long long)rs);
        }
    } else {
        uint32_t steps = __reduce_add_sync(0xFFFFFFFFu, a.steps);
        uint32_t episodes = __reduce_add_sync(0xFFFFFFFFu, a.episodes);
        uint32_t capped = __reduce_add_sync(0xFFFFFFFFu, a.capped);
        uint32_t explore = __reduce_add_sync(0xFFFFFFFFu, a.explore);
        long long rsum = reward64;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xFFFFFFFFu, rsum, o);
        if (lane == 0) {
            if (steps) atomicAdd(&ss.steps, steps);
            if (episodes) atomicAdd(&ss.episodes, episodes);
            if (capped) atomicAdd(&ss.capped, capped);
            if (explore) atomicAdd(&ss.explore, explore);
            if (rsum) atomicAdd(&ss.reward, (unsigned long long)rsum);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mdpp_stats *slot = &ctl->stats[blockIdx.x & (kStatSlots - 1)];
        const unsigned long long steps = ss.steps + (ss.packed & 0xFFFFull);
        const unsigned long long explore = ss.explore + ((ss.packed >> 16) & 0xFFFFull);
        const unsigned long long episodes = ss.episodes + ((ss.packed >> 32) & 0xFFFFull);
        const unsigned long long capped = ss.capped + (ss.packed >> 48);
        if (steps) atomicAdd((unsigned long long *)&slot->agent_steps, steps);
        if (episodes) atomicAdd((unsigned long long *)&slot->episodes_done, episodes);
        if (capped) atomicAdd((unsigned long long *)&slot->episodes_capped, capped);
        if (explore) atomicAdd((unsigned long long *)&slot->explore_steps, explore);
        if (kQ && steps) atomicAdd((unsigned long long *)&slot->q_updates, steps);
        if (ss.reward) atomicAdd((unsigned long long *)&slot->reward_sum, ss.reward);
    }
}

__device__ __forceinline__ void zero_stats(SmemStats &ss) {
    if (threadIdx.x == 0) ss = SmemStats{0ull, 0ull, 0u, 0u, 0u, 0u};
}

// ------------------------------------------------------------------------------------------------
// reset / initialize  (reference src/env_gym.py:643-672,331-352)
// ------------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(kThreads)
reset_kernel(DevCfg cfg, uint4 *__restrict__ recs, int64_t n_envs, uint32_t env_id_base,
             const uint8_t *__restrict__ env_mask, const uint8_t *__restrict__ headings, uint32_t seed,
             int zero_counters) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_envs) return;
    if (env_mask && !env_mask[e]) return;
    uint32_t hb = 0;
    if (headings) {
#pragma unroll
        for (int i = 0; i < NC; i++) hb |= (uint32_t)(headings[e * NC + i] & 3u) << (2 * i);
    } else {
        hb = heading_word(0, 3, seed, env_id_base + (uint32_t)e);
    }
    uint64_t cars = initial_cars<NC>(cfg, hb);
    uint4 r = recs[e];
    r.x = (uint32_t)cars;
    r.y = (uint32_t)(cars >> 32);
    if (zero_counters) { r.z = 0; r.w = 0; }
    else { r.z &= 0xFFFFu; r.w &= ~0xFFu; } // new episode: steps and stall run = 0, episode index kept
    recs[e] = r;
}

// ------------------------------------------------------------------------------------------------
// one agent-step of car `c` in every env, externally supplied actions (parity / gym mode)
// ------------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(kThreads)
step_car_kernel(DevCfg cfg, uint4 *__restrict__ recs, int64_t n_envs, int c, int mode, int kind,
                const uint8_t *__restrict__ actions, uint32_t *__restrict__ state_idx,
                uint32_t *__restrict__ next_idx, float *__restrict__ reward, uint8_t *__restrict__ legal,
                uint8_t *__restrict__ status, uint8_t *__restrict__ done_out, Control *ctl) {
    __shared__ SmemTables sm;
    __shared__ SmemStats ss;
    zero_stats(ss);
    load_tables(sm, cfg.tables);
    const SharedTab tab{&sm};
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    StatAcc acc;
    int64_t rsum = 0;
    if (e < n_envs) {
        uint4 rec = recs[e];
        uint64_t cars = (uint64_t)rec.x | ((uint64_t)rec.y << 32);
        uint32_t st = 0, sidx = 0xFFFFFFFFu, nidx = 0xFFFFFFFFu, lm = 0;
        float rw = 0.f;
        uint32_t a = actions[e];
        if (f_done(car_field(cars, c)) || a == MDPP_ACT_SKIP) {
            st = 1; // reference: `if env.check_car_training(num): continue` (qlearning.py:216-217)
        } else {
            View v = view_of<NC>(cars, c, mode, cfg);
            lm = legal_mask(v, v.pnode, tab);
            sidx = state_index(v, v.pnode, cfg);
            if (a >= MDPP_N_ACT || !((lm >> a) & 1u)) {
                st = 2;
            } else {
                uint32_t nn = next_node(v.pnode, a, tab);
                nidx = state_index(v, nn, cfg);
                int32_t ri = reward_int<NC>(v, nn, cars, c, kind, cfg, tab);
                rw = reward_float(ri, kind);
                cars = advance_car(cars, c, v, nn, cfg);
                uint32_t steps = rec.z >> 16;
                steps = steps < 0xFFFFu ? steps + 1 : steps;
                rec.z = (rec.z & 0xFFFFu) | (steps << 16);
                acc.steps = 1;
                rsum = ri;
            }
            rec.x = (uint32_t)cars; // ghost counters move even when the action was rejected
            rec.y = (uint32_t)(cars >> 32);
            recs[e] = rec;
        }
        if (state_idx) state_idx[e] = sidx;
        if (next_idx) next_idx[e] = nidx;
        if (reward) reward[e] = rw;
        if (legal) legal[e] = (uint8_t)lm;
        if (status) status[e] = (uint8_t)st;
        if (done_out) done_out[e] = (uint8_t)all_done<NC>(cars);
    }
    flush_stats<true, false>(acc, rsum, ctl, ss);
}

// ------------------------------------------------------------------------------------------------
// lockstep rollout, uniform-random legal policy, auto-reset (config "random-policy sweep")
// ------------------------------------------------------------------------------------------------
// 4 CTAs/SM (<= 64 registers): the unrolled per-car loop wants ~95 registers, but at 2 CTAs/SM the
// kernel sat at 24 % occupancy stalled on fixed-latency ALU dependencies (ncu: "wait")
template <int NC>
__global__ void __launch_bounds__(kThreads, 4)
rollout_random_kernel(DevCfg cfg, uint4 *__restrict__ recs, int64_t n_envs, uint32_t env_id_base, int rounds,
                      uint64_t first_round, uint32_t seed, int kind, uint8_t *__restrict__ action_out,
                      float *__restrict__ reward_out, uint8_t *__restrict__ done_out,
                      uint32_t *__restrict__ next_out, Control *ctl) {
    __shared__ SmemTables sm;
    __shared__ SmemStats ss;
    zero_stats(ss);
    load_tables(sm, cfg.tables);
    const SharedTab tab{&sm};
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    StatAcc acc;
    int64_t rsum = 0;
    if (e < n_envs) {
        uint4 rec = recs[e];
        uint64_t cars = (uint64_t)rec.x | ((uint64_t)rec.y << 32);
        uint32_t episode = rec.z & 0xFFFFu, steps = rec.z >> 16;
        const uint32_t env = env_id_base + (uint32_t)e;
        for (int r = 0; r < rounds; r++) {
            const uint64_t round = first_round + (uint64_t)r;
            uint4 p = philox4x32_r((uint32_t)round, (uint32_t)(round >> 32), 0, 0, seed, env);
#pragma unroll
            for (int c = 0; c < NC; c++) {
                uint32_t a = MDPP_ACT_SKIP, nidx = 0xFFFFFFFFu;
                float rw = 0.f;
                if (!f_done(car_field(cars, c))) {
                    View v = view_of<NC>(cars, c, 1, cfg);
                    uint32_t lm = legal_mask(v, v.pnode, tab);
                    if (lm) {
                        uint32_t w = c == 0 ? p.x : c == 1 ? p.y : c == 2 ? p.z : c == 3 ? p.w
                                     : philox4x32_r((uint32_t)round, (uint32_t)(round >> 32), 1, 0, seed, env).x;
                        // random.choice(legal) (qlearning.py:102): low 16 bits scaled to the legal count
                        a = nth_set_bit(lm, ((w & 0xFFFFu) * (uint32_t)__popc(lm)) >> 16);
                        uint32_t nn = next_node(v.pnode, a, tab);
                        nidx = state_index(v, nn, cfg);
                        int32_t ri = reward_int<NC>(v, nn, cars, c, kind, cfg, tab);
                        rw = reward_float(ri, kind);
                        cars = advance_car(cars, c, v, nn, cfg);
                        steps = steps < 0xFFFFu ? steps + 1 : steps;
                        acc.steps++;
                        acc.explore++;
                        rsum += ri;
                    }
                }
                size_t o = ((size_t)r * NC + c) * (size_t)n_envs + (size_t)e;
                if (action_out) action_out[o] = (uint8_t)a;
                if (reward_out) reward_out[o] = rw;
                if (next_out) next_out[o] = nidx;
                if (done_out) done_out[o] = (uint8_t)all_done<NC>(cars);
            }
            // end of round: is_game_ended (src/env_gym.py:622-638) or the step cap -> next episode
            bool finished = all_done<NC>(cars);
            bool capped = !finished && cfg.step_cap > 0 && steps >= (uint32_t)cfg.step_cap;
            if (finished || capped) {
                acc.episodes++;
                acc.capped += capped;
                episode = (episode + 1) & 0xFFFFu;
                steps = 0;
                cars = initial_cars<NC>(cfg, heading_word(round, 2, seed, env));
            }
        }
        rec.x = (uint32_t)cars;
        rec.y = (uint32_t)(cars >> 32);
        rec.z = episode | (steps << 16);
        recs[e] = rec;
    }
    flush_stats<false, false>(acc, rsum, ctl, ss);
}

// ------------------------------------------------------------------------------------------------
// learner: epsilon-greedy select + env step + key marks, and the pull kernels that apply the update
// ------------------------------------------------------------------------------------------------
} // namespace mdpp
#include "q_learner.cuh" // needs Control from above
#include "q_round.cuh"
#include "q_sync.cuh"
namespace mdpp {

// ------------------------------------------------------------------------------------------------
// string-API support (n_envs small): evaluate a transition from an ARBITRARY state id, and the bare
// update_car.  The reference's getlegalactions / get_successor_state / reward_function take state
// strings, not the env (src/env_gym.py:188-265,485-620); only the clash term reads the cars.
// ------------------------------------------------------------------------------------------------
template <int NC>
__device__ __forceinline__ View view_from_state(uint32_t sidx, const DevCfg &cfg) {
    View v;
    const uint32_t nl = sidx % cfg.n_local_nodes, t = sidx / cfg.n_local_nodes;
    v.dnidx = t % cfg.n_dest_nodes;
    v.rank = t / cfg.n_dest_nodes;
    v.pnode = nl + cfg.node_base;
    v.di = 0;
    v.dnode = cfg.dn_node[v.dnidx];
    v.dbox = cfg.dn_dbox[v.dnidx];
    v.sel = cfg.dbox_sel[v.dbox];
    v.occ9 = 0;
    v.occ18 = 0;
    uint32_t rank = v.rank;
    for (int j = NC - 1; j >= 1; j--) { // unrank the multiset of the others, largest code first
        uint32_t d = (uint32_t)j - 1u;
        auto binom = [](uint32_t n, int k) -> uint32_t {
            if (n < (uint32_t)k) return 0u;
            unsigned long long r = 1;
            for (int i = 1; i <= k; i++) r = r * (n - (uint32_t)k + (uint32_t)i) / (uint32_t)i;
            return (uint32_t)r;
        };
        while (binom(d + 1u, j) <= rank) d++;
        rank -= binom(d, j);
        const uint32_t code = d - (uint32_t)j + 1u;
        if (code) {
            const uint32_t box18 = (code - 1u) / cfg.n_dest_boxes + cfg.box_base;
            v.occ9 |= 1u << boxnum0(box18);
            v.occ18 |= 1u << box18;
        }
    }
    return v;
}

template <int NC>
__global__ void __launch_bounds__(kThreads)
query_kernel(DevCfg cfg, const uint4 *__restrict__ recs, int64_t n, int c, int kind,
             const uint32_t *__restrict__ state_idx, const uint8_t *__restrict__ action,
             uint8_t *__restrict__ legal, uint32_t *__restrict__ next_idx, float *__restrict__ reward,
             uint8_t *__restrict__ next_legal) {
    __shared__ SmemTables sm;
    load_tables(sm, cfg.tables);
    const SharedTab tab{&sm};
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint32_t sidx = state_idx[e];
    View v = view_from_state<NC>(sidx, cfg);
    const uint32_t lm = legal_mask(v, v.pnode, tab);
    if (legal) legal[e] = (uint8_t)lm;
    const uint32_t a = action ? action[e] : (uint32_t)MDPP_ACT_SKIP;
    uint32_t nidx = 0xFFFFFFFFu, lm2 = 0;
    float rw = 0.f;
    // like the reference, the successor of a move the base graph allows is defined even when the
    // legal filter would reject it; a move off the graph has no successor
    const bool exists = a < MDPP_N_ACT && (a != MDPP_A_F || (tab.info(v.pnode) & 0xFFu) != kNone);
    if (exists) {
        const uint4 rec = recs[e];
        const uint64_t cars = (uint64_t)rec.x | ((uint64_t)rec.y << 32);
        const uint32_t nn = next_node(v.pnode, a, tab);
        nidx = state_index(v, nn, cfg);
        lm2 = legal_mask(v, nn, tab);
        rw = reward_float(reward_int<NC>(v, nn, cars, c, kind, cfg, tab), kind);
    }
    if (next_idx) next_idx[e] = nidx;
    if (reward) reward[e] = rw;
    if (next_legal) next_legal[e] = (uint8_t)lm2;
}

// update_car (reference src/env_gym.py:386-441) on its own: place car c on `nodes[e]` and apply the
// destination / done logic.  nodes[e] == MDPP_NONE leaves env e alone.
template <int NC>
__global__ void __launch_bounds__(kThreads)
update_car_kernel(DevCfg cfg, uint4 *__restrict__ recs, int64_t n_envs, int c, const uint8_t *__restrict__ nodes) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_envs) return;
    const uint32_t node = nodes[e];
    if (node >= MDPP_NODES) return;
    uint4 rec = recs[e];
    uint64_t cars = (uint64_t)rec.x | ((uint64_t)rec.y << 32);
    const uint32_t f = car_field(cars, c);
    View v;
    v.di = f_di(f);
    v.dnode = cfg.car_dnode[c][v.di];
    cars = set_car_field(cars, c, make_field(f_node(f), v.di, 0, f_ghost(f))); // advance_car rewrites the field
    cars = advance_car(cars, c, v, node, cfg);
    if (f_done(f)) cars = set_car_field(cars, c, car_field(cars, c) | (1u << 8)); // a finished car stays finished
    rec.x = (uint32_t)cars;
    rec.y = (uint32_t)(cars >> 32);
    recs[e] = rec;
}

// per-state info words for the pull kernels (q_learner.cuh): everything the reference update of a key
// needs besides the two Q rows, precomputed once per handle from the state id alone.
template <int NC>
__global__ void __launch_bounds__(kThreads)
state_info_kernel(DevCfg cfg, uint32_t *__restrict__ sinfo, int64_t n_states) {
    __shared__ SmemTables sm;
    load_tables(sm, cfg.tables);
    const SharedTab tab{&sm};
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_states) return;
    const View v = view_from_state<NC>((uint32_t)s, cfg);
    const uint32_t lm = legal_mask(v, v.pnode, tab);
    const uint32_t fwd = tab.info(v.pnode) & 0xFFu;
    const uint32_t fwd_local = (fwd == kNone || fwd < (uint32_t)cfg.node_base ||
                                fwd - (uint32_t)cfg.node_base >= (uint32_t)cfg.n_local_nodes)
                                   ? 127u : fwd - (uint32_t)cfg.node_base;
    uint32_t bad = 0;
    if (cfg.bad_on) {
        const uint32_t id = (v.rank * cfg.n_dest_boxes + v.dbox) * cfg.n_src_boxes + ((v.pnode >> 2) - cfg.box_base);
        bad = (__ldg(cfg.bad_bitmap + (id >> 5)) >> (id & 31u)) & 1u;
    }
    sinfo[s] = lm | (fwd_local << 5) | (tab.bfs(v.dnidx, v.pnode) << 12) | (bad << 20) |
               ((uint32_t)(v.pnode == v.dnode) << 21) | ((v.pnode - (uint32_t)cfg.node_base) << 22);
}

// ------------------------------------------------------------------------------------------------
// DQN feed: the env-facing half of the noisy double DQN's collection loop
// (reference src/algorithms/dqn_open_source.py:490-526): encodestate -> convert(state) + give_mask for
// the network, then env.step(state, action) -> (convert(next_state), dqn_reward, mask(next_state)).
// The MLP itself stays in PyTorch.  Features are RL.convert's 80 {0,1} bytes (:245-313).
// ------------------------------------------------------------------------------------------------
constexpr int kFeat = 16 * MDPP_MAX_CARS; // 80

// 8 bytes of one car entry: box NUMBER as 5 bits MSB first, then 3 heading bits (N,W,E,S = 0..3, 7 = none)
__device__ __forceinline__ unsigned long long feat_entry(uint32_t number, uint32_t dir) {
    const uint32_t v = (number << 3) | dir;          // b7..b0 = n4 n3 n2 n1 n0 d2 d1 d0
    const uint32_t vr = __brev(v) >> 24;             // byte i of the entry = bit (7 - i) of v
    unsigned long long x = ((unsigned long long)vr * 0x0101010101010101ull) & 0x8040201008040201ull;
    return ((x + 0x7F7F7F7F7F7F7F7Full) >> 7) & 0x0101010101010101ull;
}

// Others in the reference's string order (ascend_array, src/env_gym.py:267-279): box number, then the
// source / destination strings ('D' < 'U').  keys[i] = sortable key | (src box18 << 20) | (dst box18 << 25).
template <int NC>
__device__ __forceinline__ void listed_others(uint64_t cars, int c, int mode, const DevCfg &cfg, uint32_t (&keys)[NC]) {
#pragma unroll
    for (int i = 0; i < NC; i++) {
        const uint32_t f = car_field(cars, i);
        uint32_t sb = 0, db = 0;
        bool listed = false;
        if (i != c) {
            if (f_done(f)) {
                if (mode != 0 && f_ghost(f) != 7u) { // counters already advanced by the observe kernel
                    listed = true;
                    sb = db = cfg.car_ghost_box18[i];
                }
            } else {
                listed = true;
                sb = f_node(f) >> 2;
                db = cfg.car_dbox18[i][f_di(f)];
            }
        }
        const uint32_t order = (boxnum0(sb) << 8) | ((sb < 9u) << 7) | ((db < 9u) << 6) | (boxnum0(db) << 2);
        keys[i] = listed ? ((order << 0) & 0xFFFFFu) | (sb << 20) | (db << 25) : 0xFFFFFFFFu;
    }
    // the low 20 bits order the entries; box ids ride along in the high bits -> sort on a rotated key
#pragma unroll
    for (int i = 0; i < NC; i++)
        if (keys[i] != 0xFFFFFFFFu) keys[i] = (keys[i] << 12) | (keys[i] >> 20);
    if constexpr (NC > 1) sort_codes<NC>(keys);
}

// Both kernels step through the shared step tables (fast_tabs.cuh), read in place through L1, and take the 8
// feature bytes of a (box number, heading) pair from a 256-entry shared-memory table instead of recomputing the
// bit spread per entry: ncu had dqn_act at 683 warp-instructions per 32 envs, issue slots 80 % busy and DRAM
// at 60 %, i.e. bound by instructions, next to dqn_observe at 79 % of DRAM.
struct FeatLut {
    unsigned long long e[256]; // index = number << 3 | dir
};
__device__ __forceinline__ void build_feat_lut(FeatLut &lut) {
    for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) lut.e[i] = feat_entry(i >> 3, i & 7u);
}

template <int NC>
__device__ __forceinline__ void write_features_lut(const FeatLut &lut, uint4 *dst, uint32_t pnode, uint32_t dnode,
                                                   const uint32_t (&keys)[NC]) {
    const uint32_t dirmap = 0x1320u; // our N,E,S,W (0..3) -> the reference's N=0 W=1 E=2 S=3, one nibble each
    unsigned long long src[MDPP_MAX_CARS], dstv[MDPP_MAX_CARS];
    src[0] = lut.e[((boxnum0(pnode >> 2) + 1u) << 3) | ((dirmap >> (4u * (pnode & 3u))) & 7u)];
    dstv[0] = lut.e[((boxnum0(dnode >> 2) + 1u) << 3) | ((dirmap >> (4u * (dnode & 3u))) & 7u)];
#pragma unroll
    for (int i = 1; i < MDPP_MAX_CARS; i++) {
        src[i] = 0x0101010101010101ull; // missing cars: all ones (:303)
        dstv[i] = 0x0101010101010101ull;
        if (i - 1 < NC) {
            const uint32_t k = keys[i - 1];
            if (k != 0xFFFFFFFFu) {
                const uint32_t sb = k & 31u, db = (k >> 5) & 31u; // after the rotation: bits 0-4 src, 5-9 dst
                src[i] = lut.e[((boxnum0(sb) + 1u) << 3) | 7u];
                dstv[i] = lut.e[((boxnum0(db) + 1u) << 3) | 7u];
            }
        }
    }
    unsigned long long all[10] = {src[0], src[1], src[2], src[3], src[4], dstv[0], dstv[1], dstv[2], dstv[3], dstv[4]};
#pragma unroll
    for (int j = 0; j < 5; j++)
        dst[j] = make_uint4((uint32_t)all[2 * j], (uint32_t)(all[2 * j] >> 32), (uint32_t)all[2 * j + 1],
                            (uint32_t)(all[2 * j + 1] >> 32));
}

// Stage a warp's 32 x 80 feature bytes in shared memory and write them out as 5 fully coalesced
// 512-byte rows (a lane's own 80 bytes are 80 apart from its neighbour's: direct stores would touch
// 32 different sectors per instruction).
template <int NC>
__device__ __forceinline__ void store_features_lut(const FeatLut &lut, uint8_t *__restrict__ out, int64_t warp_first_env,
                                                   int64_t n_envs, uint4 *stage, bool valid, uint32_t pnode,
                                                   uint32_t dnode, const uint32_t (&keys)[NC]) {
    const uint32_t lane = threadIdx.x & 31u;
    uint4 *mine = stage + lane * 5; // 80 bytes per lane
    if (valid) write_features_lut<NC>(lut, mine, pnode, dnode, keys);
    else {
#pragma unroll
        for (int j = 0; j < 5; j++) mine[j] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    uint4 *gout = reinterpret_cast<uint4 *>(out + warp_first_env * kFeat);
    const int64_t remaining = n_envs - warp_first_env;
    const uint32_t n16 = (uint32_t)((remaining >= 32 ? 32 : remaining) * 5);
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const uint32_t i = j * 32u + lane;
        if (i < n16) gout[i] = stage[i];
    }
    __syncwarp();
}

// observe: encodestate(car, 1) with its ghost side effect -> convert(state), give_mask(state), state id
template <int NC, int C>
__global__ void __launch_bounds__(kThreads)
dqn_observe_kernel(DevCfg cfg, uint4 *__restrict__ recs, int64_t n_envs, const FastTabs *__restrict__ tabs,
                   uint32_t rank_stride, uint8_t *__restrict__ features, uint8_t *__restrict__ legal,
                   uint32_t *__restrict__ state_idx) {
    __shared__ FeatLut lut;
    __shared__ uint4 stage[kThreads * 5];
    build_feat_lut(lut);
    __syncthreads();
    const StepTabs<MDPP_MAX_CARS> &S = *reinterpret_cast<const StepTabs<MDPP_MAX_CARS> *>(tabs);
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t warp_first = e - (threadIdx.x & 31);
    uint32_t keys[NC], pnode = 0, dnode = 0, lm = 0, sidx = 0xFFFFFFFFu;
    bool valid = false;
    if (e < n_envs) {
        uint4 rec = recs[e];
        uint64_t cars = (uint64_t)rec.x | ((uint64_t)rec.y << 32);
        if (!f_done(car_field(cars, C))) {
            const FastView v = fast_view<NC, C>(S, tabs, cars, 1, rank_stride);
            lm = v.lm;
            sidx = v.sidx;
            listed_others<NC>(v.cars, C, 2, cfg, keys);
            pnode = v.f & 127u;
            dnode = (v.own.x >> 10) & 127u;
            valid = true;
            if (v.cars != cars) { // only a ghost counter can have moved: most records are not written back (16 B/env)
                rec.x = (uint32_t)v.cars;
                rec.y = (uint32_t)(v.cars >> 32);
                recs[e] = rec;
            }
        }
        if (legal) legal[e] = (uint8_t)lm;
        if (state_idx) state_idx[e] = sidx;
    }
    if (!valid) {
#pragma unroll
        for (int i = 0; i < NC; i++) keys[i] = 0xFFFFFFFFu;
    }
    if (warp_first < n_envs && features)
        store_features_lut<NC>(lut, features, warp_first, n_envs, stage + (threadIdx.x & ~31) * 5, valid, pnode, dnode, keys);
}

// act: env.step(state, action, num) for the state the observe kernel produced (ghost counters are NOT
// advanced again) -> convert(next_state), reward, give_mask(next_state), env-done flag; optional
// end-of-round auto-reset with Philox headings when c is the last slot.
template <int NC, int C>
__global__ void __launch_bounds__(kThreads)
dqn_act_kernel(DevCfg cfg, uint4 *__restrict__ recs, int64_t n_envs, uint32_t env_id_base, int kind,
               const FastTabs *__restrict__ tabs, uint32_t rank_stride, const uint8_t *__restrict__ actions,
               uint8_t *__restrict__ next_features, float *__restrict__ reward, uint8_t *__restrict__ next_legal,
               uint8_t *__restrict__ done_out, uint8_t *__restrict__ status, int auto_reset, uint64_t round,
               uint32_t seed, Control *ctl) {
    __shared__ FeatLut lut;
    __shared__ CtaCounters cc;
    __shared__ uint4 stage[kThreads * 5];
    build_feat_lut(lut);
    if (threadIdx.x == 0) cc = CtaCounters{0u, 0u, 0u, 0u, 0, 0u};
    __syncthreads();
    const StepTabs<MDPP_MAX_CARS> &S = *reinterpret_cast<const StepTabs<MDPP_MAX_CARS> *>(tabs);
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t warp_first = e - (threadIdx.x & 31);
    uint32_t keys[NC], pnode = 0, dnode = 0;
    bool valid = false;
    StepCounts cnt;
    if (e < n_envs) {
        uint4 rec = recs[e];
        uint64_t cars = (uint64_t)rec.x | ((uint64_t)rec.y << 32);
        uint32_t episode = rec.z & 0xFFFFu, steps = rec.z >> 16;
        uint32_t st = 1, lm2 = 0;
        float rw = 0.f;
        const uint32_t a = actions[e];
        if (!f_done(car_field(cars, C)) && a != MDPP_ACT_SKIP) {
            const FastView v = fast_view<NC, C>(S, tabs, cars, 2, rank_stride); // mode 2: counters as the observe kernel left them
            st = 2;
            if (a < MDPP_N_ACT && ((v.lm >> a) & 1u)) {
                st = 0;
                const uint32_t mv = S.next[C][v.f & 0xFFu][a];
                const int32_t ri = fast_reward<NC, C>(v, mv, cars, kind, cfg);
                rw = reward_float(ri, kind);
                dnode = (v.own.x >> 10) & 127u;
                pnode = ((mv >> 9) & 1u) ? dnode : (mv & 127u); // the successor itself (a finished car is parked afterwards)
                // give_mask(next_state): the successor STATE STRING keeps the others and the destination
                lm2 = fast_legal(S.own[C][pnode | (v.f & 0x80u)], v.occ9, v.occ18);
                listed_others<NC>(cars, C, 2, cfg, keys);
                valid = true;
                const uint32_t nf = (mv & 0x1FFu) | (v.f & 0xE00u);
                cars = (cars & ~(0xFFFull << (12 * C))) | ((uint64_t)nf << (12 * C));
                steps = steps < 0xFFFFu ? steps + 1 : steps;
                cnt.steps = 1;
                cnt.reward = ri;
            }
        }
        const bool finished = all_done<NC>(cars);
        if (done_out) done_out[e] = (uint8_t)finished;
        if (auto_reset && C == NC - 1) {
            const bool capped = !finished && cfg.step_cap > 0 && steps >= (uint32_t)cfg.step_cap;
            if (finished || capped) {
                cnt.episodes = 1;
                cnt.capped = capped;
                episode = (episode + 1) & 0xFFFFu;
                steps = 0;
                cars = initial_cars<NC>(cfg, heading_word(round, 2, seed, env_id_base + (uint32_t)e));
            }
        }
        rec.x = (uint32_t)cars;
        rec.y = (uint32_t)(cars >> 32);
        rec.z = episode | (steps << 16);
        recs[e] = rec;
        if (reward) reward[e] = rw;
        if (next_legal) next_legal[e] = (uint8_t)lm2;
        if (status) status[e] = (uint8_t)st;
    }
    if (!valid) {
#pragma unroll
        for (int i = 0; i < NC; i++) keys[i] = 0xFFFFFFFFu;
    }
    if (warp_first < n_envs && next_features)
        store_features_lut<NC>(lut, next_features, warp_first, n_envs, stage + (threadIdx.x & ~31) * 5, valid, pnode, dnode, keys);
    flush_counts<false>(cnt, cc, MDPP_LEARN_FROZEN, ctl); // steps + reward_sum, no Q-updates
}

// ------------------------------------------------------------------------------------------------
// deadlock precompute: backward reachability of "all cars finished" over the JOINT state space.
// This is a new definition (the reference's -dd is the stall heuristic above): a joint
// state = per car (node, destination_index) or FINISHED; any unfinished car may take any action its
// legal filter allows (encodestate(., 0) view: finished cars are not listed); a state is live iff the
// all-finished state is reachable.  Joint index = sum_c v_c * V^c, v_c = node_local*2 + di, V-1 = finished.
// One thread per joint state and sweep; sweeps repeat until no bit changes (monotone fixpoint, so the
// in-place update order does not matter).
// ------------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(kThreads)
reach_sweep_kernel(DevCfg cfg, uint32_t *__restrict__ reach, uint64_t first, uint64_t n_joint, uint32_t *__restrict__ changed) {
    // joint states [first, n_joint): the whole space, or one rank's shard of it (mdpp_deadlock_sweep_range) -- a
    // thread writes the bit of its own state only and reads successor bits anywhere in the bitmap
    __shared__ SmemTables sm;
    load_tables(sm, cfg.tables);
    const SharedTab tab{&sm};
    const uint64_t idx = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_joint) return;
    if ((reach[idx >> 5] >> (idx & 31u)) & 1u) return;
    const uint32_t V = 2u * (uint32_t)cfg.n_local_nodes + 1u;
    uint64_t cars = 0, rest = idx, pw[NC];
    uint32_t digit[NC];
    bool all_fin = true;
    {
        uint64_t p = 1;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            pw[c] = p;
            p *= V;
            digit[c] = (uint32_t)(rest % V);
            rest /= V;
            const bool fin = digit[c] == V - 1u;
            all_fin = all_fin && fin;
            const uint32_t f = fin ? make_field((uint32_t)cfg.parking_node, 0, 1, 7)
                                   : make_field((digit[c] >> 1) + (uint32_t)cfg.node_base, digit[c] & 1u, 0, 5);
            cars |= (uint64_t)f << (12 * c);
        }
    }
    bool live = all_fin; // the goal itself
#pragma unroll
    for (int c = 0; c < NC; c++) {
        if (live || digit[c] == V - 1u) continue;
        uint64_t tmp = cars;
        const View v = view_of<NC>(tmp, c, 0, cfg);
        const uint32_t lm = legal_mask(v, v.pnode, tab);
        for (uint32_t a = 0; a < MDPP_N_ACT; a++) {
            if (!((lm >> a) & 1u)) continue;
            const uint32_t nn = next_node(v.pnode, a, tab);
            const uint32_t f = car_field(advance_car(cars, c, v, nn, cfg), c);
            const uint32_t nd = f_done(f) ? V - 1u : ((f_node(f) - (uint32_t)cfg.node_base) << 1) | f_di(f);
            const uint64_t j = idx - (uint64_t)digit[c] * pw[c] + (uint64_t)nd * pw[c];
            if ((*(volatile uint32_t *)(reach + (j >> 5)) >> (j & 31u)) & 1u) { live = true; break; }
        }
    }
    if (live) {
        atomicOr(reach + (idx >> 5), 1u << (idx & 31u));
        *changed = 1u;
    }
}

} // namespace mdpp

#include "reach.cuh"
#include "dqn_replay.cuh"
#include "mdpp_api.inl"
