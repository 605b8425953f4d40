// q_learner.cuh -- batched tabular Q-learning, "mark and pull" (sm_100a).
//
// Batch semantics: every env of a sub-step reads the Q table as it was when the
// sub-step began, and each touched key (state, action) receives the reference update
//     Q[s,a] + alpha * (r(s,a) + discount * max_{a' legal in s'} Q[s',a'] - Q[s,a])        (qlearning.py:107-116)
// Everything on the right-hand side is a function of the key and of the snapshot: (s, a) fixes the
// successor state string s' (same others), both bad-state flags, both BFS lengths and the goal test;
// the clash term of q_reward is 0 under legal actions.  All envs that update one key in one sub-step
// therefore compute the SAME value (the oracle counts disagreements; tests assert 0), and the work
// splits into two very different kernels:
//   * q_env_kernel   -- one lane per env: encode, legal mask, epsilon-greedy select (the only Q read,
//                       and only for lanes whose coin says "greedy"), env step, and ONE scattered
//                       byte store that marks the key as touched.  No TD arithmetic, no atomics.
//   * q_pull_*       -- one thread per KEY: recompute the update of every marked key from the
//                       snapshot ("pull") and write it; ~10^3-10^4 keys instead of ~10^6 envs.
// Small tables (dense mode) ping-pong between two table buffers: the pull kernel copies the whole
// table and substitutes the marked keys, so there is no in-place hazard and no second pass.
// Large tables (list mode) keep a touched bitmap + list and finalize in two passes (compute, apply).
// Marks / bitmaps / lists are double-buffered by sub-step parity so that the env kernel of sub-step
// t+1 may overlap the finalize of sub-step t (programmatic dependent launch).
#pragma once
#include "env_core.cuh"

namespace mdpp {

using LearnArgs = mdpp_learner;

// ---- per-state info word (workspace, built once by state_info_kernel) -----------------------------
// bits 0-4 legal mask of the state | 5-11 local node of the F successor (127 = none) | 12-19 BFS length
// to the state's destination node | 20 bad-state flag (0 when bad-states are off) | 21 primary stands on
// its destination node | 22-28 local node of the primary
constexpr uint32_t kSiBad = 1u << 20, kSiAtDest = 1u << 21;
__device__ __forceinline__ uint32_t si_lm(uint32_t i) { return i & 31u; }
__device__ __forceinline__ uint32_t si_fwd(uint32_t i) { return (i >> 5) & 127u; }
__device__ __forceinline__ int32_t si_bfs(uint32_t i) { return (int32_t)((i >> 12) & 255u); }
__device__ __forceinline__ uint32_t si_node(uint32_t i) { return (i >> 22) & 127u; }

struct Marks {
    uint32_t *words;    // dense: [n_keys], parity of the current sub-step; nonzero = key touched
    uint32_t *bits;     // list: [n_keys / 32] touched bitmap, parity of the current sub-step
    uint32_t *list;     // list: [capacity] touched keys
    uint32_t *count;    // list: number of listed keys (Control::touched_count[parity])
    uint32_t *count_other; // list: the other parity's counter (zeroed by the compute pass)
    float *vals;        // list: [capacity] new values between the compute and the apply pass
    uint32_t *visited;  // [n_states] bit a = (state, a) has been updated at least once (qvalue file writer)
    const uint32_t *sinfo; // [n_states] per-state info words
    uint32_t capacity;
    uint32_t dense;
    uint32_t n_keys;
    uint32_t replica_mask; // dense: mark replicas - 1 (power of two)
};

// epsilon schedule (reference src/utils/utils.py:33-52, model 3) on the per-env episode counter
__device__ __forceinline__ uint32_t eps_q16_of(const LearnArgs &lp, int32_t episode) {
    uint32_t e = lp.eps_q16[5];
#pragma unroll
    for (int i = 4; i >= 0; i--)
        if (episode < lp.eps_threshold[i]) e = lp.eps_q16[i];
    return e;
}

// One 16-byte table entry per node: bytes 0-7 BFS length to destination node 0..7, byte 8 F successor,
// byte 9 base legal mask.  One LDG.128 through L1; the 36 entries of a block span 5 cache lines.
struct NodeEnt {
    uint4 v;
    __device__ __forceinline__ uint32_t bfs(uint32_t d) const {
        return (uint32_t)((((unsigned long long)v.y << 32) | v.x) >> (8u * d)) & 0xFFu;
    }
    __device__ __forceinline__ uint32_t fwd() const { return v.z & 0xFFu; }
    __device__ __forceinline__ uint32_t base() const { return (v.z >> 8) & 0xFFu; }
};
__device__ __forceinline__ NodeEnt load_ent(const uint4 *__restrict__ tab, uint32_t node) {
    NodeEnt e;
    e.v = __ldg(tab + node);
    return e;
}
__device__ __forceinline__ uint32_t legal_mask_ent(const View &v, uint32_t node, const NodeEnt &ent) {
    uint32_t mask = ent.base();
    if ((v.occ9 >> boxnum0(node >> 2)) & 1u) mask &= 1u << MDPP_A_F;
    if (mask & (1u << MDPP_A_F)) {
        const uint32_t fb = ent.fwd() >> 2;
        bool blocked = (v.occ9 >> boxnum0(fb)) & 1u;
        blocked = blocked || (((v.sel >> fb) & 1u) && (v.sel & v.occ18));
        if (blocked) mask &= ~(1u << MDPP_A_F);
    }
    return mask;
}

// ------------------------------------------------------------------------------------------------
// env side of one sub-step: car slot C of every env
// ------------------------------------------------------------------------------------------------
// The kernel is bound by instruction issue (ncu: ~430 warp-instructions per 32 envs before this
// version, issue slots 72 % busy), so everything that can be decided at compile time is: the car slot
// is a template parameter (fixed shifts instead of variable ones, no register-indexed constant loads),
// the epsilon schedule has a one-compare fast path, the random legal action comes from a 32-entry
// table, counters are reduced per CTA, and on the PLAIN path the reward -- which only the pull kernel
// needs -- is not computed on this side at all.
//
// rng_mode: 0 = compute the Philox block here; 1 = car-0 launch: compute it for every live env and
// publish the words of cars 1..3; 2 = read this car's word from the side array.
// PLAIN = the training fast path: no forced actions, no per-env outputs, no learner flags, no reward_sum.

template <int NC, int C, bool PLAIN>
__global__ void __launch_bounds__(kThreads)
q_env_kernel(DevCfg cfg, uint4 *__restrict__ recs, int64_t n_envs, uint32_t env_id_base, uint64_t round,
             LearnArgs lp, const float *__restrict__ q, Marks mk, const uint4 *__restrict__ nodetab,
             uint32_t *__restrict__ rng_words, int rng_mode, const uint8_t *__restrict__ forced,
             uint8_t *__restrict__ action_out, uint32_t *__restrict__ state_out, float *__restrict__ reward_out,
             uint32_t *__restrict__ dl_state, Control *ctl) {
    __shared__ unsigned long long cta_counts; // steps | explore << 16 | episodes << 32 | capped << 48
    __shared__ long long cta_reward;
    const uint32_t lflags = PLAIN ? 0u : lp.flags;
    if constexpr (PLAIN) { forced = nullptr; action_out = nullptr; state_out = nullptr; reward_out = nullptr; }
    if (threadIdx.x == 0) { cta_counts = 0ull; cta_reward = 0ll; }
    // PDL: the pull kernel that follows may be scheduled as SM slots free up (it waits for this grid to
    // complete before it looks at the marks)
    pdl_launch_dependents();
    __syncthreads();
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t stepped = 0, explored = 0, ep_done = 0, ep_capped = 0;
    int32_t rsum = 0;
    bool first = false;   // list mode: this lane is the first toucher of its key
    uint32_t key = 0;
    if (e < n_envs) {
        uint4 rec = recs[e];
        uint64_t cars = (uint64_t)rec.x | ((uint64_t)rec.y << 32);
        uint32_t episode = rec.z & 0xFFFFu, steps = rec.z >> 16;
        const uint32_t env = env_id_base + (uint32_t)e;
        uint32_t a_out = MDPP_ACT_SKIP, s_out = 0xFFFFFFFFu;
        float r_out = 0.f;
        const bool live = (int32_t)episode < lp.episodes; // an env that used its episode budget is frozen
        // detect_deadlocks_v0 encodes with deadlock_id 0 (qlearning.py:342), training with 1 (:218)
        const int mode = (lflags & MDPP_LEARN_MODE0) ? 0 : (lflags & MDPP_LEARN_OBSERVED) ? 2 : 1;
        uint32_t word = 0;
        if (C == 0 && rng_mode == 1) { // one Philox block per env and round: word c belongs to car c
            if (live) {
                const uint4 p = philox4x32_10((uint32_t)round, (uint32_t)(round >> 32), 0, 0, lp.seed, env);
                word = p.x;
                if (NC > 1) rng_words[e] = p.y;
                if (NC > 2) rng_words[(size_t)n_envs + e] = p.z;
                if (NC > 3) rng_words[2 * (size_t)n_envs + e] = p.w;
            }
        } else if (C > 0 && C < 4 && rng_mode == 2) {
            word = rng_words[(size_t)(C - 1) * n_envs + e];
        }
        if (live && !f_done(car_field(cars, C))) {
            View v = view_of<NC>(cars, C, mode, cfg);
            const NodeEnt entP = load_ent(nodetab, v.pnode);
            const uint32_t lm = legal_mask_ent(v, v.pnode, entP);
            const uint32_t sidx = state_index(v, v.pnode, cfg);
            s_out = sidx;
            uint32_t f = forced ? forced[e] : (uint32_t)MDPP_ACT_SKIP;
            const bool forced_ok = f >= MDPP_ACT_GREEDY || (f < MDPP_N_ACT && ((lm >> f) & 1u));
            if (lm && forced_ok) {
                // ---- getaction (reference src/algorithms/qlearning.py:74-105) ----
                uint32_t a;
                bool explore;
                // epsilon schedule; the first segment (40 % of the episodes, most of the steps) costs one compare
                uint32_t eps16 = lp.eps_q16[0];
                if ((int32_t)episode >= lp.eps_threshold[0]) eps16 = eps_q16_of(lp, (int32_t)episode);
                if (f == MDPP_ACT_GREEDY) {
                    a = 0;
                    explore = false;
                } else if (f != MDPP_ACT_SKIP) {
                    a = f;
                    explore = true;
                } else {
                    const uint32_t w = rng_mode ? word : decision_word(round, C, lp.seed, env);
                    explore = (w >> 16) < eps16; // flipcoin (utils.py:255-265)
                    // random.choice(legal) (qlearning.py:102): low 16 bits scaled to the legal count
                    a = nth_set_bit_lut(cfg, lm, ((w & 0xFFFFu) * (uint32_t)__popc(lm)) >> 16);
                }
                if (!explore) { // computeactionfromqvalues: first maximum in S,L,R,F,T order (:659-667)
                    pdl_wait();  // the only Q read of this kernel: the preceding pull must be complete
                    float qs[MDPP_Q_ROW];
                    ldg_row(q + (size_t)sidx * MDPP_Q_ROW, qs); // one 256-bit load = the whole 32-byte row
                    float best = 0.f;
                    bool have = false;
#pragma unroll
                    for (int k = 0; k < MDPP_N_ACT; k++) {
                        bool ok = (lm >> k) & 1u;
                        if (ok && (!have || qs[k] > best)) { best = qs[k]; a = k; have = true; }
                    }
                }
                explored = explore;
                // ---- stall heuristic of detect_deadlocks_v0 (qlearning.py:347-360): 3*n_cars consecutive
                // 'S' actions while epsilon == 0 flag the state; the env stops there (the reference returns)
                bool stalled = false;
                if (lflags & MDPP_LEARN_DETECT) {
                    uint32_t run = rec.w & 0xFFu;
                    run = (a == MDPP_A_S && eps16 == 0u) ? run + 1u : 0u;
                    stalled = run == 3u * NC;
                    rec.w = (rec.w & ~0xFFu) | (stalled ? 0u : run);
                    if (stalled) {
                        dl_state[e] = sidx;
                        episode = (uint32_t)lp.episodes; // frozen from now on
                    }
                }
                if (!stalled) {
                    // ---- update (qlearning.py:107-116): mark the key; the pull kernel computes the value ----
                    if (!(lflags & MDPP_LEARN_FROZEN)) {
                        key = sidx * MDPP_Q_ROW + a;
                        if (mk.dense) {
                            // RED.OR into one of R CTA-indexed replicas of the mark array: a hot key is spread
                            // over R L2 addresses (same-address traffic serialises in L2, loads included)
                            atomicOr(mk.words + (blockIdx.x & mk.replica_mask) * mk.n_keys + key, 1u);
                        } else {
                            const uint32_t bit = 1u << (key & 31u);
                            first = !(atomicOr(mk.bits + (key >> 5), bit) & bit);
                        }
                    }
                    // ---- env step: successor (src/env_gym.py:256-265), update_car (:386-441) ----
                    const uint32_t rot = (0x20130u >> (4u * a)) & 3u;
                    const uint32_t nn = a == MDPP_A_F ? entP.fwd() : ((v.pnode & ~3u) | ((v.pnode + rot) & 3u));
                    if constexpr (!PLAIN) { // q_reward (:555-620) for reward_out / reward_sum
                        int32_t ri;
                        bool bad_next = false, bad_prev = false;
                        if (cfg.bad_on) { // :568-577 on the heading-stripped strings
                            const uint32_t base = (v.rank * cfg.n_dest_boxes + v.dbox) * cfg.n_src_boxes;
                            const uint32_t idn = base + ((nn >> 2) - cfg.box_base), idp = base + ((v.pnode >> 2) - cfg.box_base);
                            bad_next = (__ldg(cfg.bad_bitmap + (idn >> 5)) >> (idn & 31u)) & 1u;
                            bad_prev = (__ldg(cfg.bad_bitmap + (idp >> 5)) >> (idp & 31u)) & 1u;
                        }
                        if (bad_next) ri = -10000;
                        else if (bad_prev) ri = 0;
                        else {
                            const NodeEnt entN = load_ent(nodetab, nn);
                            ri = -10 + 50 * ((int32_t)entP.bfs(v.dnidx) - (int32_t)entN.bfs(v.dnidx));
                            if (nn == v.dnode) ri += 10010;
                            int32_t clash = 0;
#pragma unroll
                            for (int i = 0; i < NC; i++)
                                if (i != C) clash += ((f_node(car_field(cars, i)) >> 2) == (nn >> 2));
                            ri -= clash * 100000;
                        }
                        rsum = ri;
                        r_out = (float)ri;
                    }
                    cars = advance_car(cars, C, v, nn, cfg);
                    steps = steps < 0xFFFFu ? steps + 1 : steps;
                    stepped = 1;
                }
                a_out = a;
            }
        }
        if (live && C == NC - 1) { // end of the round-robin round (qlearning.py:233)
            bool finished = all_done<NC>(cars);
            bool capped = !finished && cfg.step_cap > 0 && steps >= (uint32_t)cfg.step_cap;
            if (finished || capped) {
                ep_done = 1;
                ep_capped = capped;
                episode = episode + 1;
                steps = 0;
                rec.w &= ~0xFFu; // deadlock_count = 0 per episode (qlearning.py:331)
                cars = initial_cars<NC>(cfg, heading_word(round, 2, lp.seed, env));
            }
        }
        if (live) {
            rec.x = (uint32_t)cars;
            rec.y = (uint32_t)(cars >> 32);
            rec.z = (episode & 0xFFFFu) | (steps << 16);
            recs[e] = rec;
        }
        if (action_out) action_out[e] = (uint8_t)a_out;
        if (state_out) state_out[e] = s_out;
        if (reward_out) reward_out[e] = r_out;
    }
    if (!mk.dense) { // list mode: first touchers append their key, one counter atomic per warp
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, first);
        if (bal) {
            const uint32_t lane = threadIdx.x & 31u, leader = (uint32_t)__ffs((int)bal) - 1u;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(mk.count, (uint32_t)__popc(bal));
            base = __shfl_sync(0xFFFFFFFFu, base, (int)leader);
            const uint32_t slot = base + (uint32_t)__popc(bal & ((1u << lane) - 1u));
            if (first && slot < mk.capacity) mk.list[slot] = key;
        }
    }
    // ---- counters: warp REDUX -> one shared-memory atomic per warp -> one thread flushes the CTA ----
    {
        const uint32_t packed = __reduce_add_sync(0xFFFFFFFFu, stepped | (explored << 8) | (ep_done << 16) | (ep_capped << 24));
        int32_t rs = 0;
        if constexpr (!PLAIN) rs = (int32_t)__reduce_add_sync(0xFFFFFFFFu, (unsigned)rsum);
        if ((threadIdx.x & 31u) == 0) {
            if (packed)
                atomicAdd(&cta_counts, (unsigned long long)(packed & 0xFFu) | ((unsigned long long)((packed >> 8) & 0xFFu) << 16) |
                                           ((unsigned long long)((packed >> 16) & 0xFFu) << 32) |
                                           ((unsigned long long)(packed >> 24) << 48));
            if (!PLAIN && rs) atomicAdd((unsigned long long *)&cta_reward, (unsigned long long)(long long)rs);
        }
        __syncthreads();
        if (threadIdx.x == 0 && cta_counts) {
            mdpp_stats *slot = &ctl->stats[blockIdx.x & (kStatSlots - 1)];
            const unsigned long long cc = cta_counts;
            const unsigned long long st = cc & 0xFFFFull, ex = (cc >> 16) & 0xFFFFull, ep = (cc >> 32) & 0xFFFFull, cp = cc >> 48;
            if (st) atomicAdd((unsigned long long *)&slot->agent_steps, st);
            if (st && !(lflags & MDPP_LEARN_FROZEN)) atomicAdd((unsigned long long *)&slot->q_updates, st);
            if (ex) atomicAdd((unsigned long long *)&slot->explore_steps, ex);
            if (ep) atomicAdd((unsigned long long *)&slot->episodes_done, ep);
            if (cp) atomicAdd((unsigned long long *)&slot->episodes_capped, cp);
            if (!PLAIN && cta_reward) atomicAdd((unsigned long long *)&slot->reward_sum, (unsigned long long)cta_reward);
        }
    }
    // every kernel of the chain waits for its predecessor before it exits, so that "my predecessor has
    // completed" implies "everything before it has completed" (lanes that explored never waited above)
    pdl_wait();
}

// ------------------------------------------------------------------------------------------------
// pull side: the reference update of one key, from the snapshot table
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float td_value(const float *__restrict__ q, const uint32_t *__restrict__ sinfo, uint32_t s,
                                          uint32_t a, float qa, float alpha, float discount) {
    const uint32_t i = __ldg(sinfo + s);
    const uint32_t nl = si_node(i);
    const uint32_t rot = (0x20130u >> (4u * a)) & 3u;
    const uint32_t nnl = a == MDPP_A_F ? si_fwd(i) : ((nl & ~3u) | ((nl + rot) & 3u));
    const uint32_t s2 = s - nl + nnl; // same others, same destination: only the node term moves
    const uint32_t i2 = __ldg(sinfo + s2);
    float ns[MDPP_Q_ROW];
    ldg_row(q + (size_t)s2 * MDPP_Q_ROW, ns);
    // q_reward (src/env_gym.py:555-620); clash term = 0 under legal actions
    int32_t ri;
    if (i2 & kSiBad) ri = -10000;
    else if (i & kSiBad) ri = 0;
    else ri = -10 + 50 * (si_bfs(i) - si_bfs(i2)) + ((i2 & kSiAtDest) ? 10010 : 0);
    const uint32_t lm2 = si_lm(i2);
    float nv = 0.f;
    bool have = false;
#pragma unroll
    for (int k = 0; k < MDPP_N_ACT; k++) {
        bool ok = (lm2 >> k) & 1u;
        if (ok && (!have || ns[k] > nv)) { nv = ns[k]; have = true; }
    }
    // explicit roundings: the same operation order as the reference expression, no FMA contraction
    const float target = __fadd_rn((float)ri, __fmul_rn(discount, nv));
    return __fadd_rn(qa, __fmul_rn(alpha, __fsub_rn(target, qa)));
}

__device__ __forceinline__ void mark_visited(uint32_t *__restrict__ visited, uint32_t key) {
    uint32_t *vis = visited + (key >> 3);
    const uint32_t vb = 1u << (key & 7u);
    if (!(*vis & vb)) atomicOr(vis, vb);
}

// dense mode: one thread per table word.  qnew <- qold with every marked key replaced by its update.
__global__ void __launch_bounds__(kThreads)
q_pull_dense_kernel(const float *__restrict__ qold, float *__restrict__ qnew, Marks mk, float alpha, float discount,
                    Control *ctl) {
    pdl_wait();              // the env kernel that set the marks has completed
    pdl_launch_dependents(); // the next env kernel may start; it waits for us before its first Q read
    const uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t touched = 0;
    if (key < mk.n_keys) {
        float v = qold[key];
        uint32_t marked = 0;
        if ((key & 7u) < MDPP_N_ACT) {
            for (uint32_t r = 0; r <= mk.replica_mask; r++) {
                uint32_t *w = mk.words + (size_t)r * mk.n_keys + key;
                if (*w) { marked = 1; *w = 0; }
            }
        }
        if (marked) {
            v = td_value(qold, mk.sinfo, key >> 3, key & 7u, v, alpha, discount);
            mark_visited(mk.visited, key);
            touched = 1;
        }
        qnew[key] = v;
    }
    touched = __reduce_add_sync(0xFFFFFFFFu, touched);
    if ((threadIdx.x & 31u) == 0 && touched)
        atomicAdd((unsigned long long *)&ctl->stats[blockIdx.x & (kStatSlots - 1)].touched_keys,
                  (unsigned long long)touched);
}

// list mode, pass 1: new value of every listed key into vals[] (the table is read-only here)
__global__ void __launch_bounds__(kThreads)
q_pull_list_compute_kernel(const float *__restrict__ q, Marks mk, float alpha, float discount) {
    pdl_wait();
    pdl_launch_dependents();
    const uint32_t n = min(*mk.count, mk.capacity);
    if (blockIdx.x == 0 && threadIdx.x == 0) *mk.count_other = 0; // the next sub-step appends under the other parity
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t key = mk.list[i];
        mk.vals[i] = td_value(q, mk.sinfo, key >> 3, key & 7u, __ldg(q + key), alpha, discount);
    }
}

// list mode, pass 2: write the values, clear the bitmap words (every set bit of a word is a listed key)
__global__ void __launch_bounds__(kThreads)
q_pull_list_apply_kernel(float *__restrict__ q, Marks mk, Control *ctl) {
    pdl_wait();
    pdl_launch_dependents();
    const uint32_t n = min(*mk.count, mk.capacity);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t key = mk.list[i];
        q[key] = mk.vals[i];
        mk.bits[key >> 5] = 0;
        mark_visited(mk.visited, key);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && n)
        atomicAdd((unsigned long long *)&ctl->stats[0].touched_keys, (unsigned long long)n);
}

__global__ void __launch_bounds__(kThreads)
q_copy_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, uint32_t n4) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) dst[i] = src[i];
}

} // namespace mdpp
