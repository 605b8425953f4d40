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
// Large tables (scan mode) keep one touched bitmap per sub-step in global memory (a fire-and-forget
// atomic OR per env) and a MIRROR of the table: the pull of sub-step c scans the bitmaps of sub-steps
// c-1 and c, reads table X and writes table Y -- the update for keys marked in c, a plain copy for keys
// marked only in c-1 (where Y was left stale by the previous pull) -- and the two tables swap roles.
// One pass per sub-step, no in-place hazard, nothing proportional to the table size.
// Marks are double-buffered (by sub-step parity / by bank) so that the env kernel of sub-step
// t+1 may overlap the pull of sub-step t (programmatic dependent launch).
#pragma once
#include "fast_tabs.cuh"

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

// Per-CTA touched bitmaps of one region are stored blocked: [word / 8][cta slot][8 words].  A CTA writes whole
// 32-byte sectors; the pull kernel's CTA (256 threads = 8 words = one block) reads, for every env CTA, ONE sector
// shared by its 8 warps instead of one sector per warp and env CTA (the [cta][word] layout cost 8x the L2
// sectors and made the pull's first load phase 3 us long).
__device__ __forceinline__ size_t slice_index(uint32_t word, uint32_t cta, uint32_t cta_slots) {
    return ((size_t)(word >> 3) * cta_slots + cta) * 8u + (word & 7u);
}

struct Marks {
    uint32_t *cta_bits;  // dense: [n_ctas][words_per_cta] per-CTA touched bitmaps of the current parity
    uint32_t *cta_bits2; // dense, fused round: bitmaps of the catch-up kernel (or NULL) ...
    const uint32_t *flags2; // ... valid for the CTAs whose flag is set
    uint32_t n_ctas, words_per_cta, cta_slots;
    uint32_t *bits;      // scan: [n_keys / 32] touched bitmap of the current sub-step (bit = key = state * 8 + action)
    uint32_t *bits_prev; // scan: bitmap of the previous sub-step (keys where the pull's destination table is stale) or NULL
    uint32_t scan_words; // scan: words of one bitmap (a multiple of 4)
    uint32_t scan_magic; // scan: floor(2^32 / scan_words) + 1 (division by scan_words, see scan_slot)
    uint32_t *visited;   // [n_states] bit a = (state, a) has been updated at least once (qvalue file writer)
    const uint32_t *sinfo; // [n_states] per-state info words
    uint32_t dense;
    uint32_t n_keys;
    uint32_t n_marks; // dense: n_states * 5 -- marks are indexed by the compact key  state * 5 + action
    uint32_t l1_check; // scan: test the bitmap word through L1 before the atomic
};

// epsilon schedule (reference src/utils/utils.py:33-52, model 3) on the per-env episode counter
__device__ __forceinline__ uint32_t eps_q16_of(const LearnArgs &lp, int32_t episode) {
    uint32_t e = lp.eps_q16[5];
#pragma unroll
    for (int i = 4; i >= 0; i--)
        if (episode < lp.eps_threshold[i]) e = lp.eps_q16[i];
    return e;
}

// ------------------------------------------------------------------------------------------------
// env side of one sub-step: car slot C of every env
// ------------------------------------------------------------------------------------------------
// The env side is bound by instruction issue and by what it sends to L2 (ncu, 2^20 envs: ~430
// warp-instructions per 32 envs in the first version, issue slots 72 % busy; one scattered RED per env
// caps at ~1e11/s chip-wide however the addresses are spread), so:
//   * everything that can be decided at compile time is: the car slot is a template parameter (fixed
//     shifts, no register-indexed constant loads), the epsilon schedule has a one-compare fast path,
//     the random legal action comes from a 32-entry table, the legal filter is three mask tests on a
//     precomputed node entry, counters are reduced per CTA, and on the PLAIN path the reward -- which
//     only the pull kernel needs -- is not computed on this side at all;
//   * small tables: marks go to a byte map in SHARED memory (one plain STS per env); a persistent CTA
//     per SM compresses its map to a bitmap and writes it with coalesced stores when it is done, and the
//     pull kernel ORs the per-CTA bitmaps.  No L2 atomics, ~1 MB of mark traffic per sub-step.
//
// rng_mode: 0 = compute the Philox block here; 1 = car-0 launch: compute it for every live env and
// publish the words of cars 1..3; 2 = read this car's word from the side array.
// PLAIN = the training fast path: no forced actions, no per-env outputs, no learner flags, no reward_sum.

struct EnvArgs {
    uint4 *recs;
    int64_t n_envs;
    uint32_t env_id_base;
    uint64_t round;
    const float *q;
    const FastTabs *tabs;
    uint32_t rank_stride; // n_dest_nodes * n_local_nodes
    uint32_t n_states;    // rows of the Q table (index guard)
    Control *ctl;
    uint32_t *rng_words;
    int rng_mode;
    uint32_t order_slot;  // MDPP_LEARN_RANDOM_ORDER: index of the sub-step within the round
    const uint8_t *forced;
    uint8_t *action_out;
    uint32_t *state_out;
    float *reward_out;
    uint32_t *dl_state;
};

struct StepCounts {
    uint32_t steps = 0, explore = 0, episodes = 0, capped = 0;
    int32_t reward = 0;
    uint32_t max_episode = 0; // largest episode counter this thread has written
};

// One agent-step of car C in env e (record `rec` already loaded).  Returns true and sets `key` when the env
// updated a key; the caller owns the marking.  Writes the record back and the optional outputs.
// ANY: random car order (MDPP_LEARN_RANDOM_ORDER) -- C is the car this env drew, the forced byte, the decision
// word and the heading word of a reset come from the caller, and the episode may end after any agent-step.
template <int NC, int C, bool PLAIN, bool ANY, class Tabs>
__device__ __forceinline__ bool env_substep_car(const Tabs &S, const DevCfg &cfg, const LearnArgs &lp,
                                                const EnvArgs &ea, int64_t e, uint4 rec, uint32_t &key, StepCounts &cnt,
                                                uint32_t any_forced, uint32_t any_word, uint32_t any_heading) {
    const uint32_t lflags = PLAIN ? 0u : lp.flags;
    const uint8_t *forced = PLAIN ? nullptr : ea.forced;
    uint8_t *action_out = PLAIN ? nullptr : ea.action_out;
    uint32_t *state_out = PLAIN ? nullptr : ea.state_out;
    float *reward_out = PLAIN ? nullptr : ea.reward_out;
    bool touched = false;
    uint64_t cars = (uint64_t)rec.x | ((uint64_t)rec.y << 32);
    uint32_t episode = rec.z & 0xFFFFu, steps = rec.z >> 16;
    const uint32_t env = ea.env_id_base + (uint32_t)e;
    uint32_t a_out = MDPP_ACT_SKIP, s_out = 0xFFFFFFFFu;
    float r_out = 0.f;
    const bool live = (int32_t)episode < lp.episodes; // an env that used its episode budget is frozen
    // detect_deadlocks_v0 encodes with deadlock_id 0 (qlearning.py:342), training with 1 (:218)
    const int mode = (lflags & MDPP_LEARN_MODE0) ? 0 : (lflags & MDPP_LEARN_OBSERVED) ? 2 : 1;
    uint32_t word = 0;
    if (ANY) {
        word = any_word;
    } else if (C == 0 && ea.rng_mode == 1) { // one Philox block per env and round: word c belongs to car c
        if (live) {
            const uint4 p = philox4x32_r((uint32_t)ea.round, (uint32_t)(ea.round >> 32), 0, 0, lp.seed, env);
            word = p.x;
            if (NC > 1) ea.rng_words[e] = p.y;
            if (NC > 2) ea.rng_words[(size_t)ea.n_envs + e] = p.z;
            if (NC > 3) ea.rng_words[2 * (size_t)ea.n_envs + e] = p.w;
        }
    } else if (C > 0 && C < 4 && ea.rng_mode == 2) {
        word = ea.rng_words[(size_t)(C - 1) * ea.n_envs + e];
    }
    if (live && !f_done(car_field(cars, C))) {
        const FastView v = fast_view<NC, C>(S, ea.tabs, cars, mode, ea.rank_stride);
        const uint64_t before = cars; // the reward's clash term reads the cars before update_car
        cars = v.cars;                // the encode's ghost side effects stay whether or not the car moves
        uint32_t lm = v.lm;
        const uint32_t sidx = v.sidx;
        if (sidx >= ea.n_states) { // guard: the encoded state is off the table -- the car does not act
            guard_tripped(ea.ctl);
            lm = 0;
        }
        s_out = sidx;
        uint32_t f = ANY ? any_forced : forced ? forced[e] : (uint32_t)MDPP_ACT_SKIP;
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
                const uint32_t w = (ANY || ea.rng_mode) ? word : decision_word(ea.round, C, lp.seed, env);
                explore = (w >> 16) < eps16; // flipcoin (utils.py:255-265)
                // random.choice(legal) (qlearning.py:102): low 16 bits scaled to the legal count
                a = (S.nth[lm] >> (3u * (((w & 0xFFFFu) * (uint32_t)__popc(lm)) >> 16))) & 7u;
            }
            if (!explore) { // computeactionfromqvalues: first maximum in S,L,R,F,T order (:659-667)
                pdl_wait();  // the only Q read of this kernel: the preceding pull must be complete
                float qs[MDPP_Q_ROW];
                ldg_row(ea.q + (size_t)sidx * MDPP_Q_ROW, qs); // one 256-bit load = the whole 32-byte row
                float best = 0.f;
                bool have = false;
#pragma unroll
                for (int k = 0; k < MDPP_N_ACT; k++) {
                    bool ok = (lm >> k) & 1u;
                    if (ok && (!have || qs[k] > best)) { best = qs[k]; a = k; have = true; }
                }
            }
            cnt.explore += explore;
            // ---- stall heuristic of detect_deadlocks_v0 (qlearning.py:347-360): 3*n_cars consecutive
            // 'S' actions while epsilon == 0 flag the state; the env stops there (the reference returns)
            bool stalled = false;
            if (lflags & MDPP_LEARN_DETECT) {
                uint32_t run = rec.w & 0xFFu;
                run = (a == MDPP_A_S && eps16 == 0u) ? run + 1u : 0u;
                stalled = run == 3u * NC;
                rec.w = (rec.w & ~0xFFu) | (stalled ? 0u : run);
                if (stalled) {
                    ea.dl_state[e] = sidx;
                    rec.w = (rec.w & 0xFFFFu) | (episode << 16); // the episode of the detection (the reference returns it, :379-380)
                    episode = (uint32_t)lp.episodes; // frozen from now on
                }
            }
            if (!stalled) {
                // ---- update (qlearning.py:107-116): the key is marked; the pull kernel computes the value ----
                if (!(lflags & MDPP_LEARN_FROZEN)) {
                    key = sidx * MDPP_Q_ROW + a;
                    touched = true;
                }
                // ---- env step: successor (src/env_gym.py:256-265) + update_car (:386-441) = one table entry ----
                const uint32_t mv = S.next[C][v.f & 0xFFu][a];
                if constexpr (!PLAIN) { // q_reward (:555-620) for reward_out / reward_sum
                    const int32_t ri = fast_reward<NC, C>(v, mv, before, MDPP_REWARD_Q, cfg);
                    cnt.reward += ri;
                    r_out = (float)ri;
                }
                const uint32_t nf = (mv & 0x1FFu) | (v.f & 0xE00u);
                cars = (cars & ~(0xFFFull << (12 * C))) | ((uint64_t)nf << (12 * C));
                steps = steps < 0xFFFFu ? steps + 1 : steps;
                cnt.steps += 1;
            }
            a_out = a;
        }
    }
    if (live && (ANY || C == NC - 1)) { // end of the round-robin round (qlearning.py:233); any order: :291
        bool finished = all_done<NC>(cars);
        bool capped = !finished && cfg.step_cap > 0 && steps >= (uint32_t)cfg.step_cap;
        if (finished || capped) {
            cnt.episodes += 1;
            cnt.capped += capped;
            episode = episode + 1;
            cnt.max_episode = max(cnt.max_episode, episode);
            steps = 0;
            rec.w &= ~0xFFu; // deadlock_count = 0 per episode (qlearning.py:331)
            cars = initial_cars<NC>(cfg, ANY ? any_heading : heading_word(ea.round, 2, lp.seed, env));
        }
    }
    if (live) {
        rec.x = (uint32_t)cars;
        rec.y = (uint32_t)(cars >> 32);
        rec.z = (episode & 0xFFFFu) | (steps << 16);
        ea.recs[e] = rec;
    }
    if (action_out) action_out[e] = (uint8_t)a_out;
    if (state_out) state_out[e] = s_out;
    if (reward_out) reward_out[e] = r_out;
    return touched;
}

// The sub-step of car slot C -- or, under MDPP_LEARN_RANDOM_ORDER (train_given_state_random_order,
// qlearning.py:241-295; the host always launches slot 0 then), of the car each env draws for itself:
// random.randint(0, n - 1) (:267) = the high half of word 0 of the Philox block (round, stream 4 + order_slot)
// scaled to n_cars; a finished car means no agent-step (:268-269); word 1 is the decision word, word 2 the
// headings of the next episode.  The forced byte then carries the car in its high nibble (0xF = own draw) and
// the action in the low one (0..4, 0xE greedy, 0xF own decision): 0xFF / 0xFE keep their meaning.
template <int NC, int C, bool PLAIN, class Tabs>
__device__ __forceinline__ bool env_substep(const Tabs &S, const DevCfg &cfg, const LearnArgs &lp,
                                            const EnvArgs &ea, int64_t e, uint4 rec, uint32_t &key, StepCounts &cnt) {
    if constexpr (!PLAIN && C == 0) {
        if (lp.flags & MDPP_LEARN_RANDOM_ORDER) {
            const uint4 p = philox4x32_r((uint32_t)ea.round, (uint32_t)(ea.round >> 32), 4u + ea.order_slot, 0, lp.seed,
                                          ea.env_id_base + (uint32_t)e);
            uint32_t f = ea.forced ? ea.forced[e] : 0xFFu;
            uint32_t c = ((p.x >> 16) * (uint32_t)NC) >> 16;
            if ((f >> 4) != 0xFu) c = min(f >> 4, (uint32_t)NC - 1u);
            f = (f & 0xFu) == 0xFu ? (uint32_t)MDPP_ACT_SKIP : (f & 0xFu) == 0xEu ? (uint32_t)MDPP_ACT_GREEDY : (f & 0xFu);
            switch (c) {
            case 0: return env_substep_car<NC, 0, false, true>(S, cfg, lp, ea, e, rec, key, cnt, f, p.y, p.z);
            case 1: return env_substep_car<NC, (NC > 1 ? 1 : 0), false, true>(S, cfg, lp, ea, e, rec, key, cnt, f, p.y, p.z);
            case 2: return env_substep_car<NC, (NC > 2 ? 2 : 0), false, true>(S, cfg, lp, ea, e, rec, key, cnt, f, p.y, p.z);
            case 3: return env_substep_car<NC, (NC > 3 ? 3 : 0), false, true>(S, cfg, lp, ea, e, rec, key, cnt, f, p.y, p.z);
            default: return env_substep_car<NC, (NC > 4 ? 4 : 0), false, true>(S, cfg, lp, ea, e, rec, key, cnt, f, p.y, p.z);
            }
        }
    }
    return env_substep_car<NC, C, PLAIN, false>(S, cfg, lp, ea, e, rec, key, cnt, 0u, 0u, 0u);
}

// counters of a CTA: warp REDUX -> one shared-memory atomic per warp -> one thread adds the CTA's totals to a
// privatised global slot.  Fields of the packed words hold up to 2^16 (a persistent CTA steps <= 2^16 envs).
struct CtaCounters {
    uint32_t steps, explore, episodes, capped;
    int32_t reward;
    uint32_t max_episode;
};
template <bool PLAIN>
__device__ __forceinline__ void flush_counts(const StepCounts &c, CtaCounters &cc, uint32_t lflags, Control *ctl) {
    const uint32_t s = __reduce_add_sync(0xFFFFFFFFu, c.steps), x = __reduce_add_sync(0xFFFFFFFFu, c.explore);
    const uint32_t ep = __reduce_add_sync(0xFFFFFFFFu, c.episodes | (c.capped << 16));
    int32_t rs = 0;
    if constexpr (!PLAIN) rs = (int32_t)__reduce_add_sync(0xFFFFFFFFu, (unsigned)c.reward);
    if ((threadIdx.x & 31u) == 0) {
        if (s) atomicAdd(&cc.steps, s);
        if (x) atomicAdd(&cc.explore, x);
        if (ep & 0xFFFFu) atomicAdd(&cc.episodes, ep & 0xFFFFu);
        if (ep >> 16) atomicAdd(&cc.capped, ep >> 16);
        if (!PLAIN && rs) atomicAdd(&cc.reward, rs);
    }
    if (ep) { // some lane of this warp started a new episode (rare)
        const uint32_t me = __reduce_max_sync(0xFFFFFFFFu, c.max_episode);
        if ((threadIdx.x & 31u) == 0) atomicMax(&cc.max_episode, me);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mdpp_stats *slot = &ctl->stats[blockIdx.x & (kStatSlots - 1)];
        if (cc.steps) atomicAdd((unsigned long long *)&slot->agent_steps, (unsigned long long)cc.steps);
        if (cc.steps && !(lflags & MDPP_LEARN_FROZEN)) atomicAdd((unsigned long long *)&slot->q_updates, (unsigned long long)cc.steps);
        if (cc.explore) atomicAdd((unsigned long long *)&slot->explore_steps, (unsigned long long)cc.explore);
        if (cc.episodes) atomicAdd((unsigned long long *)&slot->episodes_done, (unsigned long long)cc.episodes);
        if (cc.capped) atomicAdd((unsigned long long *)&slot->episodes_capped, (unsigned long long)cc.capped);
        if (cc.max_episode) atomicMax(&ctl->max_episode[blockIdx.x & (kStatSlots - 1)], cc.max_episode);
        if (!PLAIN && cc.reward) atomicAdd((unsigned long long *)&slot->reward_sum, (unsigned long long)(long long)cc.reward);
    }
}

// scan mode (large tables): one CTA per 256-env tile, one touched bitmap per sub-step in global memory.  The step
// tables are read in place through L1 (a 256-env CTA cannot amortise staging 40 KB; a persistent variant with
// staged tables measured 10-40 % slower when the atomic still returned a value).
// Where a key lives in a scan-mode bitmap: the compact key m = state * 5 + action in BIT PLANES -- word m % NW, bit
// m / NW (NW = words of the bitmap).  Keys that are hot together (the 5 actions of a state, the 4 headings of a box, the
// 36-72 states that share the others and the destination) are neighbours in m; with bit = m % 32 they shared a few
// words, and atomics on one address serialise in L2.  Bit planes put neighbouring keys into neighbouring WORDS, and
// spread a hot cluster over many CTAs of the scanning pull.  Returns bit << 27 | word (NW < 2^27).
__device__ __forceinline__ uint32_t scan_slot(uint32_t key /* state * 8 + action */, uint32_t nw, uint32_t magic) {
    const uint32_t m = (key >> 3) * MDPP_N_ACT + (key & 7u);
    uint32_t b = __umulhi(m, magic); // m / nw, or one more
    uint32_t w = m - b * nw;
    if ((int32_t)w < 0) {
        b -= 1u;
        w += nw;
    }
    return (b << 27) | w;
}
__device__ __forceinline__ void scan_mark(uint32_t *__restrict__ bits, uint32_t key, uint32_t l1_check, uint32_t nw, uint32_t magic) {
    // Check through L1 first: atomics on a hot bitmap word are what L2 serialises (ncu, 4 cars: 80 us with 1.3 M
    // atomic sectors for the car slot whose envs share states, 21 us with 0.4 M for the one whose envs are spread).
    // L1 is not coherent, but the word can only be stale towards "clear" -- bits are set during this launch and were
    // cleared by a pull that completed before it -- and a stale clear bit just costs the atomic.  Without the
    // test (MDPP_L1_CHECK=0): 4 cars 60 -> 86 us per pure-exploration round, 101 -> 322 at epsilon 0.5; round 1 had
    // switched it off for the 16 MB bitmaps of 5 cars ("the extra load misses"), which with these kernels costs a
    // factor of two (epsilon 0.5: 468 against 237 us per round).
    // The result of the atomic is not used: a reduction (RED), nothing for the lane to wait for.
    const uint32_t slot = scan_slot(key, nw, magic), bit = 1u << (slot >> 27);
    uint32_t *w = bits + (slot & 0x7FFFFFFu);
    if (!l1_check || !(*w & bit)) atomicOr(w, bit);
}

template <int NC, int C, bool PLAIN>
__global__ void __launch_bounds__(kThreads)
q_env_kernel(DevCfg cfg, LearnArgs lp, EnvArgs ea, Marks mk, Control *ctl) {
    __shared__ CtaCounters cc;
    const StepTabs<MDPP_MAX_CARS> &S = *reinterpret_cast<const StepTabs<MDPP_MAX_CARS> *>(ea.tabs);
    if (threadIdx.x == 0) cc = CtaCounters{0u, 0u, 0u, 0u, 0, 0u};
    // PDL: the pull kernel that follows may be scheduled as SM slots free up (it waits for this grid to
    // complete before it looks at the marks)
    pdl_launch_dependents();
    __syncthreads();
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    StepCounts cnt;
    uint32_t key = 0;
    if (e < ea.n_envs)
        if (env_substep<NC, C, PLAIN>(S, cfg, lp, ea, e, ea.recs[e], key, cnt)) scan_mark(mk.bits, key, mk.l1_check, mk.scan_words, mk.scan_magic);
    flush_counts<PLAIN>(cnt, cc, PLAIN ? 0u : lp.flags, ctl);
    // every kernel of the chain waits for its predecessor before it exits, so that "my predecessor has
    // completed" implies "everything before it has completed" (lanes that explored never waited above)
    pdl_wait();
}

// dense mode (small tables): persistent CTAs, marks in a shared-memory byte map.  CTA b owns the warp tiles
// [b * T / G, (b + 1) * T / G) of the batch (T tiles of 32 envs, G CTAs); warp w takes every n_warps-th of
// them.  At the end the CTA compresses its map to bits and writes them to its own slice of the global bitmap
// array [G][words] with coalesced stores (all words, so nothing has to be cleared between sub-steps).
// BITS: tables of up to 1.6 M keys keep one BIT per key (shared-memory atomic OR, ~2 cycles per lane)
// instead of one byte per key (plain store) -- still no traffic to L2 until the flush.
constexpr int kEnvSmemThreads = 1024;
template <int NC, int C, bool PLAIN, bool BITS>
__global__ void __launch_bounds__(kEnvSmemThreads, 1)
q_env_smem_kernel(DevCfg cfg, LearnArgs lp, EnvArgs ea, Marks mk, Control *ctl) {
    extern __shared__ uint4 smem_dyn[];           // step tables, then the map (bytes: 16 keys per uint4; bits: 128)
    __shared__ CtaCounters cc;
    StepTabs<NC> &S = *reinterpret_cast<StepTabs<NC> *>(smem_dyn);
    uint4 *smem_map4 = smem_dyn + (sizeof(StepTabs<NC>) + 15) / 16;
    uint8_t *map = reinterpret_cast<uint8_t *>(smem_map4);
    uint32_t *map32 = reinterpret_cast<uint32_t *>(smem_map4);
    stage_step_tabs<NC, kEnvSmemThreads>(S, ea.tabs);
    const uint32_t n16 = BITS ? (mk.words_per_cta + 3u) >> 2 : (mk.n_marks + 15u) >> 4;
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) smem_map4[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) cc = CtaCounters{0u, 0u, 0u, 0u, 0, 0u};
    pdl_launch_dependents();
    __syncthreads();
    const int64_t tiles = (ea.n_envs + 31) >> 5;
    const int64_t t0 = tiles * blockIdx.x / gridDim.x, t1 = tiles * (blockIdx.x + 1) / gridDim.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    StepCounts cnt;
    int64_t t = t0 + warp;
    uint4 rec = make_uint4(0u, 0u, 0u, 0u);
    if (t < t1 && t * 32 + lane < ea.n_envs) rec = ea.recs[t * 32 + lane];
    while (t < t1) {
        const int64_t e = t * 32 + lane, tn = t + n_warps;
        uint4 rec_next = make_uint4(0u, 0u, 0u, 0u); // the next tile's records are in flight during this tile's work
        if (tn < t1 && tn * 32 + lane < ea.n_envs) rec_next = ea.recs[tn * 32 + lane];
        if (e < ea.n_envs) {
            uint32_t key = 0;
            if (env_substep<NC, C, PLAIN>(S, cfg, lp, ea, e, rec, key, cnt)) {
                const uint32_t m = (key >> 3) * MDPP_N_ACT + (key & 7u); // compact key: no marks for the padding slots
                if constexpr (BITS) atomicOr(map32 + (m >> 5), 1u << (m & 31u));
                else map[m] = 1;
            }
        }
        rec = rec_next;
        t = tn;
    }
    __syncthreads();
    // bytes -> bits: 32 keys = two uint4 -> one word; (w * 0x01020408) >> 24 packs four 0/1 bytes into a nibble
    uint32_t *out = mk.cta_bits;
    for (uint32_t w = threadIdx.x; w < mk.words_per_cta; w += blockDim.x) {
        uint32_t bits = 0;
        if constexpr (BITS) {
            bits = map32[w];
        } else {
            if (2u * w < n16) {
                const uint4 lo = smem_map4[2u * w];
                bits = ((lo.x * 0x01020408u) >> 24) | (((lo.y * 0x01020408u) >> 24) << 4) |
                       (((lo.z * 0x01020408u) >> 24) << 8) | (((lo.w * 0x01020408u) >> 24) << 12);
            }
            if (2u * w + 1u < n16) {
                const uint4 hi = smem_map4[2u * w + 1u];
                bits |= (((hi.x * 0x01020408u) >> 24) << 16) | (((hi.y * 0x01020408u) >> 24) << 20) |
                        (((hi.z * 0x01020408u) >> 24) << 24) | (((hi.w * 0x01020408u) >> 24) << 28);
            }
        }
        out[slice_index(w, blockIdx.x, mk.cta_slots)] = bits;
    }
    flush_counts<PLAIN>(cnt, cc, PLAIN ? 0u : lp.flags, ctl);
    pdl_wait();
}

// ------------------------------------------------------------------------------------------------
// pull side: the reference update of one key, from the snapshot table
// ------------------------------------------------------------------------------------------------
// row of the state the action leads to: same others, same destination, only the node term moves.  In range for
// every key an env can mark; clamped branch-free (a counted guard here cost the latency-bound pull 1 us per round)
__device__ __forceinline__ uint32_t td_successor(uint32_t s, uint32_t i /* = sinfo[s] */, uint32_t a, uint32_t n_states) {
    const uint32_t nl = si_node(i);
    const uint32_t rot = (0x20130u >> (4u * a)) & 3u;
    const uint32_t nnl = a == MDPP_A_F ? si_fwd(i) : ((nl & ~3u) | ((nl + rot) & 3u));
    return min(s - nl + nnl, n_states - 1u);
}
// the new value of a key from its state's info word, the successor's info word and table row
__device__ __forceinline__ float td_update(uint32_t i, uint32_t i2, const float (&ns)[MDPP_Q_ROW], float qa, float alpha,
                                           float discount) {
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
__device__ __forceinline__ float td_value(const float *__restrict__ q, const uint32_t *__restrict__ sinfo, uint32_t s,
                                          uint32_t i /* = sinfo[s] */, uint32_t a, float qa, float alpha, float discount,
                                          uint32_t n_states) {
    const uint32_t s2 = td_successor(s, i, a, n_states);
    const uint32_t i2 = __ldg(sinfo + s2);
    float ns[MDPP_Q_ROW];
    ldg_row(q + (size_t)s2 * MDPP_Q_ROW, ns);
    return td_update(i, i2, ns, qa, alpha, discount);
}

__device__ __forceinline__ void mark_visited(uint32_t *__restrict__ visited, uint32_t key) {
    uint32_t *vis = visited + (key >> 3);
    const uint32_t vb = 1u << (key & 7u);
    if (!(*vis & vb)) atomicOr(vis, vb);
}

// dense mode: one thread per compact key m = state * 5 + action (marks carry no padding slots), one warp per
// bitmap word.  The warp ORs the per-CTA bitmaps of its word (lane l reads CTAs l, l + 32, ...: all loads in
// flight at once), then qnew <- qold with every marked key replaced by its update (ping-pong: no in-place hazard).
constexpr int kPullUnroll = 5; // 5 x 32 lanes = 160 per-CTA bitmaps per pass
__global__ void __launch_bounds__(kThreads)
q_pull_dense_kernel(const float *__restrict__ qold, float *__restrict__ qnew, Marks mk, float alpha, float discount,
                    int early_trigger, Control *ctl) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    const uint32_t word = m >> 5; // warp-uniform
    const uint32_t s = m / MDPP_N_ACT, a = m - s * MDPP_N_ACT, key = s * MDPP_Q_ROW + a;
    // early_trigger: the kernel after this one waits (griddepcontrol.wait) before it touches anything this chain
    // produces -- another pull or the catch-up kernel -- so it may be made resident right away and its launch
    // latency disappears from the critical path.  The last pull of a round is followed by a round kernel, whose
    // prologue reads records before waiting: that one is released only once the env kernel has completed.
    if (early_trigger) pdl_launch_dependents();
    pdl_wait();              // the env kernel that wrote the bitmaps has completed
    if (!early_trigger) pdl_launch_dependents(); // the next env kernel may start; it waits for us before its first Q read
    // the state's info word is requested together with the bitmap words (it does not depend on them): a marked
    // key then needs one more round trip (successor row + info), not two
    const uint32_t info = m < mk.n_marks ? __ldg(mk.sinfo + s) : 0u;
    // All loads of a lane are issued before the first one is used: written as `bits |= load` in a loop, the
    // in-order warp waited for every load in turn (5 L2 round trips, 3 us of the kernel's 4).
    uint32_t bits = 0;
    if (word < mk.words_per_cta) {
        for (uint32_t b0 = 0; b0 < mk.n_ctas; b0 += 32u * kPullUnroll) {
            uint32_t v[kPullUnroll];
#pragma unroll
            for (int j = 0; j < kPullUnroll; j++) {
                const uint32_t b = b0 + 32u * j + lane;
                v[j] = b < mk.n_ctas ? mk.cta_bits[slice_index(word, b, mk.cta_slots)] : 0u;
            }
#pragma unroll
            for (int j = 0; j < kPullUnroll; j++) bits |= v[j];
        }
        if (mk.cta_bits2) {
            for (uint32_t b0 = 0; b0 < mk.n_ctas; b0 += 32u * kPullUnroll) {
                uint32_t v[kPullUnroll], f[kPullUnroll];
#pragma unroll
                for (int j = 0; j < kPullUnroll; j++) {
                    const uint32_t b = b0 + 32u * j + lane;
                    f[j] = b < mk.n_ctas ? mk.flags2[b] : 0u;
                    v[j] = b < mk.n_ctas ? mk.cta_bits2[slice_index(word, b, mk.cta_slots)] : 0u; // a stale slice is masked below
                }
#pragma unroll
                for (int j = 0; j < kPullUnroll; j++) bits |= f[j] ? v[j] : 0u;
            }
        }
    }
    bits = __reduce_or_sync(0xFFFFFFFFu, bits);
    uint32_t touched = 0;
    if (m < mk.n_marks) {
        float v = qold[key];
        if ((bits >> lane) & 1u) {
            v = td_value(qold, mk.sinfo, s, info, a, v, alpha, discount, mk.n_keys / MDPP_Q_ROW);
            mark_visited(mk.visited, key);
            touched = 1;
        }
        qnew[key] = v;
        if (a < MDPP_Q_ROW - MDPP_N_ACT) qnew[key + MDPP_N_ACT] = qold[key + MDPP_N_ACT]; // padding slots 5..7 carried through
    }
    touched = __reduce_add_sync(0xFFFFFFFFu, touched);
    if (lane == 0 && touched)
        atomicAdd((unsigned long long *)&ctl->stats[blockIdx.x & (kStatSlots - 1)].touched_keys,
                  (unsigned long long)touched);
}

// scan mode: the pull of one sub-step in ONE pass.  X = the table as the sub-step began (complete), Y = the other
// table, equal to X except on the keys of `bits_prev` (marked in the previous sub-step: Y is one update behind there).
// Every key marked now receives its update computed from X; every key marked only before is copied; afterwards Y is
// the complete new table and X is one update behind on the keys of `bits` -- the next pull's `bits_prev`.  X is never
// written and Y never read: no in-place hazard, no second pass, and nothing proportional to the table size except
// the scan of two bitmaps.  bits == NULL: settle (copy only; both tables equal afterwards).
// Work is balanced per KEY, not per bitmap word: hot states cluster (the 36-72 states that share the others and the
// destination are neighbours, and a popular configuration marks most of their 5 actions), and a thread that walked
// the bits of its own words one dependent load chain after the other made the first version of this kernel 60 us
// long for 20 k keys.  Each thread takes one word of each bitmap, the CTA queues the set bits in shared memory, and
// the queue is processed one key per thread, every load chain of a pass in flight at once; most CTAs find nothing
// and leave at once.  (Measured and dropped, 4-car rounds at 2^20 envs: persistent CTAs walking four words per
// thread with a batched queue -- contiguous ranges 82 us per round against 62, 32-word groups or 256-word chunks
// dealt out round-robin 72-73: the same per-key latency chain with fewer CTAs in flight.)
// The table is read past L1 (ld.global.cg): these pulls may run BESIDE a long env kernel on the same SM
// (mdpp_api.inl, q_pure_scan_rounds_launch), where a line cached by the pull before last must not be trusted.
constexpr uint32_t kScanQueue = kThreads * 32; // every bit of a CTA's words may be set
__global__ void __launch_bounds__(kThreads)
q_pull_scan_kernel(const float *__restrict__ X, float *__restrict__ Y, Marks mk, float alpha, float discount, Control *ctl) {
    __shared__ uint16_t queue[kScanQueue]; // (thread << 5 | bit) << 1 | marked in this sub-step  (16 KB)
    __shared__ uint32_t n_queued;
    if (threadIdx.x == 0) n_queued = 0;
    pdl_wait();              // the env kernel that set the marks (and, by the chain rule, the previous pull) has completed
    pdl_launch_dependents(); // the next env kernel may start; it waits for us before its first Q read
    __syncthreads();
    const uint32_t n_states = mk.n_keys / MDPP_Q_ROW;
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = 0, p = 0;
    if (w < mk.scan_words) {
        if (mk.bits) c = __ldcg(mk.bits + w);
        if (mk.bits_prev) {
            p = __ldcg(mk.bits_prev + w);
            if (p) mk.bits_prev[w] = 0; // last reader
        }
    }
    uint32_t m = c | p;
    if (m) {
        uint32_t at = atomicAdd(&n_queued, (uint32_t)__popc(m));
        while (m) {
            const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1u;
            queue[at++] = (uint16_t)(((threadIdx.x * 32u + b) << 1) | ((c >> b) & 1u));
        }
    }
    const uint32_t touched = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(c));
    if ((threadIdx.x & 31u) == 0 && touched)
        atomicAdd((unsigned long long *)&ctl->stats[blockIdx.x & (kStatSlots - 1)].touched_keys, (unsigned long long)touched);
    __syncthreads();
    const uint32_t n = n_queued;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        // bit b of word w is the compact key b * NW + w (scan_slot)
        const uint32_t item = queue[i];
        const uint32_t cm = ((item >> 1) & 31u) * mk.scan_words + blockIdx.x * kThreads + (item >> 6);
        if (cm >= mk.n_marks) { // guard: no env can mark beyond the table
            guard_tripped(ctl);
            continue;
        }
        const uint32_t key = (cm / MDPP_N_ACT) * MDPP_Q_ROW + cm % MDPP_N_ACT;
        float v = __ldcg(X + key);
        if (item & 1u) {
            const uint32_t s = key >> 3, i1 = __ldg(mk.sinfo + s);
            const uint32_t s2 = td_successor(s, i1, key & 7u, n_states);
            const uint32_t i2 = __ldg(mk.sinfo + s2);
            const float4 lo = __ldcg(reinterpret_cast<const float4 *>(X + (size_t)s2 * MDPP_Q_ROW));
            const float4 hi = __ldcg(reinterpret_cast<const float4 *>(X + (size_t)s2 * MDPP_Q_ROW) + 1);
            const float ns[MDPP_Q_ROW] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            v = td_update(i1, i2, ns, v, alpha, discount);
            mark_visited(mk.visited, key);
        }
        Y[key] = v;
    }
}

__global__ void __launch_bounds__(kThreads)
q_copy_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, uint32_t n4) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) dst[i] = src[i];
}

} // namespace mdpp
