// q_round.cuh -- training fast path for small tables: a whole round-robin round per launch (sm_100a).
//
// Same batch semantics as q_learner.cuh (the sub-step-by-sub-step kernels there remain the general
// path and the one the parity tests drive step by step).  Two observations make a round cheaper:
//
// 1. An env only has to wait for the pull of sub-step c-1 before its car c acts if car c's coin says
//    "greedy" -- an exploring car never reads the Q table.  So the round kernel K0 steps car 0 of every
//    env (greedy reads see the table as the round began, which is the snapshot of sub-step 0) and goes
//    straight on to car 1 when car 1 explores, marking into a second map.  A lane whose car 1 must act
//    greedily is parked (pending bits in the record, decision word saved) and finished by the catch-up
//    kernel K1 after the pull of sub-step 0.  While epsilon = 1 (40 % of the episodes, most of the
//    steps) K1 finds nothing to do and its CTAs exit on one flag; later it does what the second
//    sub-step kernel did.  The pull of sub-step 1 ORs the marks of both kernels.
//
// 2. The env arithmetic is bound by the integer ALU pipe (half issue rate on sm_100: ncu shows
//    math_pipe_throttle and "wait" as the top stalls of the sub-step kernel at ~270 warp-instructions
//    per car-step).  Everything the reference derives from a car's (node, destination_index) is a
//    per-(car, node, destination_index) constant, so it is read from shared-memory tables instead:
//    one 16-byte entry for the primary (state-index base, destination node, F successor, base legal
//    mask, occupancy test masks, selective-block mask) and one 8-byte entry per other car (its listed
//    (box, destination box) code and occupancy bits).
//
// 3. At 2^20 envs half of a round is fixed cost (staging, map flush, three kernel boundaries).  While every
//    env explores, q_rounds_pure_kernel runs up to 8 rounds per launch and q_marks_reduce_kernel +
//    q_multi_pull_kernel (one cluster, table in distributed shared memory) apply all their sub-steps.
//
// Launch shape: one persistent CTA of 1024 threads per SM, tile assignment identical in K0 and K1.
#pragma once
#include "q_learner.cuh"

// The ghost of a finished car (listed for 5 encodes by the others) behind a rarely taken branch (1) or branch-free
// (0: nibble table + masked OR).  Same-box A/B of the round kernels: the branch is faster -- chained rounds 14.98 vs
// 15.81 us, one flushed round 25.7 vs 27.5, rounds with greedy decisions 30.2 vs 31.1 (the branch-free form costs
// ~25 more instructions per tile and round, and ghosts are rare).
#ifndef MDPP_GHOST_BRANCH
#define MDPP_GHOST_BRANCH 1
#endif

namespace mdpp {

// LeanTabs (fast_tabs.cuh): the step tables with a 9-bit "other" index (finished cars read a zero entry: no branch),
// the ghost entries and the rank tables staged in shared memory as well.
template <int NC>
struct RoundSmem : LeanTabs<NC> {
    CtaCounters cc;
    uint32_t any_deferred;
};

struct RoundArgs {
    uint4 *recs;
    uint32_t n_envs;
    uint32_t env_id_base;
    uint64_t round;
    const float *q;              // table a greedy car of THIS kernel reads (K0: car 0, K1: car 1)
    uint32_t *rng_words;         // [3][n_envs] decision words of parked lanes
    const FastTabs *tabs;
    uint32_t *bits_a, *bits_b;   // K0: [n_ctas][words] bitmaps of map 0 / map 1; K1: bits_a = its own map-1 bitmap
    uint32_t *defer_flags;       // [n_ctas] K0 writes, K1 reads: this CTA parked at least one lane
    uint32_t words_per_cta, cta_slots, map_bytes;
    uint32_t rank_stride;        // n_dest_nodes * n_local_nodes
    uint32_t n_states;           // rows of the Q table (index clamp)
    uint32_t tiles_q, tiles_r;   // CTA b owns tiles [b * q + min(b, r), ... + q + (b < r)): no division on the device
    uint32_t n_rounds;           // q_rounds_pure_kernel: rounds of this launch; sub-step (r, c) flushes into the mark
    uint32_t region_words;       // region r * NC + c, regions region_words apart starting at bits_a
};

// A table row for kernels under which the table changes (q_rounds_mixed_kernel): plain coherent loads, not the
// non-coherent path of ldg_row.  They may be served by L1 -- hot rows are read by thousands of lanes, and past L1
// (ld.global.cg) the same-address loads serialise in L2: 75 instead of 32 us per round -- and the acquire side of
// the grid barrier between sub-steps makes the table written before it visible to them.
__device__ __forceinline__ void ldcg_row(const float *p, float (&r)[MDPP_Q_ROW]) {
    const float4 lo = reinterpret_cast<const float4 *>(p)[0], hi = reinterpret_cast<const float4 *>(p)[1];
    r[0] = lo.x; r[1] = lo.y; r[2] = lo.z; r[3] = lo.w; r[4] = hi.x; r[5] = hi.y; r[6] = hi.z; r[7] = hi.w;
}

enum { kStepSkip = 0, kStepDone = 1, kStepDefer = 2 };

// One agent-step of car C on the unpacked record.  `cars` is committed (ghost counters of encodestate, the
// move) unless the car is parked.  Returns kStepDone with the compact key (state * 5 + action) when a key was touched.
// CG: the table changes while the kernel runs (q_rounds_mixed_kernel): coherent greedy reads (ldcg_row).
// CarsT: uint32_t for 1- and 2-car tables (both 12-bit fields live in the record's first word: no 64-bit shifts).
template <int NC, int C, bool CG = false, class CarsT = uint64_t>
__device__ __forceinline__ int round_car_step(const RoundSmem<NC> &S, const DevCfg &cfg, const FastTabs *gt,
                                              const RoundArgs &ra, CarsT &cars, uint32_t word, uint32_t eps16,
                                              bool allow_greedy, uint32_t &key, bool &moved, bool &explore,
                                              const float *q_now = nullptr) {
    static_assert(sizeof(CarsT) * 8 >= 12 * NC, "the cars do not fit the word");
    moved = false;
    explore = false;
    const uint32_t f = (uint32_t)(cars >> (12 * C)) & 0xFFFu;
    if (f & 0x100u) return kStepSkip; // `if env.check_car_training(num): continue` (qlearning.py:216-217)
    CarsT nc = cars;
    const uint4 own = S.own[C][f & 0xFFu];
    // ---- encodestate (src/env_gym.py:281-329): the others, with the ghost side effect of finished cars ----
    // branch-free (lean_view, fast_tabs.cuh): a finished car reads a zero entry; its ghost counter steps through a
    // nibble table (5..1 -> shown and one lower, 0 -> 7 = gone, 7 stays) and its entry is OR-ed in under a mask
    uint32_t occ9 = 0, occ18 = 0, codes[NC];
#pragma unroll
    for (int i = 0; i < NC; i++) {
        codes[i] = 0;
        if (i == C) continue;
        const uint32_t fi = (uint32_t)(cars >> (12 * i)) & 0xFFFu, g = fi >> 9;
        uint2 oe = S.oth[i][fi & 0x1FFu];
#if MDPP_GHOST_BRANCH
        if ((fi & 0x100u) && g != 7u) { // a finished car whose ghost is still counted: rare
            const uint32_t ng = g == 0u ? 7u : g - 1u;
            if (g != 0u) oe = S.ghost[i];
            nc = (nc & ~((CarsT)0xE00u << (12 * i))) | ((CarsT)(ng << 9) << (12 * i));
        }
#else
        const uint32_t done = (fi >> 8) & 1u;
        const uint32_t show = done & (0x7Eu >> g);
        const uint32_t ng = done ? ((0x75432107u >> (4u * g)) & 7u) : g;
        const uint32_t mask = 0u - show;
        oe.x |= S.ghost[i].x & mask;
        oe.y |= S.ghost[i].y & mask;
        nc = (nc & ~((CarsT)0xE00u << (12 * i))) | ((CarsT)(ng << 9) << (12 * i));
#endif
        codes[i] = oe.x >> 18;
        occ18 |= oe.x & 0x3FFFFu;
        occ9 |= oe.y;
    }
    uint32_t rank = 0;
    if constexpr (NC > 1) {
        sort_codes<NC>(codes);
        rank = codes[1];
        if constexpr (NC > 2) rank += (codes[2] + 1u) * codes[2] / 2u;
        if constexpr (NC > 3) rank += (codes[3] + 2u) * (codes[3] + 1u) * codes[3] / 6u;
        if constexpr (NC > 4) rank += (codes[4] + 3u) * (codes[4] + 2u) * (codes[4] + 1u) * codes[4] / 24u;
    }
    // in range by construction (codes and bases come from the tables); clamped all the same, branch-free: the mark
    // below is a shared-memory store (a counted guard like the sub-step kernels' costs this kernel 0.3 us)
    const uint32_t sidx = min(rank * ra.rank_stride + (own.x & 0x3FFu), ra.n_states - 1u);
    // ---- getlegalactions_filtered (:143-224) ----
    uint32_t lm = (own.x >> 24) & 31u;
    const uint32_t hit = (occ9 * 0x201u) & own.y;
    if (hit & 0x1FFu) lm &= 1u << MDPP_A_F;
    if ((hit >> 9) || (((own.y >> 18) & 1u) && (own.z & occ18))) lm &= ~(1u << MDPP_A_F);
    if (!lm) { // getaction returns None: nothing moves, the encode's side effects stay
        cars = nc;
        return kStepSkip;
    }
    // ---- getaction (src/algorithms/qlearning.py:74-105) ----
    explore = (word >> 16) < eps16; // flipcoin (utils.py:255-265)
    uint32_t a;
    if (explore) { // random.choice(legal): low 16 bits scaled to the legal count
        a = (S.nth[lm] >> (3u * (((word & 0xFFFFu) * (uint32_t)__popc(lm)) >> 16))) & 7u;
    } else {
        if (!allow_greedy) return kStepDefer; // the snapshot of this sub-step does not exist yet
        float qs[MDPP_Q_ROW];
        if constexpr (CG) {
            ldcg_row(q_now + (size_t)sidx * MDPP_Q_ROW, qs);
        } else {
            pdl_wait();
            ldg_row(ra.q + (size_t)sidx * MDPP_Q_ROW, qs);
        }
        float best = 0.f;
        bool have = false;
        a = 0;
#pragma unroll
        for (int k = 0; k < MDPP_N_ACT; k++) { // first maximum in S,L,R,F,T order (:659-667)
            bool ok = (lm >> k) & 1u;
            if (ok && (!have || qs[k] > best)) { best = qs[k]; a = k; have = true; }
        }
    }
    key = sidx * MDPP_N_ACT + a; // compact key (marks carry no padding slots)
    // ---- successor (src/env_gym.py:256-265) and update_car (:386-441): one table entry ----
    const uint32_t nf = ((uint32_t)S.next[C][f & 0xFFu][a] & 0x1FFu) | (f & 0xE00u);
    cars = (nc & ~((CarsT)0xFFFu << (12 * C))) | ((CarsT)nf << (12 * C));
    moved = true;
    return kStepDone;
}

// end-of-round test and reset on 32-bit cars (1- and 2-car tables)
template <int NC>
__device__ __forceinline__ bool all_done32(uint32_t cars) {
    constexpr uint32_t kDoneBits = NC == 1 ? 0x100u : 0x100100u;
    return (cars & kDoneBits) == kDoneBits;
}

// CSTART = 0: K0, every live env, cars 0..NC-1 while they explore.  CSTART = 1 (NC = 2): K1, parked lanes only.
constexpr int kRoundThreads = 1024, kRoundCtasPerSm = 1; // 2 x 768 (40 registers, spills) and 2 x 512 measured slower
// 54 registers x 1024 threads leave room for one CTA of the pull kernel (34 x 256) per SM: the pull of sub-step 0
// is then resident while this kernel runs instead of being launched as it drains (isolated round 27.1 -> 26.5 us;
// at 46 registers, room for both pulls, the kernel spills and loses 2 us).
template <int NC, int CSTART>
__global__ void __maxnreg__(54)
q_round_kernel(DevCfg cfg, LearnArgs lp, RoundArgs ra, Control *ctl) {
    static_assert(NC >= 1 && NC <= 2 && CSTART < NC, "the fused round is built for 1- and 2-car tables");
    extern __shared__ uint4 smem_dyn[];
    RoundSmem<NC> &S = *reinterpret_cast<RoundSmem<NC> *>(smem_dyn);
    constexpr uint32_t kTabBytes = (sizeof(RoundSmem<NC>) + 15u) & ~15u;
    constexpr int kMaps = NC - CSTART;
    uint8_t *maps = reinterpret_cast<uint8_t *>(smem_dyn) + kTabBytes; // kMaps byte maps of map_bytes each
    pdl_launch_dependents();
    if (CSTART > 0) {
        // This grid may be resident while K0 still runs (the pull of sub-step 0 releases it early): wait for that
        // pull -- and with it K0 -- to complete before looking at anything.  Then, if nothing was parked in this
        // CTA's tiles, leave at once.
        pdl_wait();
        if (ra.defer_flags[blockIdx.x] == 0u) return;
    }
    // the first tile's records are requested before the tables are staged (K0 reads them cold from HBM)
    const uint32_t t0 = blockIdx.x * ra.tiles_q + min(blockIdx.x, ra.tiles_r);
    const uint32_t t1 = t0 + ra.tiles_q + (blockIdx.x < ra.tiles_r ? 1u : 0u);
    constexpr uint32_t kBlock = kRoundThreads, n_warps = kRoundThreads / 32; // compile-time trip counts
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t t = t0 + warp;
    uint4 rec = make_uint4(0u, 0u, 0u, 0u);
    if (t < t1 && t * 32u + lane < ra.n_envs) rec = ra.recs[t * 32u + lane];
    {   // tables and maps
        stage_lean_tabs<NC, kBlock>(S, ra.tabs);
        uint4 *m4 = reinterpret_cast<uint4 *>(maps);
        const uint32_t n16 = (uint32_t)kMaps * (ra.map_bytes >> 4);
        for (uint32_t i = threadIdx.x; i < n16; i += kBlock) m4[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x == 0) { S.cc = CtaCounters{0u, 0u, 0u, 0u, 0, 0u}; S.any_deferred = 0u; }
    }
    __syncthreads();
    StepCounts cnt;
    bool deferred_any = false;
    while (t < t1) {
        const uint32_t e = t * 32u + lane, tn = t + n_warps;
        uint4 rec_next = make_uint4(0u, 0u, 0u, 0u); // the next tile's records are in flight during this tile's work
        if (tn < t1 && tn * 32u + lane < ra.n_envs) rec_next = ra.recs[tn * 32u + lane];
        const uint32_t pending = (rec.w >> 8) & 7u;
        uint32_t episode = rec.z & 0xFFFFu;
        const bool live = (int32_t)episode < lp.episodes; // an env that used its episode budget is frozen
        if (e < ra.n_envs && live && pending == (uint32_t)CSTART) {
            uint32_t cars = rec.x; // 1-2 cars: both 12-bit fields live in the first word
            uint32_t steps = rec.z >> 16;
            const uint32_t env = ra.env_id_base + e;
            // epsilon schedule; the first segment (40 % of the episodes, most of the steps) costs one compare
            uint32_t eps16 = lp.eps_q16[0];
            if ((int32_t)episode >= lp.eps_threshold[0]) eps16 = eps_q16_of(lp, (int32_t)episode);
            uint32_t w[NC];
            if (CSTART == 0) { // one Philox block per env and round: word c belongs to car c
                const uint4 p = philox4x32_r((uint32_t)ra.round, (uint32_t)(ra.round >> 32), 0, 0, lp.seed, env);
                w[0] = p.x;
                if (NC > 1) w[NC - 1] = p.y;
            } else {
                w[NC - 1] = ra.rng_words[e]; // saved by K0 when it parked the lane
            }
            uint32_t new_pending = 0;
            uint32_t key;
            bool moved, explore;
            if (CSTART == 0) {
                if (round_car_step<NC, 0>(S, cfg, ra.tabs, ra, cars, w[0], eps16, true, key, moved, explore) == kStepDone)
                    maps[key] = 1;
                steps += moved;
                cnt.steps += moved;
                cnt.explore += explore;
            }
            if (NC > 1) {
                const int r = round_car_step<NC, NC - 1>(S, cfg, ra.tabs, ra, cars, w[NC - 1], eps16, CSTART > 0, key, moved, explore);
                if (r == kStepDone) maps[(uint32_t)(kMaps - 1) * ra.map_bytes + key] = 1;
                if (r == kStepDefer) { // park: K1 finishes this lane after the pull of sub-step 0
                    new_pending = 1;
                    ra.rng_words[e] = w[NC - 1];
                    deferred_any = true;
                }
                steps += moved;
                cnt.steps += moved;
                cnt.explore += explore;
            }
            steps = steps > 0xFFFFu ? 0xFFFFu : steps;
            if (!new_pending) { // end of the round-robin round (qlearning.py:233)
                bool finished = all_done32<NC>(cars);
                bool capped = !finished && cfg.step_cap > 0 && steps >= (uint32_t)cfg.step_cap;
                if (finished || capped) {
                    cnt.episodes += 1;
                    cnt.capped += capped;
                    episode = episode + 1;
                    cnt.max_episode = max(cnt.max_episode, episode);
                    steps = 0;
                    rec.w &= ~0xFFu; // deadlock_count = 0 per episode (qlearning.py:331)
                    cars = (uint32_t)initial_cars<NC>(cfg, heading_word(ra.round, 2, lp.seed, env));
                }
            }
            rec.x = cars;
            rec.z = (episode & 0xFFFFu) | (steps << 16);
            rec.w = (rec.w & ~0x700u) | (new_pending << 8);
            ra.recs[e] = rec;
        }
        rec = rec_next;
        t = tn;
    }
    if (CSTART == 0 && NC > 1 && __any_sync(0xFFFFFFFFu, deferred_any) && lane == 0) S.any_deferred = 1u;
    __syncthreads();
    if (CSTART == 0 && NC > 1 && threadIdx.x == 0) ra.defer_flags[blockIdx.x] = S.any_deferred;
    // bytes -> bits: 32 keys = two uint4 -> one word; (w * 0x01020408) >> 24 packs four 0/1 bytes into a nibble
#pragma unroll
    for (int m = 0; m < kMaps; m++) {
        const uint4 *map4 = reinterpret_cast<const uint4 *>(maps + (uint32_t)m * ra.map_bytes);
        const uint32_t n16 = ra.map_bytes >> 4;
        uint32_t *out = m == 0 ? ra.bits_a : ra.bits_b;
        for (uint32_t wd = threadIdx.x; wd < ra.words_per_cta; wd += kBlock) {
            uint32_t bits = 0;
            if (2u * wd < n16) {
                const uint4 lo = map4[2u * wd];
                bits = ((lo.x * 0x01020408u) >> 24) | (((lo.y * 0x01020408u) >> 24) << 4) |
                       (((lo.z * 0x01020408u) >> 24) << 8) | (((lo.w * 0x01020408u) >> 24) << 12);
            }
            if (2u * wd + 1u < n16) {
                const uint4 hi = map4[2u * wd + 1u];
                bits |= (((hi.x * 0x01020408u) >> 24) << 16) | (((hi.y * 0x01020408u) >> 24) << 20) |
                        (((hi.z * 0x01020408u) >> 24) << 24) | (((hi.w * 0x01020408u) >> 24) << 28);
            }
            out[slice_index(wd, blockIdx.x, ra.cta_slots)] = bits;
        }
    }
    flush_counts<true>(cnt, S.cc, 0u, ctl);
    pdl_wait();
}

// ------------------------------------------------------------------------------------------------
// Several pure-exploration rounds per launch.
//
// At 2^20 envs half of a round is fixed cost: 1 us of table staging and map clearing, 1.9 us of bitmap
// compaction behind the last tile (the first map alone costs 1.1 us: first execution of that code), a kernel
// boundary -- for 10 us of stepping (globaltimer stamps).  While every env explores nothing an
// env does depends on the table, so the rounds of one mdpp_q_train_rounds call need no pull between them: this
// kernel keeps the CTA (tables staged once) and runs n_rounds rounds, each warp over its own tiles, with one
// __syncthreads per round.  The byte maps are double-buffered: while round r + 1 marks into the other pair,
// every thread compacts its two words of round r's maps into the CTA's bitmap slices of mark regions r * NC + c
// and clears them.  The pulls of all sub-steps follow the launch in order (the dense pull, one per region).
// Same per-round arithmetic as q_round_kernel<NC, 0>; no lane can be parked (epsilon = 1: every coin explores;
// the host launches this kernel only while its bound on the episode counters stays inside the first segment
// for all n_rounds rounds).
// ------------------------------------------------------------------------------------------------
template <int NC>
__device__ __forceinline__ void flush_round_maps(const uint8_t *maps, uint32_t *bits0, const RoundArgs &ra) {
    constexpr uint32_t kBlock = kRoundThreads;
#pragma unroll
    for (int m = 0; m < NC; m++) {
        uint4 *map4 = reinterpret_cast<uint4 *>(const_cast<uint8_t *>(maps) + (uint32_t)m * ra.map_bytes);
        const uint32_t n16 = ra.map_bytes >> 4;
        uint32_t *out = bits0 + (size_t)m * ra.region_words;
        for (uint32_t wd = threadIdx.x; wd < ra.words_per_cta; wd += kBlock) {
            uint32_t bits = 0;
            if (2u * wd < n16) {
                const uint4 lo = map4[2u * wd];
                bits = ((lo.x * 0x01020408u) >> 24) | (((lo.y * 0x01020408u) >> 24) << 4) |
                       (((lo.z * 0x01020408u) >> 24) << 8) | (((lo.w * 0x01020408u) >> 24) << 12);
                map4[2u * wd] = make_uint4(0u, 0u, 0u, 0u);
            }
            if (2u * wd + 1u < n16) {
                const uint4 hi = map4[2u * wd + 1u];
                bits |= (((hi.x * 0x01020408u) >> 24) << 16) | (((hi.y * 0x01020408u) >> 24) << 20) |
                        (((hi.z * 0x01020408u) >> 24) << 24) | (((hi.w * 0x01020408u) >> 24) << 28);
                map4[2u * wd + 1u] = make_uint4(0u, 0u, 0u, 0u);
            }
            out[slice_index(wd, blockIdx.x, ra.cta_slots)] = bits;
        }
    }
}

template <int NC>
__global__ void __maxnreg__(64)
q_rounds_pure_kernel(DevCfg cfg, LearnArgs lp, RoundArgs ra, Control *ctl) {
    static_assert(NC >= 1 && NC <= 2, "the fused rounds are built for 1- and 2-car tables");
    extern __shared__ uint4 smem_dyn[];
    RoundSmem<NC> &S = *reinterpret_cast<RoundSmem<NC> *>(smem_dyn);
    constexpr uint32_t kTabBytes = (sizeof(RoundSmem<NC>) + 15u) & ~15u;
    uint8_t *maps = reinterpret_cast<uint8_t *>(smem_dyn) + kTabBytes; // [2 buffers][NC maps] of map_bytes each
    pdl_launch_dependents(); // the pulls that follow wait for this grid to complete
    const uint32_t t0 = blockIdx.x * ra.tiles_q + min(blockIdx.x, ra.tiles_r);
    const uint32_t t1 = t0 + ra.tiles_q + (blockIdx.x < ra.tiles_r ? 1u : 0u);
    constexpr uint32_t kBlock = kRoundThreads, n_warps = kRoundThreads / 32;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t t = t0 + warp;
    uint4 rec = make_uint4(0u, 0u, 0u, 0u);
    if (t < t1 && t * 32u + lane < ra.n_envs) rec = ra.recs[t * 32u + lane];
    {
        stage_lean_tabs<NC, kBlock>(S, ra.tabs);
        uint4 *m4 = reinterpret_cast<uint4 *>(maps);
        const uint32_t n16 = 2u * NC * (ra.map_bytes >> 4);
        for (uint32_t i = threadIdx.x; i < n16; i += kBlock) m4[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x == 0) { S.cc = CtaCounters{0u, 0u, 0u, 0u, 0, 0u}; S.any_deferred = 0u; }
    }
    __syncthreads();
    StepCounts cnt;
    uint32_t parked = 0; // must stay 0 (see above); counted as an internal error otherwise
    for (uint32_t r = 0; r < ra.n_rounds; r++) {
        const uint64_t round = ra.round + r;
        uint8_t *rmaps = maps + (size_t)(r & 1u) * NC * ra.map_bytes;
#pragma unroll 2 // (the record hand-over rec = rec_next costs four moves per lap; 4 laps per trip measured slower)
        while (t < t1) {
            const uint32_t e = t * 32u + lane, tn = t + n_warps;
            uint4 rec_next = make_uint4(0u, 0u, 0u, 0u);
            if (tn < t1 && tn * 32u + lane < ra.n_envs) rec_next = ra.recs[tn * 32u + lane];
            uint32_t episode = rec.z & 0xFFFFu;
            const bool live = (int32_t)episode < lp.episodes;
            if (e < ra.n_envs && live && ((rec.w >> 8) & 7u) == 0u) {
                uint32_t cars = rec.x; // 1-2 cars: both 12-bit fields live in the first word
                uint32_t steps = rec.z >> 16;
                const uint32_t env = ra.env_id_base + e;
                // epsilon = 1 here by the host's bound: the coin is a compile-time "explore" (no schedule lookup, no
                // parking path); an env found outside the first segment all the same is counted as an internal error
                constexpr uint32_t kAlwaysExplore = 0x10000u;
                parked += (int32_t)episode >= lp.eps_threshold[0];
                const uint4 p = philox4x32_r((uint32_t)round, (uint32_t)(round >> 32), 0, 0, lp.seed, env);
                uint32_t key;
                bool moved, explore;
                int st = round_car_step<NC, 0>(S, cfg, ra.tabs, ra, cars, p.x, kAlwaysExplore, false, key, moved, explore);
                if (st == kStepDone) rmaps[key] = 1;
                steps += moved;
                cnt.steps += moved;
                if (NC > 1) {
                    st = round_car_step<NC, NC - 1>(S, cfg, ra.tabs, ra, cars, p.y, kAlwaysExplore, false, key, moved, explore);
                    if (st == kStepDone) rmaps[(uint32_t)(NC - 1) * ra.map_bytes + key] = 1;
                    steps += moved;
                    cnt.steps += moved;
                }
                steps = steps > 0xFFFFu ? 0xFFFFu : steps;
                { // end of the round-robin round (qlearning.py:233)
                    bool finished = all_done32<NC>(cars);
                    bool capped = !finished && cfg.step_cap > 0 && steps >= (uint32_t)cfg.step_cap;
                    if (finished || capped) {
                        cnt.episodes += 1;
                        cnt.capped += capped;
                        episode = episode + 1;
                        cnt.max_episode = max(cnt.max_episode, episode);
                        steps = 0;
                        rec.w &= ~0xFFu;
                        cars = (uint32_t)initial_cars<NC>(cfg, heading_word(round, 2, lp.seed, env));
                    }
                }
                rec.x = cars;
                rec.z = (episode & 0xFFFFu) | (steps << 16);
                ra.recs[e] = rec;
            }
            rec = rec_next;
            t = tn;
        }
        // this thread's first tile of the next round: the record it wrote above, requested before the flush work
        t = t0 + warp;
        if (r + 1u < ra.n_rounds && t < t1 && t * 32u + lane < ra.n_envs) rec = ra.recs[t * 32u + lane];
        __syncthreads(); // every mark of round r is in its maps; the other pair (round r - 1) was cleared before the previous barrier
        flush_round_maps<NC>(rmaps, ra.bits_a + (size_t)r * NC * ra.region_words, ra);
    }
    if (__any_sync(0xFFFFFFFFu, parked != 0u) && lane == 0) guard_tripped(ctl);
    cnt.explore = cnt.steps; // every decision of this kernel explores
    flush_counts<true>(cnt, S.cc, 0u, ctl);
    pdl_wait();
}

// ------------------------------------------------------------------------------------------------
// The pulls of all sub-steps of a q_rounds_pure_kernel launch: two launches instead of one dense pull (a kernel
// boundary and ~3 us) per sub-step.
//   q_marks_reduce_kernel  one warp per (sub-step, bitmap word): ORs the env CTAs' slices into one bitmap per
//                          sub-step (all sub-steps in parallel);
//   q_multi_pull_kernel    ONE thread-block cluster (16 CTAs x 512 threads) holds the table in distributed shared
//                          memory, one state row per thread, and applies the sub-steps in order.  Per sub-step a
//                          thread takes its row's 5 mark bits (prefetched a sub-step ahead), gathers the
//                          successor row of every marked action from the owning CTA's shared memory
//                          (ld.shared::cluster) and writes the new row into the other of two row buffers; one
//                          cluster barrier per sub-step separates the generations.  Successor ids and their info
//                          words do not change: computed once.  The rows go back to global memory at the end.
// Same arithmetic as the dense pull (td_update), so the table stays bit-identical.  (One CTA with the whole
// table in its shared memory was tried first: 4.8 us per sub-step -- 2400 updates are 13 k warp-instructions,
// too many for the issue slots of a single SM.)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
q_marks_reduce_kernel(const uint32_t *__restrict__ regions, uint32_t region_words, uint32_t n_regions, uint32_t words,
                      uint32_t words_per_cta, uint32_t cta_slots, uint32_t n_ctas, uint32_t *__restrict__ gbits) {
    pdl_launch_dependents();
    pdl_wait(); // the round kernel has completed
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t item = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); // region * words + word
    if (item >= n_regions * words) return;
    const uint32_t region = item / words, word = item - region * words;
    const uint32_t *src = regions + (size_t)region * region_words;
    uint32_t bits = 0;
    if (word < words_per_cta) {
        for (uint32_t b0 = 0; b0 < n_ctas; b0 += 32u * kPullUnroll) {
            uint32_t v[kPullUnroll];
#pragma unroll
            for (int j = 0; j < kPullUnroll; j++) {
                const uint32_t b = b0 + 32u * j + lane;
                v[j] = b < n_ctas ? src[slice_index(word, b, cta_slots)] : 0u;
            }
#pragma unroll
            for (int j = 0; j < kPullUnroll; j++) bits |= v[j];
        }
    }
    bits = __reduce_or_sync(0xFFFFFFFFu, bits);
    if (lane == 0) gbits[item] = bits;
}

constexpr int kMultiPullThreads = 512, kMultiPullCtas = 16; // 8192 states; 16 CTAs = the non-portable cluster size of sm_100
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void ld_cluster_row(const void *local_smem, uint32_t rank, float (&r)[MDPP_Q_ROW]) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(local_smem), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(ra) : "memory");
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4+16];"
                 : "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "r"(ra) : "memory");
}
struct MultiPullArgs {
    float *q;                 // the table, rows of MDPP_Q_ROW floats, updated in place at the end
    const uint32_t *sinfo;
    uint32_t *visited;
    const uint32_t *gbits;    // [n_substeps][words] reduced marks (compact keys)
    uint32_t n_substeps, words, n_states;
    float alpha, discount;
};
// this state's 5 mark bits of one sub-step's bitmap (compact keys s * 5 .. s * 5 + 4: at most two words)
__device__ __forceinline__ uint2 row_bit_words(const uint32_t *__restrict__ g, uint32_t s, uint32_t words) {
    const uint32_t w = (s * MDPP_N_ACT) >> 5;
    return make_uint2(w < words ? g[w] : 0u, w + 1u < words ? g[w + 1u] : 0u);
}
__global__ void __launch_bounds__(kMultiPullThreads, 1)
q_multi_pull_kernel(MultiPullArgs pa, Control *ctl) {
    __shared__ __align__(16) float rows[2][kMultiPullThreads][MDPP_Q_ROW]; // two generations of this CTA's rows
    const uint32_t rank = cluster_ctarank(), lane = threadIdx.x & 31u;
    const uint32_t s = rank * kMultiPullThreads + threadIdx.x;
    const bool valid = s < pa.n_states;
    const uint32_t sc = valid ? s : 0u, sh = (sc * MDPP_N_ACT) & 31u;
    float *row_ptr = pa.q + (size_t)sc * MDPP_Q_ROW;
    pdl_wait();              // the reduce kernel (and with it the round kernel and everything before) has completed
    pdl_launch_dependents(); // the next round kernel waits for this grid before it reads the table or exits
    const uint32_t info = pa.sinfo[sc];
    float q[MDPP_Q_ROW];
    {
        const float4 lo = reinterpret_cast<const float4 *>(row_ptr)[0], hi = reinterpret_cast<const float4 *>(row_ptr)[1];
        q[0] = lo.x; q[1] = lo.y; q[2] = lo.z; q[3] = lo.w; q[4] = hi.x; q[5] = hi.y; q[6] = hi.z; q[7] = hi.w;
    }
    uint2 nb = row_bit_words(pa.gbits, sc, pa.words);
    uint32_t s2[MDPP_N_ACT], i2[MDPP_N_ACT]; // successor rows and their info words: fixed for the whole launch
#pragma unroll
    for (int a = 0; a < MDPP_N_ACT; a++) {
        s2[a] = td_successor(sc, info, (uint32_t)a, pa.n_states);
        i2[a] = pa.sinfo[s2[a]];
    }
    reinterpret_cast<float4 *>(rows[0][threadIdx.x])[0] = make_float4(q[0], q[1], q[2], q[3]);
    reinterpret_cast<float4 *>(rows[0][threadIdx.x])[1] = make_float4(q[4], q[5], q[6], q[7]);
    cluster_arrive();
    uint32_t touched_bits = 0, touched = 0;
    for (uint32_t i = 0; i < pa.n_substeps; i++) {
        uint32_t mb = valid ? (__funnelshift_r(nb.x, nb.y, sh) & 31u) : 0u;
        if (i + 1u < pa.n_substeps) nb = row_bit_words(pa.gbits + (size_t)(i + 1u) * pa.words, sc, pa.words);
        cluster_wait(); // generation i is complete in every CTA
        float nq[MDPP_Q_ROW];
#pragma unroll
        for (int k = 0; k < MDPP_Q_ROW; k++) nq[k] = q[k];
#pragma unroll
        for (int a = 0; a < MDPP_N_ACT; a++) {
            if ((mb >> a) & 1u) {
                float ns[MDPP_Q_ROW];
                ld_cluster_row(rows[i & 1u][s2[a] % kMultiPullThreads], s2[a] / kMultiPullThreads, ns);
                nq[a] = td_update(info, i2[a], ns, q[a], pa.alpha, pa.discount);
            }
        }
#pragma unroll
        for (int k = 0; k < MDPP_Q_ROW; k++) q[k] = nq[k];
        // generation i + 1 goes to the other buffer: nobody reads that one before the next barrier completes
        reinterpret_cast<float4 *>(rows[(i + 1u) & 1u][threadIdx.x])[0] = make_float4(q[0], q[1], q[2], q[3]);
        reinterpret_cast<float4 *>(rows[(i + 1u) & 1u][threadIdx.x])[1] = make_float4(q[4], q[5], q[6], q[7]);
        cluster_arrive();
        touched_bits |= mb;
        touched += (uint32_t)__popc(mb);
    }
    if (touched_bits) {
        reinterpret_cast<float4 *>(row_ptr)[0] = make_float4(q[0], q[1], q[2], q[3]);
        reinterpret_cast<float4 *>(row_ptr)[1] = make_float4(q[4], q[5], q[6], q[7]);
        uint32_t *vis = pa.visited + s;
        if ((*vis & touched_bits) != touched_bits) atomicOr(vis, touched_bits);
    }
    touched = __reduce_add_sync(0xFFFFFFFFu, touched);
    if (lane == 0 && touched)
        atomicAdd((unsigned long long *)&ctl->stats[(rank * (kMultiPullThreads / 32) + (threadIdx.x >> 5)) & (kStatSlots - 1)].touched_keys,
                  (unsigned long long)touched);
    cluster_wait(); // no CTA leaves while its shared memory may still be read
}

// ------------------------------------------------------------------------------------------------
// Rounds WITH greedy decisions, several per launch: a cooperative persistent kernel.
//
// Once greedy decisions appear every sub-step needs the table of the sub-step before it, so a round is
// env pass -> pull -> env pass -> pull.  As separate launches that is 4 kernels and ~32 us per round at 2^20 envs,
// of which the stepping itself is ~12 (tools/exp_mixed.py; 80 % of a full training run is spent in such rounds).
// Here ONE cooperative grid (a CTA per SM, co-resident by construction) keeps its tables staged and runs
// n_rounds x n_cars sub-steps with grid-wide barriers (grid_barrier below) instead of kernel boundaries:
//   env pass of car c over the CTA's tiles (greedy reads by plain coherent loads: the table changes under the kernel), marks in
//   a shared-memory byte map, compacted into the CTA's bitmap slice and cleared;   grid barrier;
//   pull: warp w < 8 of CTA b owns bitmap word 8 b + w -- ORs the slices of all CTAs, then one thread per key writes
//   the other copy of the table (ping-pong, as q_pull_dense_kernel);                grid barrier.
// The decision word of car 1 travels through rng_words (same thread writes and reads it).  Same arithmetic,
// same order as the sub-step kernels: bit-identical tables, records and counters.
// ------------------------------------------------------------------------------------------------
// Grid-wide barrier of a cooperative launch (every CTA co-resident by construction): one arrival per CTA on a
// counter that only ever grows and a spin of one thread per CTA.  The value the counter has when a kernel starts
// lives next to it (word 1 of the barrier area): every cooperative kernel reads it first and its CTA 0 stores the
// value after its last barrier on the way out, so a kernel whose number of barriers depends on the data (the
// levels of reach_bfs_kernel) needs no host bookkeeping -- a host count that disagrees with the arrivals would
// leave the next kernel waiting for ever.  Release / acquire through the two fences and the acquiring load, like cooperative groups'
// grid.sync(), which measured ~4 us per barrier here against ~1 us for this one.
__device__ __forceinline__ void grid_barrier(uint32_t *ctr, uint32_t target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        // Everybody polls the counter itself.  (Releasing through a separate flag word written by the last arrival --
        // so that the pollers do not queue up in L2 with the atomics -- measured SLOWER: 34.0 vs 30.5 us per round of
        // q_rounds_mixed_kernel; the extra store is one more serial L2 round trip on the critical path.)
        uint32_t v, spins = 0;
        unsigned long long t0 = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            // A barrier that cannot complete (a CTA missing, a base that disagrees with the arrivals) must end in an
            // error the host sees, not in a GPU that spins until somebody resets it: after 4 s the kernel traps.
            if ((++spins & 0x3FFFu) == 0u) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ull) __trap();
            }
        } while ((int32_t)(v - target) < 0);
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ uint32_t barrier_base_load(const uint32_t *area) { return __ldcg(area + 1); }
// on the way out, after the kernel's last barrier (every CTA read the base before its first one)
__device__ __forceinline__ void barrier_base_store(uint32_t *area, uint32_t bar) {
    if (blockIdx.x == 0 && threadIdx.x == 0) area[1] = bar;
}

struct MixedArgs {
    uint32_t *barrier;        // the grid barrier's area: [0] arrival counter, [1] its value when a kernel starts
    float *q0, *q1;           // the table and its other copy: sub-step i reads q(i & 1), its pull writes q((i + 1) & 1)
    const uint32_t *sinfo;
    uint32_t *visited;
    uint32_t n_marks, n_ctas;
    float alpha, discount;
};

template <int NC, int C>
__device__ __forceinline__ void mixed_env_pass(const RoundSmem<NC> &S, const DevCfg &cfg, const LearnArgs &lp,
                                               const RoundArgs &ra, const float *q_now, uint64_t round, uint8_t *map,
                                               uint32_t t0, uint32_t t1, StepCounts &cnt) {
    constexpr uint32_t n_warps = kRoundThreads / 32;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t t = t0 + warp;
    uint4 rec = make_uint4(0u, 0u, 0u, 0u);
    if (t < t1 && t * 32u + lane < ra.n_envs) rec = ra.recs[t * 32u + lane];
    while (t < t1) {
        const uint32_t e = t * 32u + lane, tn = t + n_warps;
        uint4 rec_next = make_uint4(0u, 0u, 0u, 0u);
        if (tn < t1 && tn * 32u + lane < ra.n_envs) rec_next = ra.recs[tn * 32u + lane];
        uint32_t episode = rec.z & 0xFFFFu;
        const bool live = (int32_t)episode < lp.episodes;
        if (e < ra.n_envs && live) {
            uint32_t cars = rec.x; // 1-2 cars: both 12-bit fields live in the first word
            uint32_t steps = rec.z >> 16;
            const uint32_t env = ra.env_id_base + e;
            uint32_t eps16 = lp.eps_q16[0];
            if ((int32_t)episode >= lp.eps_threshold[0]) eps16 = eps_q16_of(lp, (int32_t)episode);
            uint32_t word;
            if (C == 0) { // one Philox block per env and round: word c belongs to car c
                const uint4 p = philox4x32_r((uint32_t)round, (uint32_t)(round >> 32), 0, 0, lp.seed, env);
                word = p.x;
                if (NC > 1) ra.rng_words[e] = p.y;
            } else {
                word = ra.rng_words[e];
            }
            uint32_t key;
            bool moved, explore;
            if (round_car_step<NC, C, true>(S, cfg, ra.tabs, ra, cars, word, eps16, true, key, moved, explore, q_now) == kStepDone)
                map[key] = 1;
            steps += moved;
            cnt.steps += moved;
            cnt.explore += explore;
            steps = steps > 0xFFFFu ? 0xFFFFu : steps;
            if (C == NC - 1) { // end of the round-robin round (qlearning.py:233)
                bool finished = all_done32<NC>(cars);
                bool capped = !finished && cfg.step_cap > 0 && steps >= (uint32_t)cfg.step_cap;
                if (finished || capped) {
                    cnt.episodes += 1;
                    cnt.capped += capped;
                    episode = episode + 1;
                    cnt.max_episode = max(cnt.max_episode, episode);
                    steps = 0;
                    rec.w &= ~0xFFu;
                    cars = (uint32_t)initial_cars<NC>(cfg, heading_word(round, 2, lp.seed, env));
                }
            }
            rec.x = cars;
            rec.z = (episode & 0xFFFFu) | (steps << 16);
            ra.recs[e] = rec;
        }
        rec = rec_next;
        t = tn;
    }
}

// the pull of one sub-step by the whole grid: warp w < 8 of CTA b owns bitmap word 8 b + w
__device__ __forceinline__ void mixed_pull(const RoundArgs &ra, const MixedArgs &mx, const float *qold, float *qnew,
                                           uint32_t &touched_total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (warp >= 8u) return;
    for (uint32_t word = blockIdx.x * 8u + warp; word < ra.words_per_cta; word += gridDim.x * 8u) {
    const uint32_t m = word * 32u + lane, s = m / MDPP_N_ACT, a = m - s * MDPP_N_ACT, key = s * MDPP_Q_ROW + a;
    const uint32_t info = m < mx.n_marks ? __ldg(mx.sinfo + s) : 0u;
    uint32_t bits = 0;
    for (uint32_t b0 = 0; b0 < mx.n_ctas; b0 += 32u * kPullUnroll) {
        uint32_t v[kPullUnroll];
#pragma unroll
        for (int j = 0; j < kPullUnroll; j++) {
            const uint32_t b = b0 + 32u * j + lane;
            v[j] = b < mx.n_ctas ? ra.bits_a[slice_index(word, b, ra.cta_slots)] : 0u;
        }
#pragma unroll
        for (int j = 0; j < kPullUnroll; j++) bits |= v[j];
    }
    bits = __reduce_or_sync(0xFFFFFFFFu, bits);
    uint32_t touched = 0;
    if (m < mx.n_marks) {
        float v = qold[key];
        if ((bits >> lane) & 1u) {
            const uint32_t s2 = td_successor(s, info, a, ra.n_states);
            const uint32_t i2 = __ldg(mx.sinfo + s2);
            float ns[MDPP_Q_ROW];
            ldcg_row(qold + (size_t)s2 * MDPP_Q_ROW, ns);
            v = td_update(info, i2, ns, v, mx.alpha, mx.discount);
            mark_visited(mx.visited, key);
            touched = 1;
        }
        qnew[key] = v;
        if (a < MDPP_Q_ROW - MDPP_N_ACT) qnew[key + MDPP_N_ACT] = qold[key + MDPP_N_ACT]; // padding slots carried through
    }
    touched = __reduce_add_sync(0xFFFFFFFFu, touched);
    if (lane == 0) touched_total += touched;
    }
}

template <int NC>
__global__ void __launch_bounds__(kRoundThreads, 1)
q_rounds_mixed_kernel(DevCfg cfg, LearnArgs lp, RoundArgs ra, MixedArgs mx, Control *ctl) {
    static_assert(NC >= 1 && NC <= 2, "built for 1- and 2-car tables");
    extern __shared__ uint4 smem_dyn[];
    RoundSmem<NC> &S = *reinterpret_cast<RoundSmem<NC> *>(smem_dyn);
    constexpr uint32_t kTabBytes = (sizeof(RoundSmem<NC>) + 15u) & ~15u;
    uint8_t *map = reinterpret_cast<uint8_t *>(smem_dyn) + kTabBytes;
    const uint32_t t0 = blockIdx.x * ra.tiles_q + min(blockIdx.x, ra.tiles_r);
    const uint32_t t1 = t0 + ra.tiles_q + (blockIdx.x < ra.tiles_r ? 1u : 0u);
    {
        stage_lean_tabs<NC, kRoundThreads>(S, ra.tabs);
        uint4 *m4 = reinterpret_cast<uint4 *>(map);
        for (uint32_t i = threadIdx.x; i < (ra.map_bytes >> 4); i += kRoundThreads) m4[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x == 0) { S.cc = CtaCounters{0u, 0u, 0u, 0u, 0, 0u}; S.any_deferred = 0u; }
    }
    __syncthreads();
    StepCounts cnt;
    uint32_t touched_total = 0, sub = 0, bar = barrier_base_load(mx.barrier);
    for (uint32_t r = 0; r < ra.n_rounds; r++) {
        const uint64_t round = ra.round + r;
#pragma unroll
        for (int c = 0; c < NC; c++, sub++) {
            const float *q_now = (sub & 1u) ? mx.q1 : mx.q0;
            float *q_next = (sub & 1u) ? mx.q0 : mx.q1;
            if (c == 0) mixed_env_pass<NC, 0>(S, cfg, lp, ra, q_now, round, map, t0, t1, cnt);
            else mixed_env_pass<NC, (NC > 1 ? 1 : 0)>(S, cfg, lp, ra, q_now, round, map, t0, t1, cnt);
            __syncthreads();
            flush_round_maps<1>(map, ra.bits_a, ra); // compacts the map into this CTA's slice and clears it
            grid_barrier(mx.barrier, bar += gridDim.x);
            mixed_pull(ra, mx, q_now, q_next, touched_total);
            grid_barrier(mx.barrier, bar += gridDim.x);
        }
    }
    if ((threadIdx.x & 31u) == 0 && touched_total)
        atomicAdd((unsigned long long *)&ctl->stats[blockIdx.x & (kStatSlots - 1)].touched_keys,
                  (unsigned long long)touched_total);
    barrier_base_store(mx.barrier, bar);
    flush_counts<true>(cnt, S.cc, 0u, ctl);
}

// ------------------------------------------------------------------------------------------------
// lockstep rollout under the uniform-random legal policy with auto-reset (config "random-policy sweep"),
// table-driven like the round kernel: same semantics as rollout_random_kernel (mdpp_kernels.cu), which stays
// as the fallback for layouts whose BFS lengths do not fit the 6 bits of the move table.
// Persistent CTAs: the tables are staged once per CTA, each warp walks its tiles, a tile's records stay in
// registers for all `rounds` rounds of the launch.
// ------------------------------------------------------------------------------------------------
template <int NC, int C>
__device__ __forceinline__ bool rollout_car_step(const RoundSmem<NC> &S, const DevCfg &cfg, const FastTabs *gt,
                                                 uint32_t rank_stride, int kind, uint64_t &cars, uint32_t word,
                                                 uint32_t &a_out, int32_t &ri, uint32_t &nidx) {
    const uint32_t f = (uint32_t)(cars >> (12 * C)) & 0xFFFu;
    if (f & 0x100u) return false;
    uint64_t nc = cars;
    const uint4 own = S.own[C][f & 0xFFu];
    uint32_t occ9 = 0, occ18 = 0, codes[NC], clash_boxes[NC];
#pragma unroll
    for (int i = 0; i < NC; i++) {
        codes[i] = 0;
        clash_boxes[i] = 0xFFFFu;
        if (i == C) continue;
        const uint32_t fi = (uint32_t)(cars >> (12 * i)) & 0xFFFu;
        clash_boxes[i] = (fi & 127u) >> 2; // real position, parked cars included (src/env_gym.py:613-619)
        uint2 oe = make_uint2(0u, 0u);
        if (!(fi & 0x100u)) {
            oe = S.oth[i][fi & 0xFFu];
        } else {
            uint32_t g = fi >> 9;
            if (g != 7u) {
                if (g == 0u) g = 7u;
                else { g -= 1u; oe = gt->ghost[i]; }
                nc = (nc & ~(0xE00ull << (12 * i))) | ((uint64_t)(g << 9) << (12 * i));
            }
        }
        codes[i] = oe.x >> 18;
        occ18 |= oe.x & 0x3FFFFu;
        occ9 |= oe.y;
    }
    cars = nc; // the encode's ghost side effects stay whether or not the car moves
    uint32_t rank = 0;
    if constexpr (NC > 1) {
        sort_codes<NC>(codes);
        rank = codes[1];
        if constexpr (NC > 2) rank += (codes[2] + 1u) * codes[2] / 2u;
        if constexpr (NC > 3) rank += (codes[3] + 2u) * (codes[3] + 1u) * codes[3] / 6u;
        if constexpr (NC > 4) rank += (codes[4] + 3u) * (codes[4] + 2u) * (codes[4] + 1u) * codes[4] / 24u;
    }
    uint32_t lm = (own.x >> 24) & 31u;
    const uint32_t hit = (occ9 * 0x201u) & own.y;
    if (hit & 0x1FFu) lm &= 1u << MDPP_A_F;
    if ((hit >> 9) || (((own.y >> 18) & 1u) && (own.z & occ18))) lm &= ~(1u << MDPP_A_F);
    if (!lm) return false;
    const uint32_t a = (S.nth[lm] >> (3u * (((word & 0xFFFFu) * (uint32_t)__popc(lm)) >> 16))) & 7u;
    const uint32_t mv = S.next[C][f & 0xFFu][a];
    const uint32_t pnode = f & 127u, dnode = (own.x >> 10) & 127u;
    const bool reached = (mv >> 9) & 1u;
    const uint32_t nn = reached ? dnode : (mv & 127u); // the successor itself (a finished car is parked afterwards)
    const uint32_t sidx = rank * rank_stride + (own.x & 0x3FFu);
    nidx = sidx - pnode + nn;                           // same others, same destination: only the node term moves
    // ---- q_reward / dqn_reward (src/env_gym.py:485-620), integer (dqn: before the /100) ----
    bool bad_next = false, bad_prev = false;
    if (kind == MDPP_REWARD_Q && cfg.bad_on) { // :568-577 on the heading-stripped strings
        const uint32_t base = (rank * cfg.n_dest_boxes + (own.x >> 29)) * cfg.n_src_boxes;
        const uint32_t idn = base + ((nn >> 2) - cfg.box_base), idp = base + ((pnode >> 2) - cfg.box_base);
        bad_next = (__ldg(cfg.bad_bitmap + (idn >> 5)) >> (idn & 31u)) & 1u;
        bad_prev = (__ldg(cfg.bad_bitmap + (idp >> 5)) >> (idp & 31u)) & 1u;
    }
    if (bad_next) ri = -10000;
    else if (bad_prev) ri = 0;
    else {
        ri = -10 + 50 * ((int32_t)(own.w & 0xFFu) - (int32_t)(mv >> 10));
        if (reached) ri += (kind == MDPP_REWARD_DQN) ? 110 : 10010;
        int32_t clash = 0;
#pragma unroll
        for (int i = 0; i < NC; i++)
            if (i != C) clash += (clash_boxes[i] == (nn >> 2));
        ri -= clash * ((kind == MDPP_REWARD_DQN) ? 100 : 100000);
    }
    const uint32_t nf = (mv & 0x1FFu) | (f & 0xE00u);
    cars = (cars & ~(0xFFFull << (12 * C))) | ((uint64_t)nf << (12 * C));
    a_out = a;
    return true;
}

struct RolloutOut {
    uint8_t *action;
    float *reward;
    uint8_t *done;
    uint32_t *next;
};

template <int NC, int C>
__device__ __forceinline__ void rollout_chain(const RoundSmem<NC> &S, const DevCfg &cfg, const FastTabs *gt,
                                              uint32_t rank_stride, int kind, uint64_t &cars, const uint32_t (&w)[NC],
                                              uint32_t &steps, StepCounts &cnt, long long &rsum, const RolloutOut &out,
                                              size_t o, size_t n_envs) {
    uint32_t a = MDPP_ACT_SKIP, nidx = 0xFFFFFFFFu;
    int32_t ri = 0;
    float rw = 0.f;
    if (rollout_car_step<NC, C>(S, cfg, gt, rank_stride, kind, cars, w[C], a, ri, nidx)) {
        rw = reward_float(ri, kind);
        steps = steps < 0xFFFFu ? steps + 1 : steps;
        cnt.steps++;
        rsum += ri;
    } else {
        a = MDPP_ACT_SKIP;
        nidx = 0xFFFFFFFFu;
    }
    if (out.action) out.action[o] = (uint8_t)a;
    if (out.reward) out.reward[o] = rw;
    if (out.next) out.next[o] = nidx;
    if (out.done) out.done[o] = (uint8_t)all_done<NC>(cars);
    if constexpr (C + 1 < NC) rollout_chain<NC, C + 1>(S, cfg, gt, rank_stride, kind, cars, w, steps, cnt, rsum, out, o + n_envs, n_envs);
}

constexpr int kRolloutThreads = 256;
template <int NC>
__global__ void __launch_bounds__(kRolloutThreads, 4)
rollout_fast_kernel(DevCfg cfg, uint4 *__restrict__ recs, uint32_t n_envs, uint32_t env_id_base, int rounds,
                    uint64_t first_round, uint32_t seed, int kind, RolloutOut out, const FastTabs *__restrict__ tabs,
                    uint32_t rank_stride, Control *ctl) {
    extern __shared__ uint4 smem_dyn[]; // RoundSmem<NC> (50 KB for 5 cars: above the static limit)
    RoundSmem<NC> &S = *reinterpret_cast<RoundSmem<NC> *>(smem_dyn);
    __shared__ long long cta_reward;
    {
        stage_lean_tabs<NC, kRolloutThreads>(S, tabs);
        if (threadIdx.x == 0) { S.cc = CtaCounters{0u, 0u, 0u, 0u, 0, 0u}; cta_reward = 0ll; }
    }
    __syncthreads();
    const uint32_t tiles = (n_envs + 31u) >> 5;
    const uint32_t t0 = (uint32_t)((uint64_t)tiles * blockIdx.x / gridDim.x);
    const uint32_t t1 = (uint32_t)((uint64_t)tiles * (blockIdx.x + 1u) / gridDim.x);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    StepCounts cnt;
    long long rsum = 0;
    for (uint32_t t = t0 + warp; t < t1; t += n_warps) {
        const uint32_t e = t * 32u + lane;
        if (e >= n_envs) continue;
        uint4 rec = recs[e];
        uint64_t cars = (uint64_t)rec.x | ((uint64_t)rec.y << 32);
        uint32_t episode = rec.z & 0xFFFFu, steps = rec.z >> 16;
        const uint32_t env = env_id_base + e;
        for (int r = 0; r < rounds; r++) {
            const uint64_t round = first_round + (uint64_t)r;
            const uint4 p = philox4x32_r((uint32_t)round, (uint32_t)(round >> 32), 0, 0, seed, env);
            uint32_t w[NC];
            w[0] = p.x;
            if (NC > 1) w[NC > 1 ? 1 : 0] = p.y;
            if (NC > 2) w[NC > 2 ? 2 : 0] = p.z;
            if (NC > 3) w[NC > 3 ? 3 : 0] = p.w;
            if (NC > 4) w[NC > 4 ? 4 : 0] = philox4x32_r((uint32_t)round, (uint32_t)(round >> 32), 1, 0, seed, env).x;
            rollout_chain<NC, 0>(S, cfg, tabs, rank_stride, kind, cars, w, steps, cnt, rsum, out,
                                 (size_t)r * NC * (size_t)n_envs + e, (size_t)n_envs);
            // end of round: is_game_ended (src/env_gym.py:622-638) or the step cap -> next episode
            const bool finished = all_done<NC>(cars);
            const bool capped = !finished && cfg.step_cap > 0 && steps >= (uint32_t)cfg.step_cap;
            if (finished || capped) {
                cnt.episodes++;
                cnt.capped += capped;
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
    // counters (every step of this kernel is an exploration step; rewards need 64 bits over many rounds)
    cnt.explore = cnt.steps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xFFFFFFFFu, rsum, o);
    if (lane == 0 && rsum) atomicAdd((unsigned long long *)&cta_reward, (unsigned long long)rsum);
    flush_counts<true>(cnt, S.cc, MDPP_LEARN_FROZEN, ctl); // no Q-updates on this path
    if (threadIdx.x == 0 && cta_reward)
        atomicAdd((unsigned long long *)&ctl->stats[blockIdx.x & (kStatSlots - 1)].reward_sum, (unsigned long long)cta_reward);
}

// ------------------------------------------------------------------------------------------------
// The same rollout on the lean step (fast_tabs.cuh, LeanTabs): cars unpacked in registers for the whole launch,
// branch-free others / ghosts, table-driven rank, and no clash term in the reward -- under LEGAL actions it is
// always 0 (a listed car blocks its box number, a finished car stands on the parking node, which is no box), which
// is what the learner's pull relies on too; the oracle comparison of every reward (1-5 cars) pins it.
// ------------------------------------------------------------------------------------------------
// ALL: every output array is present (the sweep's configuration): no per-store null tests, and the four addresses
// advance by pointer increments instead of being rebuilt from a 64-bit index.
template <int NC, int C, bool ALL>
__device__ __forceinline__ void lean_rollout_chain(const LeanTabs<NC> &S, const DevCfg &cfg, uint32_t rank_stride, int kind,
                                                   uint32_t (&f)[NC], const uint32_t (&w)[NC], uint32_t &steps, uint32_t &n_steps,
                                                   int32_t &rsum, RolloutOut &out, size_t n_envs) {
    uint32_t a = MDPP_ACT_SKIP, nidx = 0xFFFFFFFFu;
    float rw = 0.f;
    const uint32_t fc = f[C];
    if (!(fc & 0x100u)) {
        const LeanView v = lean_view<NC, C>(S, f, rank_stride);
        // bad-state bits of this state's (others, destination box) for EVERY primary box: n_src_boxes (<= 18)
        // consecutive bits.  The two words are requested here, as soon as the rank is known, and used after the
        // action has been chosen: issued next to their use they were the kernel's hottest stall (4 x 4 % of the
        // samples on the first use of each loaded word).
        uint32_t bad_window = 0;
        if (kind == MDPP_REWARD_Q && cfg.bad_on) {
            const uint32_t base = (v.rank * cfg.n_dest_boxes + (v.own.x >> 29)) * cfg.n_src_boxes;
            const uint32_t *wp = cfg.bad_bitmap + (base >> 5); // the bitmap carries one word of padding
            bad_window = __funnelshift_r(__ldg(wp), __ldg(wp + 1), base & 31u);
        }
        if (v.lm) {
            a = (S.nth[v.lm] >> (3u * (((w[C] & 0xFFFFu) * (uint32_t)__popc(v.lm)) >> 16))) & 7u;
            const uint32_t mv = S.next[C][fc & 0xFFu][a];
            const uint32_t pnode = fc & 127u, dnode = (v.own.x >> 10) & 127u;
            const bool reached = (mv >> 9) & 1u;
            const uint32_t nn = reached ? dnode : (mv & 127u); // the successor itself (a finished car is parked afterwards)
            nidx = v.sidx - pnode + nn;                         // same others, same destination: only the node term moves
            // ---- q_reward / dqn_reward (src/env_gym.py:485-620), integer (dqn: before the /100) ----
            int32_t ri = -10 + 50 * ((int32_t)(v.own.w & 0xFFu) - (int32_t)(mv >> 10));
            if (reached) ri += (kind == MDPP_REWARD_DQN) ? 110 : 10010;
            if (kind == MDPP_REWARD_Q && cfg.bad_on) { // :568-577 on the heading-stripped strings
                const uint32_t bn = (bad_window >> ((nn >> 2) - cfg.box_base)) & 1u;
                const uint32_t bp = (bad_window >> ((pnode >> 2) - cfg.box_base)) & 1u;
                ri = bn ? -10000 : (bp ? 0 : ri);
            }
            rw = reward_float(ri, kind);
            f[C] = (mv & 0x1FFu) | (fc & 0xE00u);
            steps += 1;
            n_steps += 1;
            rsum += ri;
        }
    }
    if (ALL || out.action) *out.action = (uint8_t)a;
    if (ALL || out.reward) *out.reward = rw;
    if (ALL || out.next) *out.next = nidx;
    if (ALL || out.done) {
        uint32_t d = f[0];
#pragma unroll
        for (int i = 1; i < NC; i++) d &= f[i];
        *out.done = (uint8_t)((d >> 8) & 1u);
    }
    // the next car-step's slot of every output: one row of n_envs further
    if (ALL || out.action) out.action += n_envs;
    if (ALL || out.reward) out.reward += n_envs;
    if (ALL || out.next) out.next += n_envs;
    if (ALL || out.done) out.done += n_envs;
    if constexpr (C + 1 < NC) lean_rollout_chain<NC, C + 1, ALL>(S, cfg, rank_stride, kind, f, w, steps, n_steps, rsum, out, n_envs);
}

template <int NC, bool ALL>
__global__ void __launch_bounds__(kRolloutThreads, 4)
rollout_lean_kernel(DevCfg cfg, uint4 *__restrict__ recs, uint32_t n_envs, uint32_t env_id_base, int rounds,
                    uint64_t first_round, uint32_t seed, int kind, RolloutOut out, const FastTabs *__restrict__ tabs,
                    uint32_t rank_stride, Control *ctl) {
    extern __shared__ uint4 smem_dyn[];
    LeanTabs<NC> &S = *reinterpret_cast<LeanTabs<NC> *>(smem_dyn);
    __shared__ long long cta_reward;
    __shared__ CtaCounters cc;
    stage_lean_tabs<NC, kRolloutThreads>(S, tabs);
    if (threadIdx.x == 0) { cc = CtaCounters{0u, 0u, 0u, 0u, 0, 0u}; cta_reward = 0ll; }
    __syncthreads();
    const uint32_t tiles = (n_envs + 31u) >> 5;
    const uint32_t t0 = (uint32_t)((uint64_t)tiles * blockIdx.x / gridDim.x);
    const uint32_t t1 = (uint32_t)((uint64_t)tiles * (blockIdx.x + 1u) / gridDim.x);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    StepCounts cnt;
    long long rsum = 0;
    uint4 rec_next = make_uint4(0u, 0u, 0u, 0u); // the next tile's records are in flight during this tile's work
    if (t0 + warp < t1 && (t0 + warp) * 32u + lane < n_envs) rec_next = recs[(t0 + warp) * 32u + lane];
    for (uint32_t t = t0 + warp; t < t1; t += n_warps) {
        const uint32_t e = t * 32u + lane;
        uint4 rec = rec_next;
        if (t + n_warps < t1 && (t + n_warps) * 32u + lane < n_envs) rec_next = recs[(t + n_warps) * 32u + lane];
        if (e >= n_envs) continue;
        uint32_t f[NC];
        {
            const uint64_t cars = (uint64_t)rec.x | ((uint64_t)rec.y << 32);
#pragma unroll
            for (int i = 0; i < NC; i++) f[i] = (uint32_t)(cars >> (12 * i)) & 0xFFFu;
        }
        uint32_t episode = rec.z & 0xFFFFu, steps = rec.z >> 16;
        const uint32_t env = env_id_base + e;
        int32_t rtile = 0;
        for (int r = 0; r < rounds; r++) {
            const uint64_t round = first_round + (uint64_t)r;
            const uint4 p = philox4x32_r((uint32_t)round, (uint32_t)(round >> 32), 0, 0, seed, env);
            uint32_t w[NC];
            w[0] = p.x;
            if (NC > 1) w[NC > 1 ? 1 : 0] = p.y;
            if (NC > 2) w[NC > 2 ? 2 : 0] = p.z;
            if (NC > 3) w[NC > 3 ? 3 : 0] = p.w;
            if (NC > 4) w[NC > 4 ? 4 : 0] = philox4x32_r((uint32_t)round, (uint32_t)(round >> 32), 1, 0, seed, env).x;
            int32_t rr = 0;
            RolloutOut o = out; // this env's slot of car 0 in round r
            const size_t off = (size_t)r * NC * (size_t)n_envs + e;
            if (ALL || o.action) o.action += off;
            if (ALL || o.reward) o.reward += off;
            if (ALL || o.next) o.next += off;
            if (ALL || o.done) o.done += off;
            lean_rollout_chain<NC, 0, ALL>(S, cfg, rank_stride, kind, f, w, steps, cnt.steps, rr, o, (size_t)n_envs);
            rtile += rr;
            if ((r & 63) == 63) { rsum += rtile; rtile = 0; } // |reward| <= 1e5 per step: 64 rounds x 5 cars fit 32 bits
            steps = min(steps, 0xFFFFu);
            // end of round: is_game_ended (src/env_gym.py:622-638) or the step cap -> next episode
            uint32_t d = f[0];
#pragma unroll
            for (int i = 1; i < NC; i++) d &= f[i];
            const bool finished = (d >> 8) & 1u;
            const bool capped = !finished && cfg.step_cap > 0 && steps >= (uint32_t)cfg.step_cap;
            if (finished || capped) {
                cnt.episodes++;
                cnt.capped += capped;
                episode = (episode + 1) & 0xFFFFu;
                steps = 0;
                const uint32_t hb = heading_word(round, 2, seed, env);
#pragma unroll
                for (int i = 0; i < NC; i++)
                    f[i] = make_field(cfg.start_box18[i] * 4 + ((hb >> (2 * i)) & 3u), cfg.start_di[i], 0, 5);
            }
        }
        rsum += rtile;
        uint64_t cars = 0;
#pragma unroll
        for (int i = 0; i < NC; i++) cars |= (uint64_t)f[i] << (12 * i);
        rec.x = (uint32_t)cars;
        rec.y = (uint32_t)(cars >> 32);
        rec.z = episode | (steps << 16);
        recs[e] = rec;
    }
    cnt.explore = cnt.steps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xFFFFFFFFu, rsum, o);
    if (lane == 0 && rsum) atomicAdd((unsigned long long *)&cta_reward, (unsigned long long)rsum);
    flush_counts<true>(cnt, cc, MDPP_LEARN_FROZEN, ctl); // no Q-updates on this path
    if (threadIdx.x == 0 && cta_reward)
        atomicAdd((unsigned long long *)&ctl->stats[blockIdx.x & (kStatSlots - 1)].reward_sum, (unsigned long long)cta_reward);
}

// ------------------------------------------------------------------------------------------------
// Pure-exploration rounds of LARGE tables (scan mode, 3-5 cars): all cars of up to `rounds` rounds in one pass over
// the batch.  While every env explores nothing an env does depends on the table, so the per-car sub-step kernels
// (each a full pass over the records: 4 cars = 4 x 15 us at 2^20 envs) collapse into the lean step of the rollout
// kernel -- cars unpacked in registers for the whole launch, tables in shared memory -- plus one fire-and-forget
// atomic OR per agent-step into the touched bitmap of its sub-step (region r * NC + c).  The pulls of those
// sub-steps follow in order (q_pull_scan_kernel).  The host launches this kernel only while its bound on the
// episode counters keeps every env inside the epsilon = 1 segment for all `rounds` rounds.
// ------------------------------------------------------------------------------------------------
struct PureScanArgs {
    uint4 *recs;
    uint32_t n_envs, env_id_base;
    uint32_t rounds;
    uint64_t first_round;
    const FastTabs *tabs;
    uint32_t rank_stride, n_states;
    uint32_t *bits0;       // bitmap of sub-step (0, 0); sub-step (r, c) = bits0 + (r * NC + c) * region_words
    uint32_t region_words;
    uint32_t scan_magic;   // floor(2^32 / region_words) + 1 (scan_slot)
    uint32_t l1_check;
};

// The mark of an agent-step is split in two.  Its bitmap word is REQUESTED (through L1, scan_mark's test) as soon as
// the key is known; it is tested, and the atomic issued, one car-step later -- ncu on the first form, which tested
// the word where it was loaded, had a third of the kernel's stall samples on those four instructions
// (profiles/sass_pure_scan4_r02.txt).  `pend_key` / `pend_val`: where the key marked by the previous car-step of this
// env lives (scan_slot; kNoKey: none) and that word; its bitmap is the region before this car-step's.
constexpr uint32_t kNoKey = 0xFFFFFFFFu;
__device__ __forceinline__ void scan_mark_resolve(uint32_t *__restrict__ bits, uint32_t slot /* scan_slot */, uint32_t val) {
    const uint32_t bit = 1u << (slot >> 27);
    if (slot != kNoKey && !(val & bit)) atomicOr(bits + (slot & 0x7FFFFFFu), bit);
}
template <int NC, int C>
__device__ __forceinline__ void lean_mark_chain(const LeanTabs<NC> &S, const PureScanArgs &pa, uint32_t (&f)[NC],
                                                const uint32_t (&w)[NC], uint32_t &steps, uint32_t &n_steps,
                                                uint32_t *__restrict__ bits, uint32_t &pend_key, uint32_t &pend_val,
                                                Control *ctl) {
    const uint32_t fc = f[C];
    uint32_t key = kNoKey, val = 0;
    if (!(fc & 0x100u)) { // `if env.check_car_training(num): continue` (qlearning.py:216-217)
        const LeanView v = lean_view<NC, C>(S, f, pa.rank_stride); // the encode's ghost side effects stay in f[]
        uint32_t lm = v.lm;
        if (v.sidx >= pa.n_states) { // guard: the encoded state is off the table -- the car does not act
            guard_tripped(ctl);
            lm = 0;
        }
        if (lm) {
            // random.choice(legal) (qlearning.py:102): low 16 bits of the decision word scaled to the legal count
            const uint32_t a = (S.nth[lm] >> (3u * (((w[C] & 0xFFFFu) * (uint32_t)__popc(lm)) >> 16))) & 7u;
            key = scan_slot(v.sidx * MDPP_Q_ROW + a, pa.region_words, pa.scan_magic);
            if (pa.l1_check) val = bits[key & 0x7FFFFFFu]; // requested now, tested a car-step later
            const uint32_t mv = S.next[C][fc & 0xFFu][a]; // successor + update_car
            f[C] = (mv & 0x1FFu) | (fc & 0xE00u);
            steps += 1;
            n_steps += 1;
        }
    }
    scan_mark_resolve(bits - pa.region_words, pend_key, pend_val); // the previous car-step's mark
    pend_key = key;
    pend_val = val;
    if constexpr (C + 1 < NC)
        lean_mark_chain<NC, C + 1>(S, pa, f, w, steps, n_steps, bits + pa.region_words, pend_key, pend_val, ctl);
}

template <int NC>
__global__ void __launch_bounds__(kRolloutThreads, 4)
q_rounds_pure_scan_kernel(DevCfg cfg, LearnArgs lp, PureScanArgs pa, Control *ctl) {
    extern __shared__ uint4 smem_dyn[];
    LeanTabs<NC> &S = *reinterpret_cast<LeanTabs<NC> *>(smem_dyn);
    __shared__ CtaCounters cc;
    pdl_launch_dependents(); // the pulls that follow wait for this grid to complete
    stage_lean_tabs<NC, kRolloutThreads>(S, pa.tabs);
    if (threadIdx.x == 0) cc = CtaCounters{0u, 0u, 0u, 0u, 0, 0u};
    __syncthreads();
    const uint32_t tiles = (pa.n_envs + 31u) >> 5;
    const uint32_t t0 = (uint32_t)((uint64_t)tiles * blockIdx.x / gridDim.x);
    const uint32_t t1 = (uint32_t)((uint64_t)tiles * (blockIdx.x + 1u) / gridDim.x);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    StepCounts cnt;
    uint4 rec_next = make_uint4(0u, 0u, 0u, 0u); // the next tile's records are in flight during this tile's work
    if (t0 + warp < t1 && (t0 + warp) * 32u + lane < pa.n_envs) rec_next = pa.recs[(t0 + warp) * 32u + lane];
    for (uint32_t t = t0 + warp; t < t1; t += n_warps) {
        const uint32_t e = t * 32u + lane;
        uint4 rec = rec_next;
        if (t + n_warps < t1 && (t + n_warps) * 32u + lane < pa.n_envs) rec_next = pa.recs[(t + n_warps) * 32u + lane];
        uint32_t episode = rec.z & 0xFFFFu, steps = rec.z >> 16;
        if (e >= pa.n_envs || (int32_t)episode >= lp.episodes) continue; // an env that used its episode budget is frozen
        uint32_t f[NC];
        {
            const uint64_t cars = (uint64_t)rec.x | ((uint64_t)rec.y << 32);
#pragma unroll
            for (int i = 0; i < NC; i++) f[i] = (uint32_t)(cars >> (12 * i)) & 0xFFFu;
        }
        const uint32_t env = pa.env_id_base + e;
        uint32_t *bits = pa.bits0;
        uint32_t pend_key = kNoKey, pend_val = 0;
        for (uint32_t r = 0; r < pa.rounds; r++) {
            const uint64_t round = pa.first_round + (uint64_t)r;
            const uint4 p = philox4x32_r((uint32_t)round, (uint32_t)(round >> 32), 0, 0, lp.seed, env);
            uint32_t w[NC];
            w[0] = p.x;
            if (NC > 1) w[NC > 1 ? 1 : 0] = p.y;
            if (NC > 2) w[NC > 2 ? 2 : 0] = p.z;
            if (NC > 3) w[NC > 3 ? 3 : 0] = p.w;
            if (NC > 4) w[NC > 4 ? 4 : 0] = philox4x32_r((uint32_t)round, (uint32_t)(round >> 32), 1, 0, lp.seed, env).x;
            lean_mark_chain<NC, 0>(S, pa, f, w, steps, cnt.steps, bits, pend_key, pend_val, ctl);
            bits += (size_t)NC * pa.region_words;
            steps = min(steps, 0xFFFFu);
            // end of the round-robin round (qlearning.py:233): is_game_ended or the step cap -> next episode
            uint32_t d = f[0];
#pragma unroll
            for (int i = 1; i < NC; i++) d &= f[i];
            const bool finished = (d >> 8) & 1u;
            const bool capped = !finished && cfg.step_cap > 0 && steps >= (uint32_t)cfg.step_cap;
            if (finished || capped) {
                cnt.episodes++;
                cnt.capped += capped;
                episode = episode + 1;
                cnt.max_episode = max(cnt.max_episode, episode);
                steps = 0;
                rec.w &= ~0xFFu; // deadlock_count = 0 per episode (qlearning.py:331)
                const uint32_t hb = heading_word(round, 2, lp.seed, env);
#pragma unroll
                for (int i = 0; i < NC; i++)
                    f[i] = make_field(cfg.start_box18[i] * 4 + ((hb >> (2 * i)) & 3u), cfg.start_di[i], 0, 5);
                if ((int32_t)episode >= lp.episodes) break; // budget used up: frozen from here on
            }
        }
        scan_mark_resolve(bits - pa.region_words, pend_key, pend_val); // the last car-step's mark
        uint64_t cars = 0;
#pragma unroll
        for (int i = 0; i < NC; i++) cars |= (uint64_t)f[i] << (12 * i);
        rec.x = (uint32_t)cars;
        rec.y = (uint32_t)(cars >> 32);
        rec.z = (episode & 0xFFFFu) | (steps << 16);
        pa.recs[e] = rec;
    }
    cnt.explore = cnt.steps; // every decision of this kernel explores
    flush_counts<true>(cnt, cc, 0u, ctl);
    pdl_wait();
}

} // namespace mdpp
