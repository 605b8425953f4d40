// fast_tabs.cuh -- table-driven env step shared by the learner, round and rollout kernels (sm_100a).
//
// ncu showed the env arithmetic bound by the integer ALU pipe (half issue rate on sm_100).  Everything the
// reference derives from a car's (node, destination_index) is a per-(car, node, destination_index) constant, so
// it is read from shared-memory tables instead of being recomputed: one 16-byte entry for the primary
// (state-index base, destination node, F successor, base legal mask, occupancy test masks, selective-block
// mask, BFS length), one 8-byte entry per other car (its listed (box, destination box) code and occupancy bits),
// and the car field after each of the 5 actions (successor + update_car folded, with the goal flag and the
// successor's BFS length for the reward).  The image is built on the host (mdpp_api.inl build_fast_tabs,
// exported as mdpp_fast_tables and replayed against the oracle env by tests/test_host_tables.py).
#pragma once
#include "env_core.cuh"

namespace mdpp {

// host-built table image (workspace); the kernels copy the first n_cars rows to shared memory
struct FastTabs {
    uint4 own[MDPP_MAX_CARS][256];  // index = node | destination_index << 7 (the low 8 bits of a car field)
    uint2 oth[MDPP_MAX_CARS][256];  // same index, car seen as an "other"
    uint16_t next[MDPP_MAX_CARS][256][8]; // [car][node | di << 7][action]: successor (src/env_gym.py:256-265) +
        // update_car (:386-441) folded: bits 0-8 = node | di << 7 | done << 8 after the move, bit 9 = the move reached
        // the car's destination node, bits 10-15 = BFS length successor -> that destination (q_reward / dqn_reward)
    uint32_t nth[32];               // per 5-bit mask: positions of its set bits, 3 bits each, lowest first
    uint2 ghost[MDPP_MAX_CARS];     // oth-format entry of the finished car's ghost
};
// own.x: bits 0-9 dnidx * n_local_nodes + node_local | 10-16 destination node | 17-23 F successor (127 none)
//        | 24-28 base legal mask | 29-31 destination-box index
// own.y: bits 0-8 1 << (own box number - 1) | 9-17 1 << (number of the box ahead - 1) | 18 box ahead is in the
//        selective-block list keyed by this car's current destination box
// own.z: that selective-block list as an 18-bit box mask        own.w: bits 0-7 BFS length node -> the car's destination
// oth.x: bits 0-17 1 << box18 | 18-25 others-code        oth.y: bits 0-8 1 << (box number - 1)

// the first NC rows of the image, staged in shared memory by the kernels that step many tiles per CTA;
// StepTabs<MDPP_MAX_CARS> is a prefix of FastTabs, so a kernel may also read the image in place (through L1)
template <int NC>
struct StepTabs {
    uint4 own[NC][256];
    uint2 oth[NC][256];
    uint16_t next[NC][256][8];
    uint32_t nth[32];
};
static_assert(sizeof(StepTabs<MDPP_MAX_CARS>) <= sizeof(FastTabs), "StepTabs<5> must be a prefix of FastTabs");

template <int NC, int BLOCK>
__device__ __forceinline__ void stage_step_tabs(StepTabs<NC> &S, const FastTabs *__restrict__ g) {
    const uint4 *src_own = &g->own[0][0];
    for (uint32_t i = threadIdx.x; i < NC * 256u; i += BLOCK) (&S.own[0][0])[i] = __ldg(src_own + i);
    const uint2 *src_oth = &g->oth[0][0];
    for (uint32_t i = threadIdx.x; i < NC * 256u; i += BLOCK) (&S.oth[0][0])[i] = __ldg(src_oth + i);
    const uint4 *src_next = reinterpret_cast<const uint4 *>(&g->next[0][0][0]); // 256 x 8 x u16 = 256 uint4 per car
    for (uint32_t i = threadIdx.x; i < NC * 256u; i += BLOCK)
        reinterpret_cast<uint4 *>(&S.next[0][0][0])[i] = __ldg(src_next + i);
    if (threadIdx.x < 32) S.nth[threadIdx.x] = __ldg(&g->nth[threadIdx.x]);
}

// What the primary car C sees (encodestate, src/env_gym.py:281-329) and may do (getlegalactions_filtered,
// :143-224), from the tables.  mode 1 = training: a finished car's ghost is listed while its counter lasts and
// the counter moves (the side effect lands in `cars`, committed by the caller); mode 0 = detect_deadlocks_v0:
// finished cars are not listed; mode 2 = the state was already encoded for this agent-step: ghosts listed, no
// second decrement.
struct FastView {
    uint4 own;       // the primary's table entry
    uint32_t f;      // its 12-bit field
    uint32_t rank, sidx, lm;
    uint32_t occ9, occ18; // boxes the listed others occupy: by box number, by floor + box
    uint64_t cars;   // records after the encode's side effects
};

// getlegalactions_filtered (src/env_gym.py:143-224) for a primary whose table entry is `own`
__device__ __forceinline__ uint32_t fast_legal(const uint4 &own, uint32_t occ9, uint32_t occ18) {
    uint32_t lm = (own.x >> 24) & 31u;
    const uint32_t hit = (occ9 * 0x201u) & own.y; // occupancy against the own box (bits 0-8) and the box ahead (9-17)
    if (hit & 0x1FFu) lm &= 1u << MDPP_A_F;
    if ((hit >> 9) || (((own.y >> 18) & 1u) && (own.z & occ18))) lm &= ~(1u << MDPP_A_F);
    return lm;
}

template <int NC, int C, class Tabs>
__device__ __forceinline__ FastView fast_view(const Tabs &S, const FastTabs *__restrict__ gt, uint64_t cars,
                                              int mode, uint32_t rank_stride) {
    FastView v;
    v.f = (uint32_t)(cars >> (12 * C)) & 0xFFFu;
    v.own = S.own[C][v.f & 0xFFu];
    v.cars = cars;
    uint32_t occ9 = 0, occ18 = 0, codes[NC];
#pragma unroll
    for (int i = 0; i < NC; i++) {
        codes[i] = 0;
        if (i == C) continue;
        const uint32_t fi = (uint32_t)(cars >> (12 * i)) & 0xFFFu;
        uint2 oe = make_uint2(0u, 0u);
        if (!(fi & 0x100u)) {
            oe = S.oth[i][fi & 0xFFu];
        } else if (mode == 1) {
            uint32_t g = fi >> 9;
            if (g != 7u) {
                if (g == 0u) g = 7u;                 // count went negative: the ghost vanishes for good
                else { g -= 1u; oe = gt->ghost[i]; } // still shown (5 encodes in total)
                v.cars = (v.cars & ~(0xE00ull << (12 * i))) | ((uint64_t)(g << 9) << (12 * i));
            }
        } else if (mode == 2 && (fi >> 9) != 7u) {
            oe = gt->ghost[i];
        }
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
    v.rank = rank;
    v.sidx = rank * rank_stride + (v.own.x & 0x3FFu);
    const uint32_t lm = fast_legal(v.own, occ9, occ18);
    v.occ9 = occ9;
    v.occ18 = occ18;
    v.lm = lm;
    return v;
}

// q_reward / dqn_reward (src/env_gym.py:485-620) of move entry `mv` as an integer (dqn: before the /100).
// `cars` = the records BEFORE update_car, as in the reference where the reward is computed first.
template <int NC, int C>
__device__ __forceinline__ int32_t fast_reward(const FastView &v, uint32_t mv, uint64_t cars, int kind, const DevCfg &cfg) {
    const bool reached = (mv >> 9) & 1u;
    const uint32_t pnode = v.f & 127u, dnode = (v.own.x >> 10) & 127u;
    const uint32_t nn = reached ? dnode : (mv & 127u); // the successor itself (a finished car is parked afterwards)
    if (kind == MDPP_REWARD_Q && cfg.bad_on) {           // :568-577 on the heading-stripped strings
        const uint32_t base = (v.rank * cfg.n_dest_boxes + (v.own.x >> 29)) * cfg.n_src_boxes;
        const uint32_t idn = base + ((nn >> 2) - cfg.box_base), idp = base + ((pnode >> 2) - cfg.box_base);
        if ((__ldg(cfg.bad_bitmap + (idn >> 5)) >> (idn & 31u)) & 1u) return -10000;
        if ((__ldg(cfg.bad_bitmap + (idp >> 5)) >> (idp & 31u)) & 1u) return 0;
    }
    int32_t r = -10 + 50 * ((int32_t)(v.own.w & 0xFFu) - (int32_t)(mv >> 10)); // :603-605
    if (reached) r += (kind == MDPP_REWARD_DQN) ? 110 : 10010;                  // :606-607 / :536-537
    int32_t clash = 0;
#pragma unroll
    for (int i = 0; i < NC; i++)                                               // :613-619 real positions, floor AND box
        if (i != C) clash += ((((uint32_t)(cars >> (12 * i)) & 127u) >> 2) == (nn >> 2));
    return r - clash * ((kind == MDPP_REWARD_DQN) ? 100 : 100000);
}

// ------------------------------------------------------------------------------------------------
// "Lean" form of the same step for the kernels that keep a tile's cars in registers across a whole round
// (rollout_fast_kernel; ncu on the previous form: 1101 warp-instructions per tile and round for 4 cars, a quarter
// of them 64-bit shifts / masks on the packed car word and branches around finished cars):
//   * the cars live UNPACKED, one 12-bit field per 32-bit register, for all car-steps of a round;
//   * an "other" is looked up with 9 index bits (node | destination_index << 7 | done << 8): the upper half of the
//     table is zero, so a finished car is simply not listed -- no branch;
//   * the ghost of a finished car (listed for 5 encodes by the others, src/env_gym.py:307-318) is branch-free: its
//     counter steps through a nibble table, its entry is OR-ed in under a mask;
//   * only the others are sorted, and the combinatorial rank comes from three small tables C(c+1,2), C(c+2,3),
//     C(c+3,4) instead of multiplications and constant divisions.
// Mode 1 (training / rollout: ghosts listed, counters move) only.  Bit-identical to fast_view by the oracle tests.
// ------------------------------------------------------------------------------------------------
#ifndef MDPP_LEAN_GHOST_BRANCH
#define MDPP_LEAN_GHOST_BRANCH 1 // rollout (lean_view): 1 = ghost behind a branch (4 cars, 2^24 envs: 9.41e10 against 8.80e10 steps/s), 0 = branch-free
#endif
constexpr int kLeanCodes = 160; // others-codes are <= 1 + 18 * 8
template <int NC>
struct LeanTabs {
    uint4 own[NC][256];
    uint2 oth[NC][512];           // [256, 512): done cars -> {0, 0}
    uint16_t next[NC][256][8];
    uint32_t nth[32];
    uint2 ghost[NC];
    uint32_t tri[kLeanCodes], tet[kLeanCodes], quad[kLeanCodes];
};

template <int NC, int BLOCK>
__device__ __forceinline__ void stage_lean_tabs(LeanTabs<NC> &S, const FastTabs *__restrict__ g) {
    const uint4 *src_own = &g->own[0][0];
    for (uint32_t i = threadIdx.x; i < NC * 256u; i += BLOCK) (&S.own[0][0])[i] = __ldg(src_own + i);
    for (uint32_t i = threadIdx.x; i < NC * 512u; i += BLOCK) {
        const uint32_t car = i >> 9, e = i & 511u;
        S.oth[car][e] = e < 256u ? __ldg(&g->oth[car][e]) : make_uint2(0u, 0u);
    }
    const uint4 *src_next = reinterpret_cast<const uint4 *>(&g->next[0][0][0]);
    for (uint32_t i = threadIdx.x; i < NC * 256u; i += BLOCK)
        reinterpret_cast<uint4 *>(&S.next[0][0][0])[i] = __ldg(src_next + i);
    if (threadIdx.x < 32) S.nth[threadIdx.x] = __ldg(&g->nth[threadIdx.x]);
    if (threadIdx.x < NC) S.ghost[threadIdx.x] = __ldg(&g->ghost[threadIdx.x]);
    for (uint32_t c = threadIdx.x; c < (uint32_t)kLeanCodes; c += BLOCK) {
        const unsigned long long x = c;
        S.tri[c] = (uint32_t)((x + 1) * x / 2);
        S.tet[c] = (uint32_t)((x + 2) * (x + 1) * x / 6);
        S.quad[c] = (uint32_t)((x + 3) * (x + 2) * (x + 1) * x / 24);
    }
}

struct LeanView {
    uint4 own;
    uint32_t rank, sidx, lm;
};

// encodestate(car C, 1) + getlegalactions_filtered on the unpacked fields f[]; the ghost side effects land in f[].
template <int NC, int C>
__device__ __forceinline__ LeanView lean_view(const LeanTabs<NC> &S, uint32_t (&f)[NC], uint32_t rank_stride) {
    LeanView v;
    v.own = S.own[C][f[C] & 0xFFu];
    uint32_t occ9 = 0, occ18 = 0, codes[NC > 1 ? NC - 1 : 1];
    codes[0] = 0;
    int k = 0;
#pragma unroll
    for (int i = 0; i < NC; i++) {
        if (i == C) continue;
        const uint32_t fi = f[i], g = fi >> 9;
        uint2 oe = S.oth[i][fi & 0x1FFu]; // zero for a finished car
        // ghost of a finished car: counter 5..1 -> shown, then one lower; 0 -> 7 (gone for good); 7 stays
#if MDPP_LEAN_GHOST_BRANCH
        if ((fi & 0x100u) && g != 7u) {
            if (g != 0u) oe = S.ghost[i];
            f[i] = (fi & 0x1FFu) | ((g == 0u ? 7u : g - 1u) << 9);
        }
#else
        const uint32_t done = (fi >> 8) & 1u;
        const uint32_t show = done & (0x7Eu >> g);                  // g in 1..6
        const uint32_t ng = done ? ((0x75432107u >> (4u * g)) & 7u) : g;
        const uint32_t mask = 0u - show;
        oe.x |= S.ghost[i].x & mask;
        oe.y |= S.ghost[i].y & mask;
        f[i] = (fi & 0x1FFu) | (ng << 9);
#endif
        codes[k++] = oe.x >> 18;
        occ18 |= oe.x & 0x3FFFFu;
        occ9 |= oe.y;
    }
    uint32_t rank = 0;
    if constexpr (NC == 2) rank = codes[0];
    if constexpr (NC > 2) {
        sort_codes<NC - 1>(codes);
        rank = codes[0] + S.tri[codes[1]];
        if constexpr (NC > 3) rank += S.tet[codes[2]];
        if constexpr (NC > 4) rank += S.quad[codes[3]];
    }
    v.rank = rank;
    v.sidx = rank * rank_stride + (v.own.x & 0x3FFu);
    v.lm = fast_legal(v.own, occ9, occ18);
    return v;
}

} // namespace mdpp
