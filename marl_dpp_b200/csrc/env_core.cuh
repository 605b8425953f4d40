// env_core.cuh -- device-side env semantics shared by every kernel (sm_100a).
//
// One env = one 16-byte record (uint4), one lane owns one env.  Cars are bit-packed, 12 bits each,
// in the first 64 bits:  node[6:0] | destination_index[7] | training_done[8] | ghost[11:9]
// (ghost = Car.training_done_count clamped: 5..0, 7 = negative).  Word 2 = episode[15:0] |
// steps_in_episode[31:16]; word 3 = reserved.  Everything the reference keeps in strings/dicts is
// integer arithmetic on these fields plus two small shared-memory tables.
//
// Reference semantics restated here (paths relative to the reference root):
//   encodestate            src/env_gym.py:281-329      -> view_of()
//   getlegalactions        src/env_gym.py:143-224      -> legal_mask()
//   get_successor_state    src/env_gym.py:256-265      -> next_node()
//   q_reward / dqn_reward  src/env_gym.py:485-620      -> reward_int()
//   update_car             src/env_gym.py:386-441      -> advance_car()
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mdpp.h"

namespace mdpp {

constexpr int kThreads = 256;
constexpr uint32_t kNone = MDPP_NONE;

// Kernel-parameter block: scalars and per-car constants (uniform accesses -> constant bank).
struct DevCfg {
    int32_t n_cars, n_local_nodes, n_dest_nodes, n_dest_boxes, n_src_boxes;
    int32_t node_base, box_base;  // subtracted from node / box18 to get the local index
    int32_t parking_node, tcdi, bad_on, step_cap;
    uint32_t car_dnidx[MDPP_MAX_CARS][2];   // destination-node index per destination_index
    uint32_t car_dnode[MDPP_MAX_CARS][2];   // destination node id
    uint32_t car_dbox[MDPP_MAX_CARS][2];    // destination-box index
    uint32_t car_dbox18[MDPP_MAX_CARS][2];  // destination box18
    uint32_t car_sel[MDPP_MAX_CARS][2];     // selective-block mask keyed by that destination box
    uint32_t car_ghost_code[MDPP_MAX_CARS]; // others-code of the finished car's ghost
    uint32_t car_ghost_box18[MDPP_MAX_CARS];
    uint32_t start_box18[MDPP_MAX_CARS], start_di[MDPP_MAX_CARS];
    uint32_t dn_node[MDPP_MAX_DEST];  // destination-node index -> node id           (query kernel only:
    uint32_t dn_dbox[MDPP_MAX_DEST];  // destination-node index -> destination-box   per-lane divergent
    uint32_t dbox_sel[MDPP_MAX_DEST]; // destination-box index -> selective mask     constant reads)
    const uint32_t *tables;      // DEV: SmemTables image
    const uint32_t *bad_bitmap;  // DEV: n_box_states bits
};

// Divergently indexed tables live in shared memory (a constant-bank lookup would serialise).
struct SmemTables {
    uint16_t nodeinfo[MDPP_NODES];             // fwd successor | base_legal << 8
    uint8_t bfs[MDPP_MAX_DEST][MDPP_NODES];    // path length node -> destination node
};
constexpr int kTableWords = (sizeof(SmemTables) + 3) / 4;

// Table access of the parity / gym / string-API kernels: a shared-memory copy per CTA.
struct SharedTab {
    const SmemTables *sm;
    __device__ __forceinline__ uint32_t info(uint32_t node) const { return sm->nodeinfo[node]; }
    __device__ __forceinline__ uint32_t bfs(uint32_t d, uint32_t node) const { return sm->bfs[d][node]; }
};
__device__ __forceinline__ void load_tables(SmemTables &sm, const uint32_t *src) {
    uint32_t *dst = reinterpret_cast<uint32_t *>(&sm);
    for (int i = threadIdx.x; i < kTableWords; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}

// ---- Philox4x32 (Salmon et al. SC'11), counter-based: no per-env generator state -------------
// MDPP_PHILOX_ROUNDS rounds: 7 by default -- the fewest with which Philox4x32 passes BigCrush in the authors' tests
// (their default of 10 is a safety margin) -- because the generator is a fifth of the round kernels' instructions
// (65 of 290 per 32-env tile and round at 10 rounds, profiles/ncu_pure_rounds_r02.txt).  The oracle is built with the
// same constant (oracle/mdpp_oracle.c); both check the 7- and the 10-round Random123 known-answer vectors.
#ifndef MDPP_PHILOX_ROUNDS
#define MDPP_PHILOX_ROUNDS 7
#endif
template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0;
        c1 = l1;
        c2 = h0 ^ c3 ^ k1;
        c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint4 philox4x32_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    return philox4x32<MDPP_PHILOX_ROUNDS>(c0, c1, c2, c3, k0, k1);
}
// RNG streams (ctr word 2): 0 = decision words of cars 0..3, 1 = car 4, 2 = headings drawn at the
// end-of-round reset, 3 = headings of mdpp_reset.  ctr = (round_lo, round_hi, stream, 0), key = (seed, env id).
__device__ __forceinline__ uint32_t decision_word(uint64_t round, int car, uint32_t seed, uint32_t env) {
    uint4 p = philox4x32_r((uint32_t)round, (uint32_t)(round >> 32), car >> 2, 0, seed, env);
    int w = car & 3;
    return w == 0 ? p.x : w == 1 ? p.y : w == 2 ? p.z : p.w;
}
__device__ __forceinline__ uint32_t heading_word(uint64_t round, uint32_t stream, uint32_t seed, uint32_t env) {
    return philox4x32_r((uint32_t)round, (uint32_t)(round >> 32), stream, 0, seed, env).x;
}

// ---- packed car fields -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t car_field(uint64_t cars, int i) { return (uint32_t)(cars >> (12 * i)) & 0xFFFu; }
__device__ __forceinline__ uint64_t set_car_field(uint64_t cars, int i, uint32_t f) {
    return (cars & ~(0xFFFull << (12 * i))) | ((uint64_t)f << (12 * i));
}
__device__ __forceinline__ uint32_t f_node(uint32_t f) { return f & 127u; }
__device__ __forceinline__ uint32_t f_di(uint32_t f) { return (f >> 7) & 1u; }
__device__ __forceinline__ uint32_t f_done(uint32_t f) { return (f >> 8) & 1u; }
__device__ __forceinline__ uint32_t f_ghost(uint32_t f) { return (f >> 9) & 7u; }
__device__ __forceinline__ uint32_t make_field(uint32_t node, uint32_t di, uint32_t done, uint32_t ghost) {
    return node | (di << 7) | (done << 8) | (ghost << 9);
}
__device__ __forceinline__ uint32_t boxnum0(uint32_t box18) { return box18 >= 9 ? box18 - 9 : box18; } // box number - 1

template <int NC>
__device__ __forceinline__ bool all_done(uint64_t cars) {
    bool d = true;
#pragma unroll
    for (int i = 0; i < NC; i++) d = d && f_done(car_field(cars, i));
    return d;
}

template <int NC>
__device__ __forceinline__ uint64_t initial_cars(const DevCfg &cfg, uint32_t heading_bits) {
    uint64_t cars = 0;
#pragma unroll
    for (int i = 0; i < NC; i++) {
        uint32_t node = cfg.start_box18[i] * 4 + ((heading_bits >> (2 * i)) & 3u);
        cars |= (uint64_t)make_field(node, cfg.start_di[i], 0, 5) << (12 * i);
    }
    return cars;
}

// What the primary car sees: the reference's state string as integers.
struct View {
    uint32_t pnode, di, dnidx, dnode, dbox, sel;  // dbox/sel: destination-box index, selective-block mask
    uint32_t rank;          // rank of the sorted multiset of the other cars' (src box, dst box) codes
    uint32_t occ9, occ18;   // boxes the listed others occupy: by box number, by floor+box
};

template <int N>
__device__ __forceinline__ void sort_codes(uint32_t (&v)[N]) {
#define MDPP_CX(a, b) { uint32_t lo = min(v[a], v[b]), hi = max(v[a], v[b]); v[a] = lo; v[b] = hi; }
    if constexpr (N == 2) { MDPP_CX(0, 1) }
    if constexpr (N == 3) { MDPP_CX(0, 1) MDPP_CX(1, 2) MDPP_CX(0, 1) }
    if constexpr (N == 4) { MDPP_CX(0, 1) MDPP_CX(2, 3) MDPP_CX(0, 2) MDPP_CX(1, 3) MDPP_CX(1, 2) }
    if constexpr (N == 5) {
        MDPP_CX(0, 1) MDPP_CX(3, 4) MDPP_CX(2, 4) MDPP_CX(2, 3) MDPP_CX(1, 4)
        MDPP_CX(0, 3) MDPP_CX(0, 2) MDPP_CX(1, 3) MDPP_CX(1, 2)
    }
#undef MDPP_CX
}

// encodestate (reference src/env_gym.py:281-329).  Slot `c` of the code array is the primary car
// itself (code 0 = "not listed"), so after sorting v[0] == 0 and v[1..NC-1] is the canonical
// multiset of the others; rank = sum_j C(v[j] + j - 1, j) (combinatorial number system).
// mode 1 mutates the ghost counters of finished cars exactly like the reference (:307-318).
template <int NC>
__device__ __forceinline__ View view_of(uint64_t &cars, int c, int mode, const DevCfg &cfg) {
    View v;
    uint32_t codes[NC];
    v.occ9 = 0;
    v.occ18 = 0;
    v.pnode = 0; v.di = 0;
#pragma unroll
    for (int i = 0; i < NC; i++) {
        uint32_t f = car_field(cars, i);
        if (i == c) {
            codes[i] = 0;
            v.pnode = f_node(f);
            v.di = f_di(f);
            continue;
        }
        uint32_t code = 0, box18 = 0;
        bool listed = false;
        if (f_done(f)) {
            if (mode == 1) {
                uint32_t g = f_ghost(f);
                if (g != 7u) {
                    if (g == 0u) {
                        g = 7u;               // count went negative: the ghost vanishes for good
                    } else {
                        g -= 1u;              // still shown (5 encodes in total)
                        listed = true;
                        code = cfg.car_ghost_code[i];
                        box18 = cfg.car_ghost_box18[i];
                    }
                    cars = set_car_field(cars, i, (f & 0x1FFu) | (g << 9));
                }
            } else if (mode == 2 && f_ghost(f) != 7u) {
                // the view a previous encodestate(., 1) of this agent-step produced: no second decrement
                listed = true;
                code = cfg.car_ghost_code[i];
                box18 = cfg.car_ghost_box18[i];
            }
        } else {
            listed = true;
            box18 = f_node(f) >> 2;
            code = 1u + (box18 - cfg.box_base) * cfg.n_dest_boxes + (f_di(f) ? cfg.car_dbox[i][1] : cfg.car_dbox[i][0]);
        }
        codes[i] = listed ? code : 0u;
        if (listed) {
            v.occ9 |= 1u << boxnum0(box18);
            v.occ18 |= 1u << box18;
        }
    }
    // two uniform constants and a select each (a register-indexed constant load is slow)
    v.dnidx = v.di ? cfg.car_dnidx[c][1] : cfg.car_dnidx[c][0];
    v.dnode = v.di ? cfg.car_dnode[c][1] : cfg.car_dnode[c][0];
    v.dbox = v.di ? cfg.car_dbox[c][1] : cfg.car_dbox[c][0];
    v.sel = v.di ? cfg.car_sel[c][1] : cfg.car_sel[c][0];
    uint32_t rank = 0;
    if constexpr (NC > 1) {
        sort_codes<NC>(codes);
        rank = codes[1];
        if constexpr (NC > 2) rank += (codes[2] + 1u) * codes[2] / 2u;
        if constexpr (NC > 3) rank += (codes[3] + 2u) * (codes[3] + 1u) * codes[3] / 6u;
        if constexpr (NC > 4) rank += (codes[4] + 3u) * (codes[4] + 2u) * (codes[4] + 1u) * codes[4] / 24u;
    }
    v.rank = rank;
    return v;
}

__device__ __forceinline__ uint32_t state_index(const View &v, uint32_t node, const DevCfg &cfg) {
    return (v.rank * cfg.n_dest_nodes + v.dnidx) * cfg.n_local_nodes + (node - cfg.node_base);
}

// getlegalactions_filtered (reference src/env_gym.py:188-224) for the primary car standing on
// `node`: 5-bit mask over S,L,R,F,T.  'B' never survives (:158-159).  S,L,R,T keep the car in its
// own box, so they die together when a listed car shares the box NUMBER (:211-217, floor ignored);
// F dies on the same test for the box ahead or on the selective-block rule (:143-186).
template <class Tab>
__device__ __forceinline__ uint32_t legal_mask(const View &v, uint32_t node, const Tab &sm) {
    const uint32_t sel = v.sel;
    uint32_t info = sm.info(node);
    uint32_t mask = info >> 8, fwd = info & 0xFFu;
    if ((v.occ9 >> boxnum0(node >> 2)) & 1u) mask &= 1u << MDPP_A_F;
    if (mask & (1u << MDPP_A_F)) {
        uint32_t fb = fwd >> 2;
        bool blocked = (v.occ9 >> boxnum0(fb)) & 1u;
        blocked = blocked || (((sel >> fb) & 1u) && (sel & v.occ18));
        if (blocked) mask &= ~(1u << MDPP_A_F);
    }
    return mask;
}

// get_successor_node (reference src/env_gym.py:132-141): L = heading-1, R = heading+1, T = heading-2
// in N,E,S,W order (src/configuration.py:147-156,251-254); F from the table.
template <class Tab>
__device__ __forceinline__ uint32_t next_node(uint32_t node, uint32_t a, const Tab &sm) {
    uint32_t rot = (0x20130u >> (4u * a)) & 3u;
    uint32_t turned = (node & ~3u) | ((node + rot) & 3u);
    return a == MDPP_A_F ? (sm.info(node) & 0xFFu) : turned;
}

// q_reward / dqn_reward as an integer (dqn: before the /100).  `cars` = the records BEFORE
// update_car, as in the reference where the reward is computed first (qlearning.py:223-229).
template <int NC, class Tab>
__device__ __forceinline__ int32_t reward_int(const View &v, uint32_t nnode, uint64_t cars, int c,
                                              int kind, const DevCfg &cfg, const Tab &sm) {
    if (kind == MDPP_REWARD_Q && cfg.bad_on) { // reference src/env_gym.py:568-577
        uint32_t base = (v.rank * cfg.n_dest_boxes + v.dbox) * cfg.n_src_boxes;
        uint32_t idn = base + ((nnode >> 2) - cfg.box_base), idp = base + ((v.pnode >> 2) - cfg.box_base);
        if ((__ldg(cfg.bad_bitmap + (idn >> 5)) >> (idn & 31u)) & 1u) return -10000;
        if ((__ldg(cfg.bad_bitmap + (idp >> 5)) >> (idp & 31u)) & 1u) return 0;
    }
    int32_t prev_steps = (int32_t)sm.bfs(v.dnidx, v.pnode), cur_steps = (int32_t)sm.bfs(v.dnidx, nnode);
    int32_t r = -10 + 50 * (prev_steps - cur_steps);                 // :603-605
    if (nnode == v.dnode) r += (kind == MDPP_REWARD_DQN) ? 110 : 10010; // :606-607 / :536-537
    int32_t clash = 0;
#pragma unroll
    for (int i = 0; i < NC; i++)                                     // :613-619 real positions, floor AND box
        if (i != c) clash += ((f_node(car_field(cars, i)) >> 2) == (nnode >> 2));
    r -= clash * ((kind == MDPP_REWARD_DQN) ? 100 : 100000);
    return r;
}

__device__ __forceinline__ float reward_float(int32_t r, int kind) {
    // reference :553 `reward/100` (float64, stored as float32 here).  For every integer |r| <= 2e6 the correctly
    // rounded SINGLE-precision quotient equals the float64 quotient rounded to float32 (checked exhaustively on the
    // host, tests/test_host_tables.py), so no FP64 instruction is needed
    return kind == MDPP_REWARD_DQN ? __fdiv_rn((float)r, 100.0f) : (float)r;
}

// update_car (reference src/env_gym.py:386-441): move; on reaching the current destination either
// finish (park at no_box_list[0]+'W', ghost counter armed at 5) or advance; destination_index toggles.
__device__ __forceinline__ uint64_t advance_car(uint64_t cars, int c, const View &v, uint32_t nnode, const DevCfg &cfg) {
    uint32_t f = car_field(cars, c);
    uint32_t di = v.di, done = 0, node = nnode;
    if (nnode == v.dnode) {
        if ((int)di == cfg.tcdi) {
            done = 1;
            node = cfg.parking_node;
        }
        di ^= 1u;
    }
    return set_car_field(cars, c, make_field(node, di, done, f_ghost(f)));
}

// Programmatic dependent launch (sm_90+): no-ops unless the kernel was launched with
// cudaLaunchAttributeProgrammaticStreamSerialization.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// One Q row = 8 floats = one 32-byte sector: a single 256-bit read-only load (LDG.E.256, sm_100+)
// instead of a 128-bit + a 32-bit request -- the learner kernel is bound by L1 wavefronts, and every
// lane of a warp gathers a different row.
__device__ __forceinline__ void ldg_row(const float *p, float (&r)[MDPP_Q_ROW]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                 : "l"(p));
}

// n-th (0-based) set bit of a 5-bit mask
__device__ __forceinline__ uint32_t nth_set_bit(uint32_t mask, uint32_t n) {
    // drop the n lowest set bits (n <= 4), then find-first-set; __fns() is a ~40-instruction emulation
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) mask = (n > k) ? (mask & (mask - 1u)) : mask;
    return (uint32_t)__ffs((int)mask) - 1u;
}

} // namespace mdpp
