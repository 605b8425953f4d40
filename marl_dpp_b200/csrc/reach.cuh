// reach.cuh -- deadlock precompute (-dd): backward reachability of "all cars finished" as a level-synchronous
// BACKWARD breadth-first search in one cooperative launch (sm_100a).
//
// Definition (new, next to the reference's stall heuristic; SURVEY.md fact 3, include/mdpp.h): joint state = per car
// (node, destination_index) or FINISHED; any unfinished car may take any action its legal filter allows under the
// encodestate(., 0) view (finished cars are not listed); a state is live iff "all finished" is reachable.
//
// reach_sweep_kernel (mdpp_kernels.cu) restates that definition directly -- every joint state re-examines its
// successors, sweep after sweep, until nothing changes: 27 sweeps x 28.4 M states for 4 cars, each with a host
// round trip for the convergence flag (30.7 ms).  Here every LIVE state is expanded exactly once, backwards:
//   * a reverse move table per car, built in shared memory from the step tables: for a car value t (node,
//     destination_index, or FINISHED) the (value v, action a) pairs with  v --a--> t;
//   * a level's frontier is a bitmap.  Each CTA takes 512-word chunks of it, compacts the set bits into a
//     shared-memory queue (so that every thread expands one state: balanced whatever the bit density) and, for
//     every car c and every reverse move (v, a) of its value, forms the predecessor p (car c at v, the others
//     unchanged), and -- unless p is already live -- checks with the table-driven legal filter that a is legal for
//     car c IN p (the other cars block boxes there).  If so p is live: one atomic OR into the result bitmap and,
//     when this thread set the bit, one into the next frontier;
//   * one grid barrier per level; the search ends, on the device, at the first level that finds nothing.
// Work: live states x cars x ~5 reverse moves, about what ONE sweep costs; no host round trip until the end.
// Bit-exact against the CPU brute force (tests/test_gpu_deadlock.py, test_gpu_fullsize.py) and against the sweeps.
#pragma once
#include "q_round.cuh"

namespace mdpp {

constexpr int kReachThreads = 512, kReachChunkWords = 512, kRevSlots = 12;
constexpr uint32_t kReachMaxV = 2u * 72u + 1u;

struct ReachArgs {
    uint32_t *reach;          // [words] result bitmap: bit j = joint state j can still reach "all cars finished"
    uint32_t *front[2];       // [words] frontier bitmaps: level l reads front[l & 1], writes front[(l + 1) & 1]
    uint32_t *counts;         // [3] states found per level (rotating)
    uint32_t *levels_out;     // number of levels the search ran
    uint32_t *errors;         // reverse-table overflow (never): counted, the search result would be incomplete
    uint32_t *barrier;        // the grid barrier's area (q_round.cuh grid_barrier)
    uint32_t words, n_joint, V, node_base, parking_node, rank_stride;
    const FastTabs *tabs;
};

template <int NC>
struct ReachSmem : StepTabs<NC> {
    uint16_t rev[NC][kReachMaxV][kRevSlots]; // v | a << 8
    uint8_t revcnt[NC][kReachMaxV + 3];
    uint16_t queue[kReachChunkWords * 32];   // bit offsets of the chunk's frontier states
    uint32_t qcount, newcount;
};

__device__ __forceinline__ uint32_t reach_field(uint32_t v, uint32_t V, uint32_t node_base, uint32_t parking) {
    return v == V - 1u ? make_field(parking, 0u, 1u, 7u) : make_field(node_base + (v >> 1), v & 1u, 0u, 5u);
}

// reverse moves of car C into its current value: predecessors of `idx` that differ in car C
template <int NC, int C>
__device__ __forceinline__ void reach_expand_car(const ReachSmem<NC> &S, const ReachArgs &ra, uint32_t idx, uint64_t cars,
                                                 const uint32_t (&digit)[NC], const uint32_t (&pw)[NC], uint32_t &found) {
    const uint32_t t = digit[C], n = S.revcnt[C][t];
    for (uint32_t k = 0; k < n; k++) {
        const uint32_t e = S.rev[C][t][k], v = e & 0xFFu, a = e >> 8;
        if (v == t) continue; // a move that changes nothing
        const uint32_t p = idx - t * pw[C] + v * pw[C];
        const uint32_t bit = 1u << (p & 31u);
        if (__ldcg(ra.reach + (p >> 5)) & bit) continue; // already live (a stale miss only costs the check below)
        const uint64_t pc = (cars & ~(0xFFFull << (12 * C))) |
                            ((uint64_t)reach_field(v, ra.V, ra.node_base, ra.parking_node) << (12 * C));
        const FastView fv = fast_view<NC, C>(S, ra.tabs, pc, 0, ra.rank_stride);
        if (!((fv.lm >> a) & 1u)) continue; // the other cars of p forbid the move
        if (!(atomicOr(ra.reach + (p >> 5), bit) & bit)) {
            atomicOr(ra.front[1] + (p >> 5), bit);
            found++;
        }
    }
    if constexpr (C + 1 < NC) reach_expand_car<NC, C + 1>(S, ra, idx, cars, digit, pw, found);
}

template <int NC>
__global__ void __launch_bounds__(kReachThreads)
reach_bfs_kernel(ReachArgs ra) {
    extern __shared__ uint4 smem_dyn[];
    ReachSmem<NC> &S = *reinterpret_cast<ReachSmem<NC> *>(smem_dyn);
    stage_step_tabs<NC, kReachThreads>(S, ra.tabs);
    if (threadIdx.x == 0) { S.qcount = 0u; S.newcount = 0u; }
    __syncthreads();
    const uint32_t V = ra.V, FIN = V - 1u;
    // reverse move table: thread (car, value t) collects every (value v, action a) of the step tables with v --a--> t
    for (uint32_t i = threadIdx.x; i < NC * V; i += kReachThreads) {
        const uint32_t c = i / V, t = i - c * V;
        uint32_t n = 0;
        for (uint32_t v = 0; v < FIN; v++) {
            const uint32_t f = reach_field(v, V, ra.node_base, ra.parking_node);
            for (uint32_t a = 0; a < MDPP_N_ACT; a++) {
                const uint32_t mv = S.next[c][f & 0xFFu][a];
                const uint32_t to = ((mv >> 8) & 1u) ? FIN : ((((mv & 127u) - ra.node_base) << 1) | ((mv >> 7) & 1u));
                if (to != t || v == t) continue;
                if (n < (uint32_t)kRevSlots) S.rev[c][t][n] = (uint16_t)(v | (a << 8));
                n++;
            }
        }
        if (n > (uint32_t)kRevSlots) {
            atomicAdd(ra.errors, 1u);
            n = kRevSlots;
        }
        S.revcnt[c][t] = (uint8_t)n;
    }
    uint32_t pw[NC];
    {
        uint32_t p = 1;
#pragma unroll
        for (int c = 0; c < NC; c++) { pw[c] = p; p *= V; }
    }
    __syncthreads();
    const uint32_t n_chunks = (ra.words + kReachChunkWords - 1) / kReachChunkWords;
    uint32_t bar = barrier_base_load(ra.barrier);
    uint32_t *cur = ra.front[0], *nxt = ra.front[1];
    ReachArgs la = ra; // front[1] of this copy is always "the next frontier" for reach_expand_car
    for (uint32_t level = 0;; level++) {
        la.front[1] = nxt;
        uint32_t found = 0;
        for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
            // ---- compact the chunk's frontier bits into the queue (one word per thread) ----
            const uint32_t w = chunk * kReachChunkWords + threadIdx.x;
            uint32_t bits = 0;
            if (threadIdx.x < kReachChunkWords && w < ra.words) {
                bits = __ldcg(cur + w);
                if (bits) cur[w] = 0u; // consumed: the buffer is the next frontier of the level after next
            }
            const uint32_t cnt = (uint32_t)__popc(bits);
            uint32_t incl = cnt; // warp-inclusive prefix sum, then one shared-memory atomic per warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if ((threadIdx.x & 31u) >= (uint32_t)o) incl += y;
            }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            uint32_t base = 0;
            if ((threadIdx.x & 31u) == 31u && total) base = atomicAdd(&S.qcount, total);
            base = __shfl_sync(0xFFFFFFFFu, base, 31);
            uint32_t slot = base + incl - cnt;
            while (bits) {
                const uint32_t b = (uint32_t)__ffs((int)bits) - 1u;
                bits &= bits - 1u;
                S.queue[slot++] = (uint16_t)(threadIdx.x * 32u + b);
            }
            __syncthreads();
            const uint32_t nq = S.qcount;
            // ---- every thread expands one frontier state at a time ----
            for (uint32_t i = threadIdx.x; i < nq; i += kReachThreads) {
                const uint32_t idx = chunk * (kReachChunkWords * 32u) + S.queue[i];
                if (idx >= ra.n_joint) continue;
                uint32_t digit[NC], rest = idx;
                uint64_t cars = 0;
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    digit[c] = rest % V;
                    rest /= V;
                    cars |= (uint64_t)reach_field(digit[c], V, ra.node_base, ra.parking_node) << (12 * c);
                }
                reach_expand_car<NC, 0>(S, la, idx, cars, digit, pw, found);
            }
            __syncthreads();
            if (threadIdx.x == 0) S.qcount = 0u;
            __syncthreads();
        }
        // ---- end of the level: how many states did the whole grid find? ----
        found = __reduce_add_sync(0xFFFFFFFFu, found);
        if ((threadIdx.x & 31u) == 0 && found) atomicAdd(&S.newcount, found);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (S.newcount) atomicAdd(ra.counts + level % 3u, S.newcount);
            S.newcount = 0u;
            if (blockIdx.x == 0) ra.counts[(level + 1u) % 3u] = 0u; // last read two barriers ago
        }
        grid_barrier(ra.barrier, bar += gridDim.x);
        const uint32_t n_new = __ldcg(ra.counts + level % 3u);
        if (n_new == 0u) {
            if (blockIdx.x == 0 && threadIdx.x == 0) *ra.levels_out = level + 1u;
            barrier_base_store(ra.barrier, bar);
            break;
        }
        uint32_t *t2 = cur;
        cur = nxt;
        nxt = t2;
    }
}

// the goal ("all cars finished" = the last joint state) seeds the result and the first frontier
__global__ void reach_seed_kernel(uint32_t *reach, uint32_t *front0, uint32_t goal) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        reach[goal >> 5] = 1u << (goal & 31u);
        front0[goal >> 5] = 1u << (goal & 31u);
    }
}

} // namespace mdpp
