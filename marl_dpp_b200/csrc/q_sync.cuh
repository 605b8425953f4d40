// q_sync.cuh -- Q-replica averaging across the GPUs of one node, one kernel over NVLink peer memory (sm_100a).
//
// every GPU trains its own replica of the Q table on its env shard; every `interval` rounds the
// replicas are averaged.  The reference has no counterpart (one process).  For the tables of the headline
// scenario (208 KiB) a library all-reduce is pure latency -- ncclAllReduce(AVG) measured 33 us, two whole
// rounds -- so the exchange is ONE kernel of our own, a member of the learner's kernel chain:
//
//   * every rank's table lives in a peer-mapped exchange buffer (CUDA IPC; mdpp_sync_alloc / mdpp_sync_open),
//     followed by a small flag area;
//   * CTA b of every rank owns the same slice of the table.  It tells CTA b of every peer "my table is final"
//     (a store into the peer's flag area), waits for the same word from every peer, reads the slice from
//     ALL ranks' buffers -- every load of a thread in flight at once: one NVLink round trip -- adds them in rank
//     order and divides by the rank count, so that every rank computes the same bits;
//   * a second handshake ("I have read your slice") precedes the only write: the slice of the own table.  Nobody
//     overwrites data a peer may still be reading, no second buffer and no copy back.
//
// Flags only ever grow (the epoch of the call), so nothing is reset between calls.  One-shot: each rank pulls
// world x table bytes, right for tables of a few MB at NVLink bandwidth; larger tables (4-5 cars) take the
// caller's NCCL path.
//
// Overlap: inside mdpp_q_train_rounds the kernel runs on a stream of its own, forked from the caller's stream after
// the last pull and joined again before the next kernel that touches the table (mdpp_api.inl q_sync_launch /
// sync_join).  The next env kernel never reads the table while every env explores, so a pure-exploration round runs
// WHILE the replicas are exchanged; the pulls of that round wait for the join.  Its CTAs are small enough (128
// threads, <= 56 registers) to be resident beside the env kernel's 1024-thread CTA on the same SM.
#pragma once
#include "env_core.cuh"

namespace mdpp {

// CTA shape: the exchange must be RESIDENT BESIDE the env kernel it overlaps.  A round-kernel CTA holds 1024 threads
// x 54-56 registers of its SM, i.e. 8192 registers stay free: 128 threads x 64 registers.  (The first version used
// 512-thread CTAs of 52 registers: on the SMs that hosted one, the env kernel's CTA could not start before the
// exchange had finished, and every exchange sat on the critical path -- +25 us per round at interval 1.)
constexpr int kSyncThreads = 128, kSyncPer = 4, kSyncMaxCtas = 148, kSyncMaxRanks = MDPP_SYNC_MAX_RANKS;
constexpr size_t kSyncFlagWords = 2u * kSyncMaxCtas * kSyncMaxRanks; // [phase][cta][rank]
constexpr size_t kSyncFlagBytes = kSyncFlagWords * 4u;

struct SyncArgs {
    float *table[kSyncMaxRanks];    // every rank's table; [rank] is the own one
    uint32_t *flags[kSyncMaxRanks]; // every rank's flag area
    uint32_t rank, world, epoch;
    uint32_t n4;                    // float4 elements of the table
    uint32_t per;                   // chunks of kSyncThreads float4 per CTA (1..kSyncPer)
    unsigned long long timeout_ns;  // a peer that never arrives must not hang the GPU: give up, count an error
    unsigned long long *errors;     // Control::stats[0].internal_errors
    uint32_t mode;                  // MDPP_SYNC_MODE bit 0: system-scope fences around every flag (comparison runs)
};

__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One phase of the handshake between the CTAs `b` of all ranks: thread t < world signals rank t and waits for it.
//
// Flags travel as relaxed system-scope stores and are polled with relaxed system-scope loads, WITHOUT fences:
//   * "my table is final": the table was written by kernels that completed before this one started; a peer reads it
//     through NVLink from the owner's L2, the point of coherence, where those writes are;
//   * "I have read your slice": every thread consumed the values it loaded (the sums are formed) before the CTA
//     barrier in front of the signal, so the loads have completed; the waiting side writes its slice only after its
//     poll loop has exited (stores are not speculated).
// This is the established shape of latency-critical all-reduce barriers over NVLink peer memory (volatile flag
// stores / loads at the start barrier, no fence before a kernel's final barrier).  System-scope fences made each
// phase ~5 us more expensive: 20.4 us per exchange on 2 GPUs against 9.8 without (ncclAllReduce of the same table:
// 18.1).  MDPP_SYNC_MODE=1 selects the fenced form (release / acquire around every flag) for comparison.
__device__ __forceinline__ void sync_handshake(const SyncArgs &a, uint32_t phase, uint32_t b) {
    __syncthreads(); // everything this CTA did before (its reads of the peers' slices) is ordered before the signal
    if (threadIdx.x < a.world) {
        const uint32_t t = threadIdx.x;
        const size_t slot = ((size_t)phase * kSyncMaxCtas + b) * kSyncMaxRanks;
        const uint32_t *mine = a.flags[a.rank] + slot + t;
        const bool fences = (a.mode & 1u) != 0u;
        if (fences) asm volatile("fence.acq_rel.sys;" ::: "memory");
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.flags[t] + slot + a.rank), "r"(a.epoch) : "memory");
        uint32_t v, spins = 0;
        unsigned long long t0 = 0;
        do {
            asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((++spins & 0xFFFu) == 0u) { // a peer that never arrives: give up after timeout_ns, count an error
                const unsigned long long now = global_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > a.timeout_ns) {
                    atomicAdd(a.errors, 1ull);
                    break;
                }
            }
        } while ((int32_t)(v - a.epoch) < 0);
        if (fences) asm volatile("fence.acq_rel.sys;" ::: "memory");
    }
    __syncthreads();
}

// the exchange of slice `b` as rank a.rank: a.per (<= kSyncPer) chunks of kSyncThreads float4 per CTA
template <int WORLD>
__device__ __forceinline__ void sync_slice(const SyncArgs &a, uint32_t b) {
    sync_handshake(a, 0u, b); // every rank's table is final
    const uint32_t base = b * (kSyncThreads * a.per) + threadIdx.x;
    float4 m[kSyncPer];
    const float w = (float)WORLD;
#pragma unroll
    for (int k = 0; k < kSyncPer; k++) {
        const uint32_t i = base + (uint32_t)k * kSyncThreads;
        if ((uint32_t)k >= a.per || i >= a.n4) continue;
        float4 v[WORLD]; // every rank's copy of this float4: all loads in flight at once (one NVLink round trip)
#pragma unroll
        for (int r = 0; r < WORLD; r++) v[r] = ld_relaxed_sys_f4(reinterpret_cast<const float4 *>(a.table[r]) + i);
        float4 s = v[0]; // rank order, explicit roundings: the same bits on every rank
#pragma unroll
        for (int r = 1; r < WORLD; r++) {
            s.x = __fadd_rn(s.x, v[r].x);
            s.y = __fadd_rn(s.y, v[r].y);
            s.z = __fadd_rn(s.z, v[r].z);
            s.w = __fadd_rn(s.w, v[r].w);
        }
        m[k] = make_float4(__fdiv_rn(s.x, w), __fdiv_rn(s.y, w), __fdiv_rn(s.z, w), __fdiv_rn(s.w, w));
    }
    sync_handshake(a, 1u, b); // every rank has read this slice of every table
#pragma unroll
    for (int k = 0; k < kSyncPer; k++) {
        const uint32_t i = base + (uint32_t)k * kSyncThreads;
        if ((uint32_t)k < a.per && i < a.n4) reinterpret_cast<float4 *>(a.table[a.rank])[i] = m[k];
    }
}

template <int WORLD>
__global__ void __launch_bounds__(kSyncThreads)
q_sync_kernel(SyncArgs a) {
    // Release the successor FIRST: the next env kernel then starts as early as it does in an unsynchronised chain --
    // while the last pull still runs -- and overlaps the pull AND the exchange (it touches records and its own
    // parity of the mark regions only, and waits for this grid before it reads the table or exits).  With the
    // release after the wait the env kernel could not start before the pull had completed: +20 us per exchange on
    // 2 GPUs, the pull's overlap lost and a launch latency exposed.
    pdl_launch_dependents();
    pdl_wait();              // launched in a kernel chain (mdpp_q_sync_replicas): the last pull has completed
    sync_slice<WORLD>(a, blockIdx.x);
}

// Every rank of a group that lives in ONE process on ONE GPU, as a single cooperative launch (all CTAs co-resident
// by construction): CTA x is slice x % ctas_per_rank of rank x / ctas_per_rank and runs the very code a rank's own
// kernel runs, handshakes included.  Kernels that wait on one another must not be separate launches on one GPU
// (nothing guarantees that they run at the same time), so this is how the protocol is exercised where there are
// fewer GPUs than ranks (mdpp_sync_emulate; tests/test_gpu_sync.py).
struct SyncGroupArgs {
    SyncArgs r[kSyncMaxRanks];
    uint32_t ctas_per_rank;
};
template <int WORLD>
__global__ void __launch_bounds__(kSyncThreads)
q_sync_group_kernel(SyncGroupArgs g) {
    sync_slice<WORLD>(g.r[blockIdx.x / g.ctas_per_rank], blockIdx.x % g.ctas_per_rank);
}

} // namespace mdpp
