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
//     (a release store into the peer's flag area), waits for the same word from every peer, reads the slice from
//     ALL ranks' buffers -- every load of a thread in flight at once: one NVLink round trip -- adds them in rank
//     order and divides by the rank count, so that every rank computes the same bits;
//   * a second handshake ("I have read your slice") precedes the only write: the slice of the own table.  Nobody
//     overwrites data a peer may still be reading, no second buffer and no copy back.
//
// Flags only ever grow (the epoch of the call), so nothing is reset between calls.  One-shot: each rank pulls
// world x table bytes, right for tables of a few MB at NVLink bandwidth; larger tables (4-5 cars) take the
// caller's NCCL path.
//
// Chain membership: the kernel releases its successor at once and then waits for its predecessor (the last pull:
// the local table is final) with griddepcontrol.wait.  The next env kernel never reads the table before its
// own griddepcontrol.wait, so a pure-exploration round runs WHILE the replicas are exchanged and the exchange
// disappears from the critical path; the pulls of that round start after the env kernel, hence after this one.
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
    uint32_t mode;                  // experiments (MDPP_SYNC_MODE): 0 = acquire polling, 1 = relaxed polling + one fence,
                                    // 2 = 1 without waiting in the second handshake (UNSAFE, timing only)
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
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

// one phase of the handshake between the CTAs `b` of all ranks: thread t < world signals rank t and waits for it
__device__ __forceinline__ void sync_handshake(const SyncArgs &a, uint32_t phase, uint32_t b) {
    __syncthreads(); // everything this CTA did before (its reads of the peers' slices) is ordered before the signal
    if (threadIdx.x < a.world) {
        const uint32_t t = threadIdx.x;
        const size_t slot = ((size_t)phase * kSyncMaxCtas + b) * kSyncMaxRanks;
        const uint32_t *mine = a.flags[a.rank] + slot + t;
        if ((a.mode & 3u) == 0u) {
            __threadfence_system();
            st_release_sys(a.flags[t] + slot + a.rank, a.epoch);
            const unsigned long long t0 = global_ns();
            while ((int32_t)(ld_acquire_sys(mine) - a.epoch) < 0) {
                if (global_ns() - t0 > a.timeout_ns) {
                    atomicAdd(a.errors, 1ull);
                    break;
                }
            }
        } else {
            // one release fence, a relaxed store; poll with relaxed loads and fence once after the flag arrived
            // (variant 3: no fence at all -- the tables were written by kernels that completed before this one started,
            // and the second signal follows instructions that consumed every loaded value)
            const bool fences = (a.mode & 3u) != 3u;
            if (fences) asm volatile("fence.acq_rel.sys;" ::: "memory");
            asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.flags[t] + slot + a.rank), "r"(a.epoch) : "memory");
            if (!((a.mode & 3u) == 2u && phase == 1u)) {
                uint32_t v, spins = 0;
                unsigned long long t0 = 0;
                do {
                    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
                    if ((++spins & 0xFFFu) == 0u) {
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
        }
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
    if (a.mode & 8u) { // experiment: wait first, release after (the first version's order)
        pdl_wait();
        pdl_launch_dependents();
    } else {
        pdl_launch_dependents();
        pdl_wait();          // the last pull has completed: the own table is final
    }
    if (a.mode & 4u) return; // experiment: no exchange at all, only the chain membership
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
