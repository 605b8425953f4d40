/*
 * mdpp.h -- C ABI of the B200-native (sm_100a) batched MARL-DPP simulator + tabular learner.
 *
 * Drop-in boundary.  The reference (DosepackAIR/MARL-DPP) is pure Python with no FFI; its
 * boundary for this path is two duck-typed contracts:
 *   (1) the gym-style env  `Environment(env_conf)`  -- reference src/env_gym.py:4-99
 *   (2) the algorithm plugin `RL(**params, sess=, env_config=)` -- reference app.py:42-72,91-96
 * The Python facade in marl_dpp_b200/ keeps both contracts and calls the entry points below
 * through ctypes (INTEGRATION.md shows the stub a reference maintainer would add).  Each entry
 * point names the reference interface it replaces.
 *
 * Conventions: every pointer marked DEV is caller-owned device memory; `stream` is a
 * cudaStream_t passed as void*; calls are asynchronous on that stream unless the name ends in
 * `_host`; the return value is 0 on success, otherwise an MDPP_E* code with text available
 * from mdpp_last_error_string() (thread local).  The library allocates no device memory: the
 * caller provides one workspace blob sized by mdpp_workspace_bytes().  One host thread per
 * handle.  There is no CPU fallback: without a CUDA device every compute entry point fails
 * with MDPP_ENODEVICE.
 */
#ifndef MDPP_H
#define MDPP_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDPP_ABI_VERSION 1
#define MDPP_MAX_CARS 5   /* 12 bits per car in one 64-bit word; reference DQN TOTAL_CARS=5, settings.py:7 */
#define MDPP_NODES 72     /* 2 floors x 9 boxes x 4 headings, reference src/configuration.py:100-118 */
#define MDPP_BOXES 18
#define MDPP_NONE 255
#define MDPP_MAX_DEST 8
#define MDPP_Q_ROW 8      /* floats per Q row: S,L,R,F,T + 3 padding slots = one 32-byte sector */

enum {
    MDPP_OK = 0,
    MDPP_EINVAL = 1,
    MDPP_ENODEVICE = 2,
    MDPP_ECUDA = 3,
    MDPP_EWORKSPACE = 4
};

/* action indices on this ABI: the reference's 6 actions minus 'B', which its legal filter always
 * removes (reference src/env_gym.py:158-159); order S,L,R,F,T = reference argMax tie-break order */
enum { MDPP_A_S = 0, MDPP_A_L = 1, MDPP_A_R = 2, MDPP_A_F = 3, MDPP_A_T = 4, MDPP_N_ACT = 5 };
#define MDPP_ACT_SKIP 0xFF
#define MDPP_ACT_GREEDY 0xFE /* mdpp_q_substep `forced`: take the argmax, no coin */

/* reward kinds, reference src/env_gym.py:101-112 select_reward_function */
enum { MDPP_REWARD_Q = 0, MDPP_REWARD_DQN = 1 };

/* Static tables of one (layout block, initial state) pair.  Built on the host from the layout
 * dict (reference src/layout.py schema) by marl_dpp_b200/tables.py; replaces Configuration
 * (reference src/configuration.py:44-145) and the per-episode BFS memo (src/env_gym.py:443-483). */
typedef struct {
    int32_t abi_version;
    int32_t n_cars;          /* 1..MDPP_MAX_CARS */
    int32_t n_local_nodes;   /* nodes in the state-id universe (36 for a one-floor block) */
    int32_t n_dest_nodes;    /* distinct destination nodes */
    int32_t n_dest_boxes;    /* distinct destination boxes */
    int32_t n_src_boxes;     /* boxes another car can be listed in (9 for a one-floor block) */
    int32_t parking_node;    /* no_box_list[0] + 'W', reference src/env_gym.py:439 */
    int32_t tcdi;            /* training_complete_destination_index */
    int32_t bad_states_on;   /* give_bad_state_reward, reference src/env_gym.py:71-83 */
    int32_t step_cap;        /* agent-steps per episode before a forced reset; 0 = none (qlearning.py:274) */
    uint8_t node_local[MDPP_NODES];          /* node -> local index, MDPP_NONE outside the universe */
    uint8_t fwd[MDPP_NODES];                 /* 'F' successor or MDPP_NONE */
    uint8_t base_legal[MDPP_NODES];          /* 5-bit mask over S,L,R,F,T */
    uint8_t bfs_len[MDPP_MAX_DEST][MDPP_NODES]; /* len(BFS actions) node -> destination node */
    uint8_t dest_node[MDPP_MAX_DEST];        /* destination-node index -> node id */
    uint8_t dest_node_box[MDPP_MAX_DEST];    /* destination-node index -> destination-box index */
    uint8_t dest_box[MDPP_MAX_DEST];         /* destination-box index -> box18 */
    uint8_t src_box_idx[MDPP_BOXES];         /* box18 -> index in the others universe or MDPP_NONE */
    uint32_t sel_mask[MDPP_MAX_DEST];        /* per destination box: 18-bit selective-block mask */
    uint8_t car_dest[MDPP_MAX_CARS][2];      /* per car, per destination_index: destination-node index */
    uint8_t start_box[MDPP_MAX_CARS];        /* box18 of the initial state */
    uint8_t start_di[MDPP_MAX_CARS];         /* initial destination_index */
    int64_t n_states;        /* rows of the Q table */
    int64_t n_box_states;    /* bits of the bad-state bitmap */
} mdpp_config;

/* Learner parameters (reference src/algorithms/qlearning.py:20-27 and src/utils/utils.py:33-52). */
typedef struct {
    float alpha, discount;      /* 0.9, 0.15 */
    int32_t episodes;           /* per-env episode budget, 0..65535 (the per-env counter is 16 bits of the record;
                                   larger values are rejected with MDPP_EINVAL); an env that finished it is frozen */
    int32_t eps_threshold[5];   /* int(episodes*{0.4,0.6,0.7,0.85,0.99}) */
    uint32_t eps_q16[6];        /* epsilon * 65536 per schedule segment: {1,.8,.5,.3,.15,0} */
    uint32_t seed;
    uint32_t flags;             /* MDPP_LEARN_* */
} mdpp_learner;

#define MDPP_LEARN_FROZEN 1u /* select + step but no TD update: validate_given_state (qlearning.py:384-437) */
#define MDPP_LEARN_MODE0 2u  /* encodestate(car, 0): finished cars are not listed (qlearning.py:342) */
#define MDPP_LEARN_DETECT 4u /* stall heuristic of detect_deadlocks_v0 (qlearning.py:347-360) */
#define MDPP_LEARN_OBSERVED 8u /* the state was already encoded for this agent-step (mdpp_step_car peek or
                                  mdpp_dqn_observe): do not advance the ghost counters a second time */
#define MDPP_LEARN_RANDOM_ORDER 16u /* train_given_state_random_order (qlearning.py:241-295): in every sub-step each
                                  env draws the car that acts (random.randint, :267; a finished car = no agent-step,
                                  :268-269) and the episode may end after any agent-step (:291).  `car` of
                                  mdpp_q_substep is then the index of the sub-step within the round (a round is
                                  n_cars sub-steps; Philox stream 4 + index: word 0 car, word 1 decision, word 2
                                  headings of the next episode).  The forced byte carries the car in its high
                                  nibble (0xF = own draw) and the action in the low one (0..4, 0xE greedy, 0xF own
                                  decision), so MDPP_ACT_SKIP / MDPP_ACT_GREEDY keep their meaning. */

/* counters accumulated on the device, fetched with mdpp_get_stats_host */
typedef struct {
    uint64_t agent_steps;       /* update_car calls, the BASELINE metric's unit */
    uint64_t q_updates;
    uint64_t episodes_done;
    uint64_t episodes_capped;
    int64_t reward_sum;         /* sum of integer q_rewards (dqn: x100) */
    uint64_t explore_steps;
    uint64_t touched_keys;      /* distinct (state, action) keys finalised */
    uint64_t internal_errors;   /* index guards tripped inside the kernels (a key or state id out of range): always 0 */
} mdpp_stats;

typedef struct mdpp_handle mdpp_handle;

/* ---- lifetime ---------------------------------------------------------------------------- */
int64_t mdpp_workspace_bytes(const mdpp_config *cfg, int64_t n_envs);
/* env_id_base: global id of env 0 (rank * n_envs for data-parallel shards) -- keys the RNG so that
 * results do not depend on the rank count.  bad_bitmap_host: n_box_states bits or NULL. */
int mdpp_create(const mdpp_config *cfg, int64_t n_envs, int64_t env_id_base,
                void *workspace /*DEV*/, int64_t workspace_bytes,
                const uint32_t *bad_bitmap_host, void *stream, mdpp_handle **out);
int mdpp_destroy(mdpp_handle *h);
const char *mdpp_last_error_string(void);
int mdpp_abi_version(void);
/* device pointer to the packed env records, 16 bytes per env (see DESIGN.md "Data layout").  Read-only for the
 * caller: the library keeps a host-side bound on the episode counters (it decides whether a training round can
 * skip its catch-up kernel); change records through mdpp_reset / the stepping entry points only. */
void *mdpp_env_records(mdpp_handle *h);
/* device pointer to [n_states] u32: bit a set = key (state, a) has been updated (qvalue file writer) */
void *mdpp_visited_bits(mdpp_handle *h);
/* device pointer to [n_envs] u32: state id at which MDPP_LEARN_DETECT flagged env e, 0xFFFFFFFF = none */
void *mdpp_deadlock_states(mdpp_handle *h);
/* number of learner kernels (env / round / pull / copy) launched through this handle so far */
uint64_t mdpp_launch_count(mdpp_handle *h);

/* ---- env: reset / initialize (reference src/env_gym.py:643-672,331-352) -------------------- */
/* headings: [n_envs][n_cars] u8 in 0..3 (N,E,S,W) or NULL => Philox.  env_mask: [n_envs] u8, NULL = all.
 * Resets cars to the initial state; episode counters are zeroed only when zero_counters != 0. */
int mdpp_reset(mdpp_handle *h, const uint8_t *env_mask /*DEV*/, const uint8_t *headings /*DEV*/,
               uint32_t seed, int zero_counters, void *stream);

/* ---- env: one agent-step of car `car_slot` in every env, externally supplied actions --------
 * Replaces encodestate + getlegalactions + get_successor_state + reward_function + update_car
 * (reference src/algorithms/qlearning.py:218-229) == Environment.step (src/env_gym.py:85-99).
 * deadlock_id as in encodestate (src/env_gym.py:281); outputs are optional (NULL = skip):
 * status: 0 stepped, 1 car already done (nothing happens), 2 illegal action (nothing happens). */
int mdpp_step_car(mdpp_handle *h, int car_slot, int deadlock_id, int reward_kind,
                  const uint8_t *actions /*DEV [n_envs]*/,
                  uint32_t *state_idx /*DEV*/, uint32_t *next_idx /*DEV*/, float *reward /*DEV*/,
                  uint8_t *legal /*DEV*/, uint8_t *status /*DEV*/, uint8_t *all_done /*DEV*/,
                  void *stream);
/* same call with HOST buffers (pinned or pageable): copies in, steps, copies out, synchronises */
int mdpp_step_car_host(mdpp_handle *h, int car_slot, int deadlock_id, int reward_kind,
                       const uint8_t *actions, uint32_t *state_idx, uint32_t *next_idx, float *reward,
                       uint8_t *legal, uint8_t *status, uint8_t *all_done, void *stream);

/* ---- string-API support: the reference's getlegalactions / get_successor_state / reward_function take
 * state STRINGS, not the env (src/env_gym.py:188-265,485-620).  Entry i evaluates action[i] from state
 * state_idx[i]; only the reward's clash term reads env i's cars.  action may be NULL (legal only).
 * mdpp_update_car is the bare update_car (src/env_gym.py:386-441); nodes[e] = MDPP_NONE skips env e. */
int mdpp_query_transitions(mdpp_handle *h, int car_slot, int reward_kind, const uint32_t *state_idx /*DEV*/,
                           const uint8_t *action /*DEV*/, int64_t n, uint8_t *legal /*DEV*/,
                           uint32_t *next_idx /*DEV*/, float *reward /*DEV*/, uint8_t *next_legal /*DEV*/,
                           void *stream);
int mdpp_update_car(mdpp_handle *h, int car_slot, const uint8_t *nodes /*DEV [n_envs]*/, void *stream);

/* ---- env: lockstep rollout under the uniform-random legal policy (reference eps = 1 behaviour,
 * src/algorithms/qlearning.py:99-102) with Philox4x32 (MDPP_PHILOX_ROUNDS = 7 rounds) and auto-reset.  `rounds` full
 * round-robin rounds (qlearning.py:211-233) per launch; outputs (optional) are SoA
 * [rounds][n_cars][n_envs]: action (MDPP_ACT_SKIP when the car did not move), reward, env-done
 * flag after the step, next-state id. */
int mdpp_rollout_random(mdpp_handle *h, int rounds, uint64_t first_round, uint32_t seed, int reward_kind,
                        uint8_t *action /*DEV*/, float *reward /*DEV*/, uint8_t *done /*DEV*/,
                        uint32_t *next_idx /*DEV*/, void *stream);

/* ---- learner: epsilon-greedy select + step + TD update ("mark and pull") ---------------------
 * One sub-step = car `car_slot` of every env: getaction (qlearning.py:74-105), the env step and
 * update (qlearning.py:107-116), all envs reading the Q table as it was when the sub-step began.
 * (state, action) and that snapshot determine the updated value completely, so every env that
 * updates one key computes the same value: mdpp_q_substep only MARKS the keys its envs touch and
 * mdpp_q_finalize computes the reference update once per marked key and writes it (exactly the
 * reference's sequential update when there is one env).  forced: optional [n_envs] u8,
 * MDPP_ACT_SKIP = choose by coin/argmax, MDPP_ACT_GREEDY = argmax, otherwise take that action as an
 * exploration step (replays a recorded decision stream).  Optional outputs as in mdpp_step_car plus
 * the chosen action.
 * Q rows are 8 floats: slots 0..4 = Q(s, S/L/R/F/T); slots 5..7 are padding to one 32-byte sector
 * (carried through unchanged).  q is 32-byte aligned. */
int mdpp_q_substep(mdpp_handle *h, float *q /*DEV [n_states][8]*/, const mdpp_learner *lp, int car_slot,
                   uint64_t round, const uint8_t *forced /*DEV*/, uint8_t *action_out /*DEV*/,
                   uint32_t *state_idx /*DEV*/, float *reward /*DEV*/, void *stream);
int mdpp_q_finalize(mdpp_handle *h, float *q /*DEV*/, const mdpp_learner *lp, void *stream);
/* `rounds` full rounds (n_cars x (sub-step + finalize)) starting at round index first_round.  The kernels of
 * consecutive sub-steps overlap through programmatic dependent launch; small tables ping-pong between q and a
 * workspace copy and are back in q when the call's work completes.  With flags == 0 on 1- and 2-car tables a
 * whole round is one env-kernel launch while every env is certainly in an epsilon = 1 segment (a host-side
 * bound on the episode counters, tightened at every mdpp_get_stats_host: fetch the counters now and then), and
 * up to 8 such rounds share one launch, their updates following in two more.  Large tables (3-5 cars) keep one
 * touched bitmap per sub-step and a MIRROR of the table in the workspace; a pull writes the keys marked in its own
 * and in the previous sub-step from one copy into the other, and while every env explores all cars of several
 * rounds are stepped in one pass over the batch.  The result does not depend on how the rounds are grouped into
 * calls.  Every learner call leaves the current table in q. */
int mdpp_q_train_rounds(mdpp_handle *h, float *q /*DEV*/, const mdpp_learner *lp, uint64_t first_round,
                        int rounds, void *stream);
/* Large tables only (a no-op otherwise): between learner calls the mirror is a complete copy of q, which saves
 * copying the table at the start of every call.  A caller that WRITES q itself between two learner calls (loads a
 * qvalue file, averages replicas with a library all-reduce) says so before its next learner call; the library then
 * copies the table once.  Passing a different q pointer than the previous call has the same effect. */
int mdpp_q_table_written(mdpp_handle *h);
/* end-to-end variant: learner parameters from host memory, stats copied back, synchronised */
int mdpp_q_train_rounds_host(mdpp_handle *h, float *q /*DEV*/, const mdpp_learner *lp_host,
                             uint64_t first_round, int rounds, mdpp_stats *stats_host, void *stream);

/* Profiling hook: launches ONLY the env-side kernel of one fused round (1- and 2-car tables, flags == 0) so
 * that bench.py can time the dominant kernel alone with CUDA events.  The marks it leaves are never
 * pulled: use it on a handle you do not train further. */
int mdpp_q_profile_env_kernel(mdpp_handle *h, float *q /*DEV*/, const mdpp_learner *lp, uint64_t round, void *stream);
/* the same for the multi-round launch: ONLY the env kernel of `rounds` (2..8) pure-exploration rounds */
int mdpp_q_profile_env_rounds(mdpp_handle *h, float *q /*DEV*/, const mdpp_learner *lp, uint64_t round, int rounds,
                              void *stream);

/* ---- multi-GPU: Q-replica averaging over NVLink peer memory ------------------------------------------
 * No reference counterpart (the reference is one process): every GPU of a node trains its own
 * replica on its env shard (env_id_base keys the RNG) and every `interval` rounds the replicas are averaged --
 * sum in rank order / world, the same bits on every rank -- by ONE kernel that reads every rank's table through
 * peer-mapped memory (csrc/q_sync.cuh).  The table must therefore live in an exchange buffer:
 *   mdpp_sync_buffer_bytes   bytes of the buffer for a scenario: the Q table [n_states][8] floats at offset 0,
 *                            then MDPP_SYNC_FLAG_BYTES of flags (zero-initialised by mdpp_sync_alloc);
 *   mdpp_sync_alloc          cudaMalloc on the current device (the ONE device allocation this library makes:
 *                            legacy CUDA IPC needs memory it allocated itself) + the 64-byte IPC handle the other
 *                            ranks open; mdpp_sync_free releases it;
 *   mdpp_sync_open / _close  map / unmap a peer's buffer (cudaIpcOpenMemHandle with lazy peer access);
 *   mdpp_sync_attach         bufs[r] = rank r's buffer as seen from this process (bufs[rank] = the own one; in a
 *                            single process plain device pointers do).  interval > 0: mdpp_q_train_rounds[_host]
 *                            averages before every round R > 0 with R % interval == 0 -- `q` must then be
 *                            bufs[rank] -- as a member of its kernel chain: a pure-exploration round runs while
 *                            the replicas are exchanged.  Every rank must make the same calls in the same order.
 *   mdpp_q_sync_replicas     one averaging now (all ranks call it; asynchronous on `stream`).
 * One-shot exchange, for tables up to MDPP_SYNC_MAX_TABLE_BYTES (1-2 cars); mdpp_sync_attach
 * fails with MDPP_EINVAL beyond that and the caller falls back to a library all-reduce. */
#define MDPP_SYNC_MAX_RANKS 8
#define MDPP_SYNC_FLAG_BYTES (2 * 148 * MDPP_SYNC_MAX_RANKS * 4)
#define MDPP_SYNC_MAX_TABLE_BYTES (148 * 128 * 4 * 16)
#define MDPP_IPC_HANDLE_BYTES 64
int64_t mdpp_sync_buffer_bytes(const mdpp_config *cfg);
int mdpp_sync_alloc(int64_t bytes, void **dev_ptr_out, uint8_t *ipc_handle_out /*[MDPP_IPC_HANDLE_BYTES]*/);
int mdpp_sync_free(void *dev_ptr);
int mdpp_sync_open(const uint8_t *ipc_handle /*[MDPP_IPC_HANDLE_BYTES]*/, void **peer_ptr_out);
int mdpp_sync_close(void *peer_ptr);
int mdpp_sync_attach(mdpp_handle *h, int rank, int world, void *const *bufs /*HOST array [world] of DEV pointers*/,
                     int interval);
int mdpp_q_sync_replicas(mdpp_handle *h, float *q /*DEV, == bufs[rank]*/, void *stream);
/* number of replica averagings launched through this handle so far */
uint64_t mdpp_sync_count(mdpp_handle *h);
/* Fewer GPUs than ranks (tests): the averaging of ALL ranks of a group whose handles live in this process, as one
 * cooperative launch that runs every rank's slices with the same code, handshakes included.  (Kernels that wait
 * on one another must never be separate launches on one GPU.)  handles[r] must be attached as rank r of `world`. */
int mdpp_sync_emulate(mdpp_handle *const *handles, int world, void *stream);

/* ---- DQN feed: the env-facing half of the noisy double DQN's collection loop (reference
 * src/algorithms/dqn_open_source.py:490-526); the MLP stays in PyTorch.  observe = encodestate(car, 1)
 * (ghost side effect included) -> RL.convert(state) as 80 {0,1} bytes per env (:245-313) + give_mask
 * (:200-209) + state id.  act = env.step(state, action, car) on the observed state -> convert(next_state),
 * reward, give_mask(next_state), env-done flag, status as in mdpp_step_car; with auto_reset the last car
 * slot starts the next episode (Philox headings, key (seed, env id), counter (round, 2)).
 * features / next_features: [n_envs][80] u8, 16-byte aligned; rows of envs that did not move are zero. */
#define MDPP_DQN_FEATURES 80
int mdpp_dqn_observe(mdpp_handle *h, int car_slot, uint8_t *features /*DEV*/, uint8_t *legal /*DEV*/,
                     uint32_t *state_idx /*DEV*/, void *stream);
int mdpp_dqn_act(mdpp_handle *h, int car_slot, int reward_kind, const uint8_t *actions /*DEV*/,
                 uint8_t *next_features /*DEV*/, float *reward /*DEV*/, uint8_t *next_legal /*DEV*/,
                 uint8_t *done /*DEV*/, uint8_t *status /*DEV*/, int auto_reset, uint64_t round, uint32_t seed,
                 void *stream);

/* ---- DQN replay memory and Double-DQN target (reference src/algorithms/dqn_open_source.py:315-394) -----------
 * remember() + the `deque(OrderedSet(self.memory))` de-duplication of replay() (:322,:343): the memory is a SET of
 * samples (state, action, next state, reward).  Batched, the compact key  state id * 5 + action  enumerates the
 * samples (next state and reward follow from the key under legal actions), so the set is one presence bit per key
 * plus a per-key payload, all caller-owned device arrays: present [n_keys/32] u32 (zeroed by the caller), m_next
 * [n_keys] u32, m_reward [n_keys] f32, m_mask [n_keys] u8 (give_mask of the next state as a bit mask), m_features
 * [n_keys][160] u8 = convert(state) | convert(next_state) or NULL, counters [3] u64 (zeroed by the caller): samples
 * offered, new keys, duplicates whose stored (next state, reward) differ (distinct tuples in the reference; none
 * occur under legal actions).  n_keys = n_states * 5 rounded up to a multiple of 32.
 *   mdpp_replay_insert   one sub-step's samples (the outputs of mdpp_dqn_observe / mdpp_dqn_act; status[i] != 0 = no
 *                        sample) into the set: one atomic OR per sample, the first of a key stores the payload;
 *   mdpp_replay_compact  list[rank] = the keys of the set in ascending order, *count = its size (the minibatch
 *                        of :347-350 is `count choose batch_size` positions of that list);
 *   mdpp_dqn_target      Double-DQN target (:360-379) of a minibatch: a* = argmax of the ONLINE network's Q(s', .)
 *                        over the valid actions, target = reward + discount * Q_target(s', a*) with the reference's NumPy
 *                        arithmetic (float32 product, float64 sum with the float64 reward k/100, float32 store), y = prediction with y[i][action[i]] = target.  index_mode 0
 *                        reproduces the reference line for line: it takes a* as the POSITION inside the masked
 *                        sub-array (`np.argmax(qvalues[masks[i]])`) and uses that position as the action index;
 *                        index_mode 1 maps the position back to the action.  counters [2] u64 or NULL: invalid_count
 *                        (:367), empty masks (the reference raises there; index 0 is used).
 * All arrays [n][5] float32 row-major; action int64 (the dtype torch indexes with). */
int mdpp_replay_insert(int64_t n, const uint32_t *state_idx /*DEV*/, const uint8_t *action /*DEV*/,
                       const uint32_t *next_idx /*DEV*/, const float *reward /*DEV*/, const uint8_t *next_mask /*DEV*/,
                       const uint8_t *status /*DEV or NULL*/, const uint8_t *features /*DEV [n][80] or NULL*/,
                       const uint8_t *next_features /*DEV [n][80] or NULL*/, int64_t n_keys, uint32_t *present /*DEV*/,
                       uint32_t *m_next /*DEV*/, float *m_reward /*DEV*/, uint8_t *m_mask /*DEV*/,
                       uint8_t *m_features /*DEV or NULL*/, uint64_t *counters /*DEV [3]*/, void *stream);
int mdpp_replay_compact(int64_t n_keys, const uint32_t *present /*DEV*/, uint32_t *list /*DEV [n_keys]*/,
                        uint32_t *count /*DEV*/, void *stream);
int mdpp_dqn_target(int64_t n, const float *q_online_next /*DEV*/, const float *q_target_next /*DEV*/,
                    const uint8_t *next_mask /*DEV*/, const float *reward /*DEV*/, const int64_t *action /*DEV*/,
                    const float *prediction /*DEV*/, double discount, int index_mode, float *y /*DEV*/,
                    float *target_out /*DEV or NULL*/, uint64_t *counters /*DEV [2] or NULL*/, void *stream);

/* ---- deadlock precompute (-dd): backward reachability bitmap over the joint state space -------------
 * NEW definition next to the reference's stall heuristic: joint state = per car
 * (node, destination_index) or FINISHED, index = sum_c v_c * V^c with v_c = node_local*2 + di and
 * V-1 = finished, V = 2*n_local_nodes + 1.  Bit set = "all cars finished" is reachable when any
 * unfinished car may take any action its legal filter (encodestate(., 0) view) allows.  Synchronous:
 * sweeps until a fixpoint; *sweeps_out (host) receives the sweep count. */
int64_t mdpp_joint_state_count(const mdpp_config *cfg);
int mdpp_deadlock_reachability(mdpp_handle *h, uint32_t *reach_bitmap /*DEV*/, int32_t *sweeps_out, void *stream);
/* The same bitmap, found backwards: a level-synchronous breadth-first search from "all cars finished" over reverse
 * moves, each candidate predecessor checked with the legal filter, in ONE cooperative launch with a device-side
 * end test (csrc/reach.cuh).  Every live state is expanded once, so the whole search costs about one sweep of the
 * function above.  For joint spaces below 2^32 states (mdpp_deadlock_scratch_bytes returns 0 otherwise); scratch:
 * caller-owned device memory of mdpp_deadlock_scratch_bytes(cfg) bytes, 16-byte aligned.  Synchronous;
 * *levels_out (host) receives the number of search levels. */
/* One forward sweep over the joint states [first_state, first_state + n_range) only (first_state a multiple of 32):
 * bits of other states are read, never written; *changed (device word, set to 1 when a bit was set, never cleared
 * here).  The building block of the search sharded over GPUs: every rank sweeps its own word range of a
 * replicated bitmap and the ranks exchange their ranges until no sweep changes anything.  Asynchronous. */
int mdpp_deadlock_sweep_range(mdpp_handle *h, uint32_t *reach_bitmap /*DEV*/, int64_t first_state, int64_t n_range,
                              uint32_t *changed /*DEV*/, void *stream);
int64_t mdpp_deadlock_scratch_bytes(const mdpp_config *cfg);
int mdpp_deadlock_reachability_bfs(mdpp_handle *h, uint32_t *reach_bitmap /*DEV*/, void *scratch /*DEV*/,
                                   int64_t scratch_bytes, int32_t *levels_out, void *stream);

/* ---- state-id codec (for the qvalue .txt format, reference src/algorithms/qlearning.py:159-185) */
/* decode: state id -> primary node, destination node, others as (src box18, dst box18) pairs in the
 * reference's ascend_array order; returns the number of others or a negative error */
int mdpp_decode_state(const mdpp_config *cfg, int64_t state_idx, int32_t *pnode, int32_t *dnode,
                      int32_t *others /*[(MDPP_MAX_CARS-1)*2]*/);
int64_t mdpp_encode_state(const mdpp_config *cfg, int32_t pnode, int32_t dnode, int32_t n_others,
                          const int32_t *others);

/* ---- table image of the table-driven kernels (csrc/q_round.cuh FastTabs), built on the host: for tests and
 * inspection without a device.  out_host == NULL returns the image size; otherwise fills it and returns the size
 * (0 if a BFS length does not fit the move table: the rollout then uses its arithmetic kernel), negative on error.
 * Layout: own[5][256] x 16 B, oth[5][256] x 8 B, next[5][256][8] x 2 B, nth[32] x 4 B, ghost[5] x 8 B. */
int64_t mdpp_fast_tables(const mdpp_config *cfg, void *out_host, int64_t bytes);

/* ---- stats ---------------------------------------------------------------------------------- */
int mdpp_get_stats_host(mdpp_handle *h, mdpp_stats *out_host, int reset, void *stream);

#ifdef __cplusplus
}
#endif
#endif
