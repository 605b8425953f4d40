/*
 * TEST INFRASTRUCTURE ONLY -- batched driver on top of the scalar oracle (mdpp_oracle.c).
 *
 * Restates, on the CPU and one env at a time, the BATCH semantics the CUDA path defines
 * (DESIGN.md "Batch semantics"): N independent envs stepped in lockstep, car slot by car slot,
 * decisions drawn from Philox4x32-7 (MDPP_PHILOX_ROUNDS) keyed (seed, env id) with counter (round, stream), per-env
 * episode counters driving the reference's epsilon schedule, auto-reset at the end of a round,
 * and -- for the learner -- every env of a sub-step reading the Q table as it was when the
 * sub-step began, with each touched key receiving the value its contributing envs computed.  All
 * contributors of one key compute the SAME value: (state, action) fixes the successor state string,
 * the reward (the clash term is 0 under legal actions: a listed car blocks its box number, a finished
 * car is parked on a no-box node) and both Q rows are read from the same snapshot.  The batch
 * definition is therefore "the value" -- formally the maximum, so that it stays order-independent
 * should the claim ever fail; orc_batch_disagreements() counts keys whose contributors differed and
 * the tests assert it stays 0.
 * The per-env pieces are the reference-pinned scalar functions; nothing here is shared with the
 * product code.  pthreads spread envs over host cores for the CPU baseline in bench.py.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "mdpp_oracle.h"

/* from mdpp_oracle.c */
void orc_env_share_bad_states(orc_env *e, const orc_env *owner);
int64_t orc_q_index(orc_q *q, const orc_state *s, int action);
double orc_q_at(const orc_q *q, int64_t i);
void orc_q_set_at(orc_q *q, int64_t i, double v);

#define ACT_SKIP 0xFF
#define ACT_GREEDY 0xFE
static const int ACT5[5] = {0, 1, 2, 3, 5}; /* S L R F T in the oracle's 6-action numbering */

typedef struct {
    const orc_cfg *cfg;
    int64_t n_envs;
    int n_cars;
    uint32_t env_id_base;
    int step_cap;
    int32_t start_box[ORC_MAX_CARS], dest_index[ORC_MAX_CARS], choose[3 * ORC_MAX_CARS];
    orc_env **envs;
    int32_t *episode, *steps;
    int32_t *stall;      /* consecutive greedy 'S' at epsilon 0 (detect_deadlocks_v0, qlearning.py:347-360) */
    orc_state *dl_state; /* state flagged by the stall heuristic; pnode == ORC_NONE: none */
    int32_t *dl_episode; /* episode in which it was flagged (the reference returns it, qlearning.py:379-380); -1: none */
    /* stats */
    uint64_t agent_steps, episodes_done, episodes_capped, explore_steps;
    int64_t reward_sum;
    uint64_t disagreements; /* (sub-step, key) groups whose contributors computed different values */
    uint64_t touched_keys;  /* distinct keys written, summed over sub-steps */
} orc_batch;

static void init_env(orc_batch *b, int64_t e, uint32_t heading_bits) {
    int32_t nodes[ORC_MAX_CARS];
    for (int i = 0; i < b->n_cars; i++) nodes[i] = b->start_box[i] * 4 + (int)((heading_bits >> (2 * i)) & 3u);
    orc_env_reset(b->envs[e]);
    orc_env_initialize_cars(b->envs[e], b->n_cars);
    orc_env_initialize(b->envs[e], b->dest_index, nodes, b->choose);
}

orc_batch *orc_batch_create(const orc_cfg *cfg, int64_t n_envs, int n_cars, const int32_t *start_box,
                            const int32_t *dest_index, const int32_t *choose3, const char *bad_text, int step_cap,
                            uint32_t env_id_base) {
    orc_batch *b = (orc_batch *)calloc(1, sizeof(orc_batch));
    b->cfg = cfg;
    b->n_envs = n_envs;
    b->n_cars = n_cars;
    b->env_id_base = env_id_base;
    b->step_cap = step_cap;
    memcpy(b->start_box, start_box, sizeof(int32_t) * (size_t)n_cars);
    memcpy(b->dest_index, dest_index, sizeof(int32_t) * (size_t)n_cars);
    memcpy(b->choose, choose3, sizeof(int32_t) * 3 * (size_t)n_cars);
    b->envs = (orc_env **)calloc((size_t)n_envs, sizeof(orc_env *));
    b->episode = (int32_t *)calloc((size_t)n_envs, sizeof(int32_t));
    b->steps = (int32_t *)calloc((size_t)n_envs, sizeof(int32_t));
    b->stall = (int32_t *)calloc((size_t)n_envs, sizeof(int32_t));
    b->dl_state = (orc_state *)calloc((size_t)n_envs, sizeof(orc_state));
    b->dl_episode = (int32_t *)calloc((size_t)n_envs, sizeof(int32_t));
    for (int64_t e = 0; e < n_envs; e++) b->dl_state[e].pnode = ORC_NONE;
    for (int64_t e = 0; e < n_envs; e++) b->dl_episode[e] = -1;
    for (int64_t e = 0; e < n_envs; e++) {
        b->envs[e] = orc_env_create(cfg);
        if (e == 0) orc_env_set_bad_states(b->envs[0], bad_text); /* parsed once, read by every env of the batch */
        else orc_env_share_bad_states(b->envs[e], b->envs[0]);
    }
    return b;
}
void orc_batch_destroy(orc_batch *b) {
    for (int64_t e = b->n_envs - 1; e >= 0; e--) orc_env_destroy(b->envs[e]); /* env 0 owns the shared set: last */
    free(b->envs);
    free(b->episode);
    free(b->steps);
    free(b->stall);
    free(b->dl_state);
    free(b->dl_episode);
    free(b);
}
orc_env *orc_batch_env(orc_batch *b, int64_t e) { return b->envs[e]; }
int32_t orc_batch_episode(const orc_batch *b, int64_t e) { return b->episode[e]; }
int32_t orc_batch_steps(const orc_batch *b, int64_t e) { return b->steps[e]; }
const orc_state *orc_batch_deadlock_state(const orc_batch *b, int64_t e) { return &b->dl_state[e]; }
int32_t orc_batch_deadlock_episode(const orc_batch *b, int64_t e) { return b->dl_episode[e]; }
uint64_t orc_batch_disagreements(const orc_batch *b) { return b->disagreements; }
uint64_t orc_batch_touched_keys(const orc_batch *b) { return b->touched_keys; }
/* bulk view of every env for the full-size parity tests: per car (node, destination_index, done, done_count clamped
 * to -1), then the episode and step counters */
void orc_batch_export(const orc_batch *b, int16_t *cars4 /* [n_envs][n_cars][4] */, int32_t *episode, int32_t *steps) {
    for (int64_t e = 0; e < b->n_envs; e++) {
        for (int c = 0; c < b->n_cars; c++) {
            int16_t *o = cars4 + ((size_t)e * (size_t)b->n_cars + (size_t)c) * 4;
            int cnt = orc_env_car_done_count(b->envs[e], c);
            o[0] = (int16_t)orc_env_car_node(b->envs[e], c);
            o[1] = (int16_t)orc_env_car_dest_index(b->envs[e], c);
            o[2] = (int16_t)orc_env_car_done(b->envs[e], c);
            o[3] = (int16_t)(cnt < -1 ? -1 : cnt);
        }
        if (episode) episode[e] = b->episode[e];
        if (steps) steps[e] = b->steps[e];
    }
}
void orc_batch_stats(const orc_batch *b, uint64_t *out5) {
    out5[0] = b->agent_steps;
    out5[1] = b->episodes_done;
    out5[2] = b->episodes_capped;
    out5[3] = b->explore_steps;
    out5[4] = (uint64_t)b->reward_sum;
}

static uint32_t philox_word(uint64_t round, uint32_t stream, uint32_t seed, uint32_t env, int word) {
    uint32_t ctr[4] = {(uint32_t)round, (uint32_t)(round >> 32), stream, 0}, key[2] = {seed, env}, out[4];
    orc_philox4x32_r(ctr, key, out);
    return out[word];
}
/* decision word of car c in round r: cars 0..3 share one Philox block, car 4 uses the next stream */
static uint32_t decision_word(uint64_t round, int car, uint32_t seed, uint32_t env) {
    return philox_word(round, (uint32_t)(car >> 2), seed, env, car & 3);
}

/* headings: [n_envs][n_cars] values 0..3 or NULL (Philox stream 3, round 0) */
void orc_batch_reset(orc_batch *b, const uint8_t *headings, uint32_t seed, int zero_counters) {
    for (int64_t e = 0; e < b->n_envs; e++) {
        uint32_t hb = 0;
        if (headings)
            for (int i = 0; i < b->n_cars; i++) hb |= (uint32_t)(headings[e * b->n_cars + i] & 3u) << (2 * i);
        else
            hb = philox_word(0, 3, seed, b->env_id_base + (uint32_t)e, 0);
        init_env(b, e, hb);
        b->steps[e] = 0;
        b->stall[e] = 0;
        if (zero_counters) {
            b->episode[e] = 0;
            b->dl_state[e].pnode = ORC_NONE;
            b->dl_episode[e] = -1;
        }
    }
}

static int nth_set_bit(int mask, int n) {
    for (int a = 0; a < 8; a++)
        if ((mask >> a) & 1) {
            if (n == 0) return a;
            n--;
        }
    return -1;
}
/* 6-action oracle mask -> 5-bit S,L,R,F,T mask of the ABI */
static int mask5(int m6) { return (m6 & 0xF) | (((m6 >> 5) & 1) << 4); }

static void end_of_episode_check(orc_batch *b, int64_t e, uint32_t heading_word, uint64_t *episodes, uint64_t *capped) {
    int finished = orc_env_is_game_ended(b->envs[e]);
    int cap = !finished && b->step_cap > 0 && b->steps[e] >= b->step_cap;
    if (finished || cap) {
        (*episodes)++;
        *capped += (uint64_t)cap;
        b->episode[e] += 1;
        b->steps[e] = 0;
        b->stall[e] = 0; /* deadlock_count = 0 per episode, qlearning.py:331 */
        init_env(b, e, heading_word);
    }
}
static void end_of_round(orc_batch *b, int64_t e, uint64_t round, uint32_t seed, uint64_t *episodes, uint64_t *capped) {
    /* the heading word is only drawn when it is used; it is a pure function of (round, env) */
    end_of_episode_check(b, e, philox_word(round, 2, seed, b->env_id_base + (uint32_t)e, 0), episodes, capped);
}

/* ---- random-policy rollout (the reference's epsilon = 1 behaviour, qlearning.py:99-102) -------- */
typedef struct {
    orc_batch *b;
    int64_t lo, hi;
    int rounds, dqn;
    uint64_t first_round;
    uint32_t seed;
    uint8_t *action, *done;
    float *reward;
    orc_state *next;
    uint64_t steps, episodes, capped;
    int64_t rsum;
} rollout_job;

static void *rollout_worker(void *arg) {
    rollout_job *j = (rollout_job *)arg;
    orc_batch *b = j->b;
    const int nc = b->n_cars;
    for (int64_t e = j->lo; e < j->hi; e++) {
        orc_env *env = b->envs[e];
        uint32_t envid = b->env_id_base + (uint32_t)e;
        for (int r = 0; r < j->rounds; r++) {
            uint64_t round = j->first_round + (uint64_t)r;
            for (int c = 0; c < nc; c++) {
                size_t o = ((size_t)r * (size_t)nc + (size_t)c) * (size_t)b->n_envs + (size_t)e;
                int a5 = ACT_SKIP;
                double rw = 0;
                orc_state ns;
                memset(&ns, 0, sizeof ns);
                ns.pnode = ORC_NONE;
                if (!orc_env_car_done(env, c)) {
                    orc_state st;
                    orc_env_encodestate(env, c, 1, &st);
                    int lm = mask5(orc_env_legal(env, &st));
                    if (lm) {
                        uint32_t w = decision_word(round, c, j->seed, envid);
                        int cnt = __builtin_popcount((unsigned)lm);
                        a5 = nth_set_bit(lm, (int)(((w & 0xFFFFu) * (uint32_t)cnt) >> 16));
                        orc_env_successor_state(env, &st, ACT5[a5], &ns);
                        rw = orc_env_reward(env, &st, &ns, c, j->dqn);
                        orc_env_update_car(env, c, ns.pnode);
                        if (b->steps[e] < 0xFFFF) b->steps[e]++;
                        j->steps++;
                        j->rsum += (int64_t)llround(j->dqn ? rw * 100.0 : rw);
                    }
                }
                if (j->action) j->action[o] = (uint8_t)a5;
                if (j->reward) j->reward[o] = (float)rw;
                if (j->next) j->next[o] = ns;
                if (j->done) j->done[o] = (uint8_t)orc_env_is_game_ended(env);
            }
            end_of_round(b, e, round, j->seed, &j->episodes, &j->capped);
        }
    }
    return NULL;
}

void orc_batch_rollout_random(orc_batch *b, int rounds, uint64_t first_round, uint32_t seed, int dqn, int threads,
                              uint8_t *action, float *reward, uint8_t *done, orc_state *next) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    rollout_job jobs[256];
    pthread_t tid[256];
    for (int t = 0; t < threads; t++) {
        rollout_job *j = &jobs[t];
        memset(j, 0, sizeof *j);
        j->b = b;
        j->lo = b->n_envs * t / threads;
        j->hi = b->n_envs * (t + 1) / threads;
        j->rounds = rounds;
        j->dqn = dqn;
        j->first_round = first_round;
        j->seed = seed;
        j->action = action;
        j->reward = reward;
        j->done = done;
        j->next = next;
        if (threads == 1) rollout_worker(j);
        else pthread_create(&tid[t], NULL, rollout_worker, j);
    }
    for (int t = 0; t < threads; t++) {
        if (threads > 1) pthread_join(tid[t], NULL);
        b->agent_steps += jobs[t].steps;
        b->explore_steps += jobs[t].steps;
        b->episodes_done += jobs[t].episodes;
        b->episodes_capped += jobs[t].capped;
        b->reward_sum += jobs[t].rsum;
    }
}

/* ---- learner sub-step ---------------------------------------------------------------------- */
typedef struct {
    float alpha, discount;
    int32_t episodes;
    int32_t eps_threshold[5];
    uint32_t eps_q16[6];
    uint32_t seed;
    uint32_t flags; /* 1 frozen (no update), 2 encodestate(.,0), 4 stall detection, 16 random car order */
} orc_learner;

static uint32_t eps_q16_of(const orc_learner *lp, int32_t episode) {
    uint32_t e = lp->eps_q16[5];
    for (int i = 4; i >= 0; i--)
        if (episode < lp->eps_threshold[i]) e = lp->eps_q16[i];
    return e;
}

typedef struct {
    int valid;       /* this env updates a key in this sub-step */
    int action;      /* oracle action number of the key */
    orc_state st;    /* its state */
    float newv;
} pending;

/* phase 1 of a sub-step over the envs [lo, hi): every env decides and steps against the Q table as it stands.
 * Reads the table only (orc_q_peek: no inserts), touches only its own envs and its own counters, so disjoint
 * ranges may run on different host threads (the BFS memo of the graph must have been filled before). */
typedef struct {
    orc_batch *b;
    const orc_q *q;
    const orc_learner *lp;
    int slot;
    uint64_t round;
    const uint8_t *forced;
    uint8_t *action_out;
    orc_state *state_out;
    float *reward_out;
    pending *pend;
    int64_t lo, hi;
    uint64_t agent_steps, explore_steps, episodes_done, episodes_capped;
    int64_t reward_sum;
} substep_job;

static void *substep_worker(void *arg) {
    substep_job *j = (substep_job *)arg;
    orc_batch *b = j->b;
    const orc_q *q = j->q;
    const orc_learner *lp = j->lp;
    const int slot = j->slot;
    const uint64_t round = j->round;
    const uint8_t *forced = j->forced;
    pending *pend = j->pend;
    const int nc = b->n_cars;
    const int any_order = (lp->flags & 16u) != 0;
    for (int64_t e = j->lo; e < j->hi; e++) {
        orc_env *env = b->envs[e];
        uint32_t envid = b->env_id_base + (uint32_t)e;
        pend[e].valid = 0;
        int a_out = ACT_SKIP;
        float r_out = 0.f;
        orc_state st;
        memset(&st, 0, sizeof st);
        st.pnode = ORC_NONE;
        int live = b->episode[e] < lp->episodes;
        int c = slot, f = forced ? forced[e] : ACT_SKIP;
        uint32_t blk[4] = {0, 0, 0, 0};
        if (any_order) {
            uint32_t ctr[4] = {(uint32_t)round, (uint32_t)(round >> 32), 4u + (uint32_t)slot, 0}, key[2] = {lp->seed, envid};
            orc_philox4x32_r(ctr, key, blk);
            c = (int)(((blk[0] >> 16) * (uint32_t)nc) >> 16);
            if ((f >> 4) != 0xF) c = f >> 4;
            f = (f & 0xF) == 0xF ? ACT_SKIP : (f & 0xF) == 0xE ? ACT_GREEDY : (f & 0xF);
            if (c >= nc) c = nc - 1;
        }
        if (live && !orc_env_car_done(env, c)) {
            orc_env_encodestate(env, c, (lp->flags & 2u) ? 0 : 1, &st);
            int lm = mask5(orc_env_legal(env, &st));
            int forced_ok = f >= ACT_GREEDY || (f < 5 && ((lm >> f) & 1));
            if (lm && forced_ok) {
                int a5, explore;
                if (f == ACT_GREEDY) {
                    a5 = 0;
                    explore = 0;
                } else if (f != ACT_SKIP) {
                    a5 = f;
                    explore = 1;
                } else {
                    uint32_t w = any_order ? blk[1] : decision_word(round, c, lp->seed, envid);
                    explore = (w >> 16) < eps_q16_of(lp, b->episode[e]);
                    a5 = nth_set_bit(lm, (int)(((w & 0xFFFFu) * (uint32_t)__builtin_popcount((unsigned)lm)) >> 16));
                }
                if (!explore) { /* first maximum in S,L,R,F,T order; absent keys read 0.0 */
                    float best = 0.f;
                    int have = 0;
                    for (int k = 0; k < 5; k++) {
                        if (!((lm >> k) & 1)) continue;
                        float v = (float)orc_q_peek(q, &st, ACT5[k], NULL);
                        if (!have || v > best) { best = v; a5 = k; have = 1; }
                    }
                }
                j->explore_steps += (uint64_t)explore;
                int stalled = 0;
                if (lp->flags & 4u) { /* qlearning.py:347-360 */
                    if (a5 == 0 && eps_q16_of(lp, b->episode[e]) == 0) b->stall[e] += 1;
                    else b->stall[e] = 0;
                    if (b->stall[e] == 3 * nc) {
                        stalled = 1;
                        b->stall[e] = 0;
                        b->dl_state[e] = st;
                        b->dl_episode[e] = b->episode[e];
                        b->episode[e] = lp->episodes;
                    }
                }
                a_out = a5;
                if (stalled) goto after_step;
                orc_state ns;
                orc_env_successor_state(env, &st, ACT5[a5], &ns);
                double rw = orc_env_reward(env, &st, &ns, c, 0);
                int lm2 = mask5(orc_env_legal(env, &ns));
                float nv = 0.f;
                int have = 0;
                for (int k = 0; k < 5; k++) {
                    if (!((lm2 >> k) & 1)) continue;
                    float v = (float)orc_q_peek(q, &ns, ACT5[k], NULL);
                    if (!have || v > nv) { nv = v; have = 1; }
                }
                float qa = (float)orc_q_peek(q, &st, ACT5[a5], NULL);
                float target = (float)rw + lp->discount * nv;
                float diff = target - qa;
                float newv = qa + lp->alpha * diff;
                if (!(lp->flags & 1u)) {
                    pend[e].valid = 1;
                    pend[e].action = ACT5[a5];
                    pend[e].st = st;
                    pend[e].newv = newv;
                }
                orc_env_update_car(env, c, ns.pnode);
                if (b->steps[e] < 0xFFFF) b->steps[e]++;
                j->agent_steps++;
                j->reward_sum += (int64_t)rw;
                r_out = (float)rw;
            after_step:;
            }
        }
        live = b->episode[e] < lp->episodes || live;
        if (live && any_order) end_of_episode_check(b, e, blk[2], &j->episodes_done, &j->episodes_capped);
        else if (live && c == nc - 1) end_of_round(b, e, round, lp->seed, &j->episodes_done, &j->episodes_capped);
        if (j->action_out) j->action_out[e] = (uint8_t)a_out;
        if (j->state_out) j->state_out[e] = st;
        if (j->reward_out) j->reward_out[e] = r_out;
    }
    return NULL;
}

/* host threads of phase 1 (full-size parity tests: 2^20 envs); 1 = everything on the calling thread */
static int g_substep_threads = 1;
void orc_batch_set_threads(int n) { g_substep_threads = n < 1 ? 1 : n > 256 ? 256 : n; }

/* One sub-step for car slot c.  q must have been created with f32 = 1.  forced: NULL or [n_envs]
 * (ACT_SKIP = own decision, else take that action as exploration).  Outputs optional.
 *
 * flags & 16: random car order (train_given_state_random_order, qlearning.py:241-295).  `slot` is then only the
 * index of the sub-step within the round (a round is n_cars sub-steps): every env draws its own car --
 * random.randint(0, n - 1) (:267) = the high half of word 0 of the Philox block (round, stream 4 + slot) scaled to
 * n_cars; a finished car means no agent-step (:268-269); the decision word is word 1 of that block; the end of the
 * episode is checked after every agent-step (:291), the headings of the next one come from word 2.  forced then
 * carries the car in its high nibble (0xF = the env's own draw) and the action in the low one (0..4, 0xE greedy,
 * 0xF own decision), so 0xFF / 0xFE keep their meaning. */
void orc_batch_q_substep(orc_batch *b, orc_q *q, const orc_learner *lp, int slot, uint64_t round, const uint8_t *forced,
                         uint8_t *action_out, orc_state *state_out, float *reward_out) {
    pending *pend = (pending *)malloc(sizeof(pending) * (size_t)b->n_envs);
    /* phase 1: every env decides and steps against the Q table as it stands (no writes yet) */
    int T = g_substep_threads;
    if (b->n_envs < 4096) T = 1;
    if (T > 1) /* fill the BFS memo up front: the workers only read it */
        for (int s = 0; s < ORC_NODES; s++)
            for (int d = 0; d < ORC_NODES; d++) orc_cfg_bfs_len(b->cfg, s, d);
    substep_job jobs[256];
    pthread_t tid[256];
    for (int t = 0; t < T; t++) {
        substep_job *j = &jobs[t];
        memset(j, 0, sizeof *j);
        j->b = b; j->q = q; j->lp = lp; j->slot = slot; j->round = round; j->forced = forced;
        j->action_out = action_out; j->state_out = state_out; j->reward_out = reward_out; j->pend = pend;
        j->lo = b->n_envs * t / T;
        j->hi = b->n_envs * (t + 1) / T;
        if (T == 1) substep_worker(j);
        else pthread_create(&tid[t], NULL, substep_worker, j);
    }
    for (int t = 0; t < T; t++) {
        if (T > 1) pthread_join(tid[t], NULL);
        b->agent_steps += jobs[t].agent_steps;
        b->explore_steps += jobs[t].explore_steps;
        b->episodes_done += jobs[t].episodes_done;
        b->episodes_capped += jobs[t].episodes_capped;
        b->reward_sum += jobs[t].reward_sum;
    }
    /* phase 2: per key, the value the contributing envs computed (identical for all of them;
     * formally the maximum).  Keys enter the map in env order, as if the envs had run one after another. */
    int64_t *keyidx = (int64_t *)malloc(sizeof(int64_t) * (size_t)b->n_envs);
    for (int64_t e = 0; e < b->n_envs; e++)
        keyidx[e] = pend[e].valid ? orc_q_index(q, &pend[e].st, pend[e].action) : -1;
    int64_t nq = orc_q_size(q);
    uint32_t *cnt = (uint32_t *)calloc((size_t)nq + 1, sizeof(uint32_t));
    float *val = (float *)calloc((size_t)nq + 1, sizeof(float));
    uint8_t *differ = (uint8_t *)calloc((size_t)nq + 1, 1);
    for (int64_t e = 0; e < b->n_envs; e++) {
        if (keyidx[e] < 0) continue;
        const int64_t k = keyidx[e];
        if (cnt[k] == 0) val[k] = pend[e].newv;
        else if (pend[e].newv != val[k]) {
            differ[k] = 1;
            if (pend[e].newv > val[k]) val[k] = pend[e].newv;
        }
        cnt[k] += 1;
    }
    for (int64_t k = 0; k < nq; k++) {
        if (!cnt[k]) continue;
        b->touched_keys++;
        b->disagreements += differ[k];
        orc_q_set_at(q, k, (double)val[k]);
    }
    free(cnt);
    free(val);
    free(differ);
    free(keyidx);
    free(pend);
}

void orc_batch_q_round(orc_batch *b, orc_q *q, const orc_learner *lp, uint64_t round) {
    for (int c = 0; c < b->n_cars; c++) orc_batch_q_substep(b, q, lp, c, round, NULL, NULL, NULL, NULL);
}

/* ---- CPU baseline: independent learner replicas, one per host thread --------------------------
 * The reference is single-threaded with a module-global env (qlearning.py:17-18); the way to use
 * every host core is one independent learner per core.  Each thread owns a
 * batch of `envs_per_thread` envs and its own Q table and runs `rounds` full rounds. */
typedef struct {
    orc_batch *b;
    orc_q *q;
    const orc_learner *lp;
    uint64_t first_round;
    int rounds;
} bench_job;

static void *bench_worker(void *arg) {
    bench_job *j = (bench_job *)arg;
    for (int r = 0; r < j->rounds; r++) orc_batch_q_round(j->b, j->q, j->lp, j->first_round + (uint64_t)r);
    return NULL;
}

typedef struct {
    int threads;
    orc_batch **b;
    orc_q **q;
    orc_learner lp;
    uint64_t round;
} orc_bench;

orc_bench *orc_bench_create(const orc_cfg *cfg, int threads, int64_t envs_per_thread, int n_cars,
                            const int32_t *start_box, const int32_t *dest_index, const int32_t *choose3,
                            const char *bad_text, int step_cap, const orc_learner *lp) {
    for (int s = 0; s < ORC_NODES; s++) /* fill the BFS memo up front: workers only read it */
        for (int d = 0; d < ORC_NODES; d++) orc_cfg_bfs_len(cfg, s, d);
    orc_bench *o = (orc_bench *)calloc(1, sizeof(orc_bench));
    o->threads = threads;
    o->b = (orc_batch **)calloc((size_t)threads, sizeof(orc_batch *));
    o->q = (orc_q **)calloc((size_t)threads, sizeof(orc_q *));
    o->lp = *lp;
    for (int t = 0; t < threads; t++) {
        o->b[t] = orc_batch_create(cfg, envs_per_thread, n_cars, start_box, dest_index, choose3, bad_text, step_cap,
                                   (uint32_t)((int64_t)t * envs_per_thread));
        orc_batch_reset(o->b[t], NULL, lp->seed, 1);
        o->q[t] = orc_q_create(1);
        orc_q_set_params(o->q[t], lp->alpha, lp->discount);
    }
    return o;
}
void orc_bench_destroy(orc_bench *o) {
    for (int t = 0; t < o->threads; t++) {
        orc_batch_destroy(o->b[t]);
        orc_q_destroy(o->q[t]);
    }
    free(o->b);
    free(o->q);
    free(o);
}
/* runs `rounds` rounds on every replica in parallel; returns the agent-steps taken */
uint64_t orc_bench_run(orc_bench *o, int rounds) {
    bench_job jobs[256];
    pthread_t tid[256];
    uint64_t before = 0, after = 0;
    int T = o->threads > 256 ? 256 : o->threads;
    for (int t = 0; t < T; t++) before += o->b[t]->agent_steps;
    for (int t = 0; t < T; t++) {
        jobs[t].b = o->b[t];
        jobs[t].q = o->q[t];
        jobs[t].lp = &o->lp;
        jobs[t].first_round = o->round;
        jobs[t].rounds = rounds;
        pthread_create(&tid[t], NULL, bench_worker, &jobs[t]);
    }
    for (int t = 0; t < T; t++) pthread_join(tid[t], NULL);
    o->round += (uint64_t)rounds;
    for (int t = 0; t < T; t++) after += o->b[t]->agent_steps;
    return after - before;
}

/* ---- deadlock reachability: CPU brute force over the reference-pinned transition functions --------
 * Same NEW definition as the CUDA bitmap (include/mdpp.h): joint state = per car (node, dest index)
 * or FINISHED; any unfinished car may take any action of getlegalactions on its encodestate(., 0)
 * state; live iff all-finished is reachable.  Joint index = sum_c v_c * V^c, v_c = node_local*2 + di,
 * V - 1 = finished.  Forward successors are enumerated once through the scalar env functions, then a
 * backward closure runs over the reverse adjacency.  Returns the number of live states. */
typedef struct {
    const orc_cfg *cfg;
    int n_cars, n_local_nodes, node_base;
    const int32_t *choose3;
    uint64_t lo, hi, V;
    const uint64_t *pw;
    uint32_t *succ;
    uint8_t *nsucc;
} reach_job;

/* forward successors of the joint states [lo, hi), through the scalar env functions */
static void *reach_worker(void *arg) {
    reach_job *j = (reach_job *)arg;
    const int n_cars = j->n_cars, node_base = j->node_base, maxs = n_cars * 5;
    const uint64_t V = j->V;
    orc_env *env = orc_env_create(j->cfg);
    int32_t di0[ORC_MAX_CARS] = {0}, nodes0[ORC_MAX_CARS] = {0};
    for (uint64_t idx = j->lo; idx < j->hi; idx++) {
        uint64_t rest = idx;
        int digit[ORC_MAX_CARS];
        orc_env_reset(env);
        orc_env_initialize_cars(env, n_cars);
        orc_env_initialize(env, di0, nodes0, j->choose3);
        for (int c = 0; c < n_cars; c++) {
            digit[c] = (int)(rest % V);
            rest /= V;
            if (digit[c] == (int)V - 1) orc_env_set_car(env, c, 0, 0, 1, -1); /* finished: never listed in mode 0 */
            else orc_env_set_car(env, c, digit[c] / 2 + node_base, digit[c] & 1, 0, 5);
        }
        for (int c = 0; c < n_cars; c++) {
            if (digit[c] == (int)V - 1) continue;
            orc_state st, ns;
            orc_env_encodestate(env, c, 0, &st);
            int legal = orc_env_legal(env, &st);
            for (int a = 0; a < ORC_ACT; a++) {
                if (!((legal >> a) & 1)) continue;
                orc_env_successor_state(env, &st, a, &ns);
                orc_env_update_car(env, c, ns.pnode);
                int nd = orc_env_car_done(env, c) ? (int)V - 1
                                                  : (orc_env_car_node(env, c) - node_base) * 2 + orc_env_car_dest_index(env, c);
                j->succ[idx * (uint64_t)maxs + j->nsucc[idx]++] =
                    (uint32_t)(idx - (uint64_t)digit[c] * j->pw[c] + (uint64_t)nd * j->pw[c]);
                orc_env_set_car(env, c, digit[c] / 2 + node_base, digit[c] & 1, 0, 5); /* undo */
            }
        }
    }
    orc_env_destroy(env);
    return NULL;
}

int64_t orc_reachability(const orc_cfg *cfg, int n_cars, const int32_t *choose3, int n_local_nodes, int node_base,
                         uint32_t *bitmap_out) {
    const uint64_t V = 2ull * (uint64_t)n_local_nodes + 1ull;
    uint64_t n = 1, pw[ORC_MAX_CARS];
    for (int c = 0; c < n_cars; c++) { pw[c] = n; n *= V; }
    const int maxs = n_cars * 5;
    uint32_t *succ = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n * (size_t)maxs);
    uint8_t *nsucc = (uint8_t *)calloc((size_t)n, 1);
    {
        int T = n < 100000 ? 1 : g_substep_threads;
        reach_job jobs[256];
        pthread_t tid[256];
        for (int t = 0; t < T; t++) {
            reach_job *j = &jobs[t];
            j->cfg = cfg; j->n_cars = n_cars; j->n_local_nodes = n_local_nodes; j->node_base = node_base;
            j->choose3 = choose3; j->V = V; j->pw = pw; j->succ = succ; j->nsucc = nsucc;
            j->lo = n * (uint64_t)t / (uint64_t)T;
            j->hi = n * (uint64_t)(t + 1) / (uint64_t)T;
            if (T == 1) reach_worker(j);
            else pthread_create(&tid[t], NULL, reach_worker, j);
        }
        if (T > 1)
            for (int t = 0; t < T; t++) pthread_join(tid[t], NULL);
    }
    /* reverse adjacency (CSR) + backward BFS from the all-finished state */
    uint64_t *start = (uint64_t *)calloc((size_t)n + 2, sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++)
        for (int k = 0; k < nsucc[i]; k++) start[succ[i * (uint64_t)maxs + k] + 2]++;
    for (uint64_t i = 2; i < n + 2; i++) start[i] += start[i - 1];
    uint32_t *pred = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(start[n + 1] + 1));
    for (uint64_t i = 0; i < n; i++)
        for (int k = 0; k < nsucc[i]; k++) pred[start[succ[i * (uint64_t)maxs + k] + 1]++] = (uint32_t)i;
    uint8_t *live = (uint8_t *)calloc((size_t)n, 1);
    uint32_t *queue = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
    uint64_t head = 0, tail = 0;
    live[n - 1] = 1;
    queue[tail++] = (uint32_t)(n - 1);
    while (head < tail) {
        uint32_t u = queue[head++];
        for (uint64_t k = start[u]; k < start[u + 1]; k++) {
            uint32_t p = pred[k];
            if (!live[p]) { live[p] = 1; queue[tail++] = p; }
        }
    }
    memset(bitmap_out, 0, (((size_t)n + 31) / 32) * 4);
    for (uint64_t i = 0; i < n; i++)
        if (live[i]) bitmap_out[i >> 5] |= 1u << (i & 31);
    free(succ); free(nsucc); free(start); free(pred); free(live); free(queue);
    return (int64_t)tail;
}
