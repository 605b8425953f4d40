/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle") of the MARL-DPP hot path.
 *
 * This is the checker, never the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  It is a scalar, one-env-at-
 * a-time restatement of the reference's env + tabular Q-learner, written to follow the
 * reference's control flow (cited file:line, paths relative to the reference root), and is
 * PINNED against golden vectors recorded from the live reference (tests/golden/, generated
 * by oracle/make_golden.py).  Parity status: pinned (tables, trajectories, Q runs).
 *
 * Node ids: floor*36 + (box-1)*4 + heading, floor U=0 D=1, box 1..9, heading N,E,S,W = 0..3
 * (reference environment_node_name_list, src/layout.py:10).  ORC_NONE = the 'UYY'/'DYY' marker.
 */
#ifndef MDPP_ORACLE_H
#define MDPP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NODES 72
#define ORC_BOXES 18
#define ORC_NONE 255
#define ORC_MAX_CARS 7   /* reference cars_info has 7 rows, src/layout.py:27-34 */
#define ORC_ACT 6        /* S L R F B T, reference src/env_gym.py:6-21 */

typedef struct {
    int32_t rows, cols;                /* must be 3,3 */
    int32_t n_no_box, no_box[ORC_BOXES];          /* box18 ids */
    int32_t n_disc, disc_nodes[ORC_NODES];        /* node ids */
    int32_t n_station, station_box[8], station_entrance[8]; /* entrance: bit h set = heading h listed */
    int32_t n_dest, dest_nodes[2][8];             /* station_destination_list, 2 lists of n_dest */
    int32_t tcdi;                                 /* training_complete_destination_index */
    int32_t n_sel, sel_key[4], sel_n[4], sel_boxes[4][ORC_BOXES]; /* selective_blocked_boxes_dictionary */
} orc_layout;

/* canonical env state as the reference's state string carries it */
typedef struct {
    uint8_t pnode, dnode;              /* primary car: source node, destination node */
    uint8_t n_others;
    uint8_t others[ORC_MAX_CARS - 1][2]; /* (src box18, dst box18), reference-sorted */
} orc_state;

typedef struct orc_cfg orc_cfg;
typedef struct orc_env orc_env;
typedef struct orc_q orc_q;

/* ---- static graph (reference src/configuration.py:44-266) ---- */
orc_cfg *orc_cfg_create(const orc_layout *lay);
void orc_cfg_destroy(orc_cfg *c);
int orc_cfg_successor(const orc_cfg *c, int node, int action); /* ORC_NONE if 'xYY' */
int orc_cfg_in_no_node_list(const orc_cfg *c, int node);
int orc_cfg_base_legal(const orc_cfg *c, int node);             /* 6-bit mask, bit a = action a */
int orc_cfg_bfs_len(const orc_cfg *c, int src, int dst);        /* len(actions) of breadth_first_search */

/* ---- environment (reference src/env_gym.py) ---- */
orc_env *orc_env_create(const orc_cfg *c);
void orc_env_destroy(orc_env *e);
/* text = contents of the bad-states file, or NULL => give_bad_state_reward False (env_gym.py:71-83) */
void orc_env_set_bad_states(orc_env *e, const char *text);
void orc_env_reset(orc_env *e);
void orc_env_initialize_cars(orc_env *e, int n);
/* start_nodes already carry the heading the reference draws with random.shuffle (env_gym.py:345-349) */
void orc_env_initialize(orc_env *e, const int32_t *dest_index, const int32_t *start_nodes, const int32_t *choose3);
void orc_env_encodestate(orc_env *e, int car, int deadlock_id, orc_state *out);
int orc_env_legal(const orc_env *e, const orc_state *s);        /* filtered, 6-bit mask */
void orc_env_successor_state(const orc_env *e, const orc_state *s, int action, orc_state *out);
double orc_env_reward(const orc_env *e, const orc_state *prev, const orc_state *next, int car, int dqn);
void orc_env_update_car(orc_env *e, int car, int node);
int orc_env_is_game_ended(const orc_env *e);
int orc_env_car_done(const orc_env *e, int car);
int orc_env_car_node(const orc_env *e, int car);
int orc_env_car_dest_index(const orc_env *e, int car);
int orc_env_car_done_count(const orc_env *e, int car);
void orc_env_set_car(orc_env *e, int car, int node, int dest_index, int done, int done_count);
int orc_state_to_string(const orc_state *s, char *buf, int cap);
int orc_state_from_string(const char *str, orc_state *out);

/* ---- tabular Q-learner (reference src/algorithms/qlearning.py:34-116,600-667) ---- */
orc_q *orc_q_create(int f32);          /* f32=0: float64 like the reference; 1: same ops rounded to float */
void orc_q_destroy(orc_q *q);
void orc_q_set_params(orc_q *q, double alpha, double discount);
double orc_q_get(orc_q *q, const orc_state *s, int action);       /* inserts 0.0 like Counter.__getitem__ */
double orc_q_peek(const orc_q *q, const orc_state *s, int action, int *present);
void orc_q_set(orc_q *q, const orc_state *s, int action, double v);
int orc_q_greedy(orc_q *q, const orc_env *e, const orc_state *s); /* computeactionfromqvalues; -1 if none */
double orc_q_value(orc_q *q, const orc_env *e, const orc_state *s); /* computevaluefromqvalues */
void orc_q_update(orc_q *q, const orc_env *e, const orc_state *s, int a, const orc_state *ns, double r);
int64_t orc_q_size(const orc_q *q);
void orc_q_entry(const orc_q *q, int64_t i, orc_state *s, int *action, double *v); /* insertion order */

/* DQN featurizer (reference src/algorithms/dqn_open_source.py:245-313 RL.convert) */
void orc_convert(const orc_state *s, int total_cars, uint8_t *out);

/* epsilon schedule (reference src/utils/utils.py:12-52) */
double orc_epsilon(int episodes, int episode, int model);

/* Philox4x32 (Salmon et al., SC'11; Random123 reference constants); the batch semantics use MDPP_PHILOX_ROUNDS = 7 rounds */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4], int rounds);
void orc_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]); /* MDPP_PHILOX_ROUNDS rounds */
int orc_philox_rounds(void);

#ifdef __cplusplus
}
#endif
#endif
