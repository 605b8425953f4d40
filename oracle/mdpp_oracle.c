/*
 * TEST INFRASTRUCTURE ONLY -- see mdpp_oracle.h.  Scalar CPU restatement of the reference's
 * static graph, env and tabular Q-learner.  Every function cites the reference lines it
 * follows (paths relative to the reference root).  Compile with -ffp-contract=off so that the
 * float64 TD update rounds exactly like CPython's.
 */
#include "mdpp_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* node helpers                                                                          */
/* ------------------------------------------------------------------------------------ */
static inline int node_floor(int n) { return n / 36; }
static inline int node_boxnum(int n) { return (n % 36) / 4 + 1; } /* 1..9 */
static inline int node_box18(int n) { return n / 4; }
static inline int node_head(int n) { return n & 3; }
static inline int make_node(int floor, int boxnum, int h) { return floor * 36 + (boxnum - 1) * 4 + h; }
static inline int box18_floor(int b) { return b / 9; }
static inline int box18_num(int b) { return b % 9 + 1; }
static const char HEAD_CH[4] = {'N', 'E', 'S', 'W'};
static inline char floor_ch(int f) { return f ? 'D' : 'U'; }

struct orc_cfg {
    orc_layout lay;
    uint8_t act[ORC_NODES][ORC_ACT]; /* action_dictionary */
    uint8_t no_node[ORC_NODES];
    uint8_t bfs[ORC_NODES][ORC_NODES]; /* 0 = not yet computed */
};

/* reference src/configuration.py:256-262 */
static void num_to_rowcol(const orc_layout *l, int number, int *row, int *col) {
    *col = (number - 1) / l->rows;
    *row = (number - 1) % l->rows;
}
static int rowcol_to_num(const orc_layout *l, int row, int col) { return col * l->rows + (row + 1); }

static int station_index(const orc_layout *l, int floor, int boxnum) {
    /* first matching station decides: `for num, i in enumerate(self.station_box_list)` + break/return */
    for (int i = 0; i < l->n_station; i++)
        if (box18_num(l->station_box[i]) == boxnum && box18_floor(l->station_box[i]) == floor) return i;
    return -1;
}
static int is_disconnected(const orc_layout *l, int node) {
    for (int i = 0; i < l->n_disc; i++)
        if (l->disc_nodes[i] == node) return 1;
    return 0;
}

/* reference src/configuration.py:158-205 give_front_action_node */
static int front_node(const orc_layout *l, int node) {
    int f = node_floor(node), number = node_boxnum(node), h = node_head(node);
    int st = station_index(l, f, number);
    if (st >= 0 && !((l->station_entrance[st] >> h) & 1)) return ORC_NONE; /* :162-169 */
    int row, col;
    num_to_rowcol(l, number, &row, &col);
    if (h == 0) row -= 1; /* N :172-173 */
    if (h == 2) row += 1; /* S */
    if (h == 1) col += 1; /* E */
    if (h == 3) col -= 1; /* W */
    if (row == l->rows || row == -1 || col == l->cols || col == -1) return ORC_NONE; /* :181,201-205 */
    int out = make_node(f, rowcol_to_num(l, row, col), h);
    if (is_disconnected(l, out)) return ORC_NONE; /* :186-187 */
    st = station_index(l, f, node_boxnum(out));
    if (st >= 0) { /* :190-199: heading must be the opposite of a listed entrance */
        int opp = (h + 2) & 3;
        if (!((l->station_entrance[st] >> opp) & 1)) return ORC_NONE;
    }
    return out;
}

/* reference src/configuration.py:207-249 give_back_action_node */
static int back_node(const orc_layout *l, int node) {
    int f = node_floor(node), number = node_boxnum(node), h = node_head(node);
    int st = station_index(l, f, number);
    if (st >= 0) { /* :210-220 */
        int opp = (h + 2) & 3;
        if (!((l->station_entrance[st] >> opp) & 1)) return ORC_NONE;
    }
    int row, col;
    num_to_rowcol(l, number, &row, &col);
    if (h == 0) row += 1;
    if (h == 2) row -= 1;
    if (h == 1) col -= 1;
    if (h == 3) col += 1;
    if (row == l->rows || row == -1 || col == l->cols || col == -1) return ORC_NONE;
    int out = make_node(f, rowcol_to_num(l, row, col), h);
    st = station_index(l, f, node_boxnum(out));
    if (st >= 0 && !((l->station_entrance[st] >> h) & 1)) return ORC_NONE; /* :237-243 */
    return out;
}

orc_cfg *orc_cfg_create(const orc_layout *lay) {
    if (lay->rows != 3 || lay->cols != 3) return NULL;
    orc_cfg *c = (orc_cfg *)calloc(1, sizeof(orc_cfg));
    c->lay = *lay;
    for (int n = 0; n < ORC_NODES; n++) {
        int base = n & ~3, h = n & 3;
        /* reference src/configuration.py:126-129: [S, L, R, F, B, T] */
        c->act[n][0] = (uint8_t)n;
        c->act[n][1] = (uint8_t)(base | ((h + 3) & 3)); /* :147-150, index-1 with Python wraparound */
        c->act[n][2] = (uint8_t)(base | ((h + 1) & 3)); /* :152-156 */
        c->act[n][3] = (uint8_t)front_node(lay, n);
        c->act[n][4] = (uint8_t)back_node(lay, n);
        c->act[n][5] = (uint8_t)(base | ((h + 2) & 3)); /* :251-254, index-2 */
    }
    /* reference src/configuration.py:138-141: every node of a no_box is a no_node */
    for (int i = 0; i < lay->n_no_box; i++)
        for (int h = 0; h < 4; h++) c->no_node[lay->no_box[i] * 4 + h] = 1;
    return c;
}
void orc_cfg_destroy(orc_cfg *c) { free(c); }
int orc_cfg_successor(const orc_cfg *c, int node, int action) { return c->act[node][action]; }
int orc_cfg_in_no_node_list(const orc_cfg *c, int node) { return c->no_node[node]; }

/* reference src/env_gym.py:114-130 get_legal_actions */
int orc_cfg_base_legal(const orc_cfg *c, int node) {
    int m = 0;
    for (int a = 0; a < ORC_ACT; a++) {
        int s = c->act[node][a];
        if (s == ORC_NONE) continue;
        if (c->no_node[s]) continue;
        m |= 1 << a;
    }
    return m;
}

/* reference src/env_gym.py:443-483 breadth_first_search; returns len(actions).
 * The goal test runs on generated successors (:471), 'S' is the first action tried, so
 * len(src,src) = 1; exhaustion returns ['S'] (:483) so "no path" is also 1. */
int orc_cfg_bfs_len(const orc_cfg *cc, int src, int dst) {
    orc_cfg *c = (orc_cfg *)cc;
    if (c->bfs[src][dst]) return c->bfs[src][dst];
    int queue[ORC_NODES], depth[ORC_NODES], closed[ORC_NODES];
    int head = 0, tail = 0, result = 1;
    memset(closed, 0, sizeof closed);
    queue[tail] = src;
    depth[tail++] = 0;
    closed[src] = 1;
    while (head < tail) {
        int n = queue[head], d = depth[head];
        head++;
        int legal = orc_cfg_base_legal(c, n);
        for (int a = 0; a < ORC_ACT; a++) {
            if (!((legal >> a) & 1)) continue;
            if (a == 4) continue; /* :465 'B' skipped */
            int item = c->act[n][a];
            if (item == dst) { result = d + 1; goto done; }
            if (closed[item]) continue;
            closed[item] = 1;
            queue[tail] = item;
            depth[tail++] = d + 1;
        }
    }
done:
    c->bfs[src][dst] = (uint8_t)result;
    return result;
}

/* ------------------------------------------------------------------------------------ */
/* environment                                                                           */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    int node, dest_index, choose[3], done, done_count;
} orc_car;

struct orc_env {
    const orc_cfg *cfg;
    int n_cars;
    orc_car cars[ORC_MAX_CARS];
    /* bad states: open-addressing set of heading-stripped state strings */
    int bad_on;
    char **bad_tab;
    uint32_t bad_cap;
    char *bad_text;
    int bad_shared; /* the set belongs to another env of the same batch (orc_env_share_bad_states) */
};

static uint32_t str_hash(const char *s) {
    uint32_t h = 2166136261u;
    for (; *s; s++) h = (h ^ (uint8_t)*s) * 16777619u;
    return h;
}

orc_env *orc_env_create(const orc_cfg *c) {
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    e->cfg = c;
    return e;
}
void orc_env_destroy(orc_env *e) {
    if (!e->bad_shared) {
        free(e->bad_tab);
        free(e->bad_text);
    }
    free(e);
}

/* reference src/env_gym.py:71-83: file lines, '' skipped; failure to open => flag False */
/* a batch of envs reads ONE parsed bad-states set: env `e` borrows `owner`'s (which must outlive it) */
void orc_env_share_bad_states(orc_env *e, const orc_env *owner) {
    if (!e->bad_shared) {
        free(e->bad_tab);
        free(e->bad_text);
    }
    e->bad_on = owner->bad_on;
    e->bad_tab = owner->bad_tab;
    e->bad_cap = owner->bad_cap;
    e->bad_text = owner->bad_text;
    e->bad_shared = 1;
}
void orc_env_set_bad_states(orc_env *e, const char *text) {
    if (!e->bad_shared) {
        free(e->bad_tab);
        free(e->bad_text);
    }
    e->bad_shared = 0;
    e->bad_tab = NULL;
    e->bad_text = NULL;
    e->bad_on = 0;
    if (!text) return;
    e->bad_on = 1;
    size_t len = strlen(text), lines = 1;
    e->bad_text = (char *)malloc(len + 1);
    memcpy(e->bad_text, text, len + 1);
    for (size_t i = 0; i < len; i++) lines += text[i] == '\n';
    e->bad_cap = 16;
    while (e->bad_cap < 4 * lines) e->bad_cap <<= 1;
    e->bad_tab = (char **)calloc(e->bad_cap, sizeof(char *));
    char *p = e->bad_text;
    while (*p) {
        char *q = p;
        while (*q && *q != '\n') q++;
        int last = (*q == 0);
        *q = 0;
        if (*p) {
            uint32_t h = str_hash(p) & (e->bad_cap - 1);
            while (e->bad_tab[h] && strcmp(e->bad_tab[h], p)) h = (h + 1) & (e->bad_cap - 1);
            e->bad_tab[h] = p;
        }
        if (last) break;
        p = q + 1;
    }
}
static int bad_contains(const orc_env *e, const char *s) {
    uint32_t h = str_hash(s) & (e->bad_cap - 1);
    while (e->bad_tab[h]) {
        if (!strcmp(e->bad_tab[h], s)) return 1;
        h = (h + 1) & (e->bad_cap - 1);
    }
    return 0;
}

/* reference src/env_gym.py:643-644 reset -> __init__: cars cleared; bad-states kept (hasattr guard :71) */
void orc_env_reset(orc_env *e) {
    e->n_cars = 0;
    memset(e->cars, 0, sizeof e->cars);
}

/* reference src/env_gym.py:655-672 + Car.__init__ src/configuration.py:19-31.
 * The cars_info start node / dest index are always overwritten by initialize(). */
void orc_env_initialize_cars(orc_env *e, int n) {
    e->n_cars = n;
    for (int i = 0; i < n; i++) {
        memset(&e->cars[i], 0, sizeof(orc_car));
        e->cars[i].done = 0;
        e->cars[i].done_count = 5; /* training_done_count, configuration.py:30 */
    }
}

/* reference src/env_gym.py:331-352 (heading choice supplied by the caller) */
void orc_env_initialize(orc_env *e, const int32_t *dest_index, const int32_t *start_nodes, const int32_t *choose3) {
    for (int i = 0; i < e->n_cars; i++) {
        e->cars[i].node = start_nodes[i];
        e->cars[i].dest_index = dest_index[i];
        for (int k = 0; k < 3; k++) e->cars[i].choose[k] = choose3[3 * i + k];
    }
}

static int car_dest_node(const orc_env *e, const orc_car *c, int di) {
    return e->cfg->lay.dest_nodes[di][c->choose[di]];
}

/* ordering of ascend_array, reference src/env_gym.py:267-279: sort (int(box number), [src, dst]);
 * ties compare the strings, where 'D' < 'U' and digits compare as characters. */
static int other_less(const uint8_t *a, const uint8_t *b) {
    int ka[5] = {box18_num(a[0]), !box18_floor(a[0]), !box18_floor(a[1]), box18_num(a[1]), 0};
    int kb[5] = {box18_num(b[0]), !box18_floor(b[0]), !box18_floor(b[1]), box18_num(b[1]), 0};
    for (int i = 0; i < 4; i++)
        if (ka[i] != kb[i]) return ka[i] < kb[i];
    return 0;
}
static void sort_others(orc_state *s) {
    for (int i = 1; i < s->n_others; i++) { /* insertion sort */
        uint8_t t0 = s->others[i][0], t1 = s->others[i][1];
        int j = i - 1;
        uint8_t t[2] = {t0, t1};
        while (j >= 0 && other_less(t, s->others[j])) {
            s->others[j + 1][0] = s->others[j][0];
            s->others[j + 1][1] = s->others[j][1];
            j--;
        }
        s->others[j + 1][0] = t0;
        s->others[j + 1][1] = t1;
    }
}

/* reference src/env_gym.py:281-329 encodestate */
void orc_env_encodestate(orc_env *e, int car, int deadlock_id, orc_state *out) {
    const orc_layout *l = &e->cfg->lay;
    memset(out, 0, sizeof *out);
    out->pnode = (uint8_t)e->cars[car].node;
    out->dnode = (uint8_t)car_dest_node(e, &e->cars[car], e->cars[car].dest_index); /* :298-300 */
    for (int i = 0; i < e->n_cars; i++) {
        orc_car *c = &e->cars[i];
        if (i == car) continue;
        if (deadlock_id == 1 && c->done) { /* :307-318 ghost of a finished car, 5 encodes */
            c->done_count -= 1;
            if (c->done_count < 0) continue;
            int fin = car_dest_node(e, c, l->tcdi);
            out->others[out->n_others][0] = (uint8_t)node_box18(fin);
            out->others[out->n_others][1] = (uint8_t)node_box18(fin);
            out->n_others++;
            continue;
        }
        if (deadlock_id == 0 && c->done) continue; /* :319-321 */
        out->others[out->n_others][0] = (uint8_t)node_box18(c->node); /* :322-328 */
        out->others[out->n_others][1] = (uint8_t)node_box18(car_dest_node(e, c, c->dest_index));
        out->n_others++;
    }
    sort_others(out); /* :329 ascend_array */
}

/* reference src/env_gym.py:143-186 */
static int selective_blocked(const orc_env *e, const orc_state *s, int action) {
    const orc_layout *l = &e->cfg->lay;
    if (action == 0 || action == 1 || action == 2 || action == 5) return 0; /* S L R T :156-157 */
    if (action == 4) return 1;                                               /* B :158-159 */
    int succ = e->cfg->act[s->pnode][action];
    int dest_box = node_box18(s->dnode), succ_box = node_box18(succ);
    for (int k = 0; k < l->n_sel; k++) { /* :176-185 */
        if (l->sel_key[k] != dest_box) continue;
        int in = 0;
        for (int j = 0; j < l->sel_n[k]; j++) in |= l->sel_boxes[k][j] == succ_box;
        if (!in) continue;
        for (int o = 0; o < s->n_others; o++)
            for (int j = 0; j < l->sel_n[k]; j++)
                if (l->sel_boxes[k][j] == s->others[o][0]) return 1;
    }
    return 0;
}

/* reference src/env_gym.py:188-224 getlegalactions_filtered (blocked_nodes_list is always empty: :404) */
int orc_env_legal(const orc_env *e, const orc_state *s) {
    int base = orc_cfg_base_legal(e->cfg, s->pnode), m = 0;
    for (int a = 0; a < ORC_ACT; a++) {
        if (!((base >> a) & 1)) continue;
        int succ = e->cfg->act[s->pnode][a];
        int crashed = selective_blocked(e, s, a);
        for (int o = 0; o < s->n_others; o++) /* :211-217 box NUMBER compare, floor ignored */
            if (node_boxnum(succ) == box18_num(s->others[o][0])) crashed = 1;
        if (!crashed) m |= 1 << a;
    }
    return m;
}

/* reference src/env_gym.py:256-265 */
void orc_env_successor_state(const orc_env *e, const orc_state *s, int action, orc_state *out) {
    *out = *s;
    out->pnode = e->cfg->act[s->pnode][action];
}

static int put_node(char *b, int n) {
    if (n == ORC_NONE) return sprintf(b, "?YY");
    return sprintf(b, "%c%d%c", floor_ch(node_floor(n)), node_boxnum(n), HEAD_CH[node_head(n)]);
}
static int put_box(char *b, int box) { return sprintf(b, "%c%d", floor_ch(box18_floor(box)), box18_num(box)); }

static int render(const orc_state *s, char *buf, int strip) {
    char *p = buf;
    if (strip) {
        p += put_box(p, node_box18(s->pnode));
        *p++ = 'x';
        p += put_box(p, node_box18(s->dnode));
    } else {
        p += put_node(p, s->pnode);
        *p++ = 'x';
        p += put_node(p, s->dnode);
    }
    for (int o = 0; o < s->n_others; o++) {
        *p++ = 'z';
        p += put_box(p, s->others[o][0]);
        *p++ = 'x';
        p += put_box(p, s->others[o][1]);
    }
    *p = 0;
    return (int)(p - buf);
}
/* reference src/env_gym.py:237-251 array_to_state */
int orc_state_to_string(const orc_state *s, char *buf, int cap) {
    char tmp[128];
    int n = render(s, tmp, 0);
    if (n + 1 > cap) return -1;
    memcpy(buf, tmp, (size_t)n + 1);
    return n;
}
static int parse_box(const char **pp) {
    const char *p = *pp;
    int f = (*p == 'D');
    p++;
    int num = 0;
    while (*p >= '0' && *p <= '9') num = num * 10 + (*p++ - '0');
    *pp = p;
    return f * 9 + num - 1;
}
static int parse_head(const char **pp) {
    char c = *(*pp)++;
    return c == 'N' ? 0 : c == 'E' ? 1 : c == 'S' ? 2 : 3;
}
/* reference src/env_gym.py:226-235 state_to_array */
int orc_state_from_string(const char *str, orc_state *out) {
    const char *p = str;
    memset(out, 0, sizeof *out);
    int b = parse_box(&p);
    out->pnode = (uint8_t)(b * 4 + parse_head(&p));
    if (*p++ != 'x') return -1;
    b = parse_box(&p);
    out->dnode = (uint8_t)(b * 4 + parse_head(&p));
    while (*p == 'z') {
        p++;
        if (out->n_others >= ORC_MAX_CARS - 1) return -1;
        out->others[out->n_others][0] = (uint8_t)parse_box(&p);
        if (*p++ != 'x') return -1;
        out->others[out->n_others][1] = (uint8_t)parse_box(&p);
        out->n_others++;
    }
    return *p ? -1 : 0;
}

/* reference src/env_gym.py:555-620 q_reward and :485-553 dqn_reward */
double orc_env_reward(const orc_env *e, const orc_state *prev, const orc_state *next, int car, int dqn) {
    if (!dqn && e->bad_on) { /* :568-577 (commented out in dqn_reward :498-507) */
        char b[128];
        render(next, b, 1);
        if (bad_contains(e, b)) return -10000.0;
        render(prev, b, 1);
        if (bad_contains(e, b)) return 0;
    }
    int current_steps = orc_cfg_bfs_len(e->cfg, next->pnode, next->dnode);  /* :583-590 */
    int previous_steps = orc_cfg_bfs_len(e->cfg, prev->pnode, next->dnode); /* :584,591-596 */
    long reward = -10 + 50L * (previous_steps - current_steps);            /* :603-605 */
    if (next->pnode == next->dnode) reward += dqn ? 110 : 10010;            /* :606-607 / :536-537 */
    for (int i = 0; i < e->n_cars; i++) {                                   /* :613-619 */
        if (i == car) continue;
        if (node_boxnum(next->pnode) == node_boxnum(e->cars[i].node) &&
            node_floor(next->pnode) == node_floor(e->cars[i].node))
            reward -= dqn ? 100 : 100000;
    }
    return dqn ? (double)reward / 100 : (double)reward; /* :553 */
}

/* reference src/env_gym.py:386-441 update_car (block_destination_nodes forced to 0 at :404) */
void orc_env_update_car(orc_env *e, int car, int node) {
    const orc_layout *l = &e->cfg->lay;
    orc_car *c = &e->cars[car];
    c->node = node; /* :416 */
    if (c->node == car_dest_node(e, c, c->dest_index)) { /* :420 */
        if (c->dest_index == l->tcdi) {                  /* :437-439 */
            c->done = 1;
            c->node = l->no_box[0] * 4 + 3;              /* no_box_list[0] + 'W' */
        }
        c->dest_index = (c->dest_index + 1) % 2;         /* :440-441, len(destination_nodes_list) == 2 */
    }
}

/* reference src/env_gym.py:622-638 */
int orc_env_is_game_ended(const orc_env *e) {
    int count = 0;
    for (int i = 0; i < e->n_cars; i++) count += e->cars[i].done != 0;
    return count == e->n_cars;
}
int orc_env_car_done(const orc_env *e, int car) { return e->cars[car].done; } /* :646-653 */
int orc_env_car_node(const orc_env *e, int car) { return e->cars[car].node; }
int orc_env_car_dest_index(const orc_env *e, int car) { return e->cars[car].dest_index; }
int orc_env_car_done_count(const orc_env *e, int car) { return e->cars[car].done_count; }
void orc_env_set_car(orc_env *e, int car, int node, int dest_index, int done, int done_count) {
    e->cars[car].node = node;
    e->cars[car].dest_index = dest_index;
    e->cars[car].done = done;
    e->cars[car].done_count = done_count;
}

/* ------------------------------------------------------------------------------------ */
/* tabular Q-learner                                                                     */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    orc_state s;
    uint8_t action;
} qkey; /* 16 bytes, zero padded */

struct orc_q {
    int f32;
    double alpha, discount;
    qkey *keys; /* insertion order, like the Python dict */
    double *vals;
    int64_t n, cap;
    int64_t *tab; /* open addressing, -1 empty */
    int64_t tcap;
};

static uint64_t key_hash(const qkey *k) {
    const uint8_t *p = (const uint8_t *)k;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(qkey); i++) h = (h ^ p[i]) * 1099511628211ull;
    return h ^ (h >> 29);
}
static void make_key(qkey *k, const orc_state *s, int action) {
    memset(k, 0, sizeof *k);
    k->s.pnode = s->pnode;
    k->s.dnode = s->dnode;
    k->s.n_others = s->n_others;
    for (int i = 0; i < s->n_others; i++) {
        k->s.others[i][0] = s->others[i][0];
        k->s.others[i][1] = s->others[i][1];
    }
    k->action = (uint8_t)action;
}

orc_q *orc_q_create(int f32) {
    orc_q *q = (orc_q *)calloc(1, sizeof(orc_q));
    q->f32 = f32;
    q->alpha = 0.9;      /* reference src/algorithms/qlearning.py:21 */
    q->discount = 0.15;  /* :22 */
    q->cap = 1024;
    q->keys = (qkey *)malloc(sizeof(qkey) * (size_t)q->cap);
    q->vals = (double *)malloc(sizeof(double) * (size_t)q->cap);
    q->tcap = 4096;
    q->tab = (int64_t *)malloc(sizeof(int64_t) * (size_t)q->tcap);
    for (int64_t i = 0; i < q->tcap; i++) q->tab[i] = -1;
    return q;
}
void orc_q_destroy(orc_q *q) {
    free(q->keys);
    free(q->vals);
    free(q->tab);
    free(q);
}
void orc_q_set_params(orc_q *q, double alpha, double discount) {
    q->alpha = alpha;
    q->discount = discount;
}
static int64_t q_find(const orc_q *q, const qkey *k, int64_t *slot) {
    uint64_t h = key_hash(k) & (uint64_t)(q->tcap - 1);
    while (q->tab[h] >= 0) {
        if (!memcmp(&q->keys[q->tab[h]], k, sizeof(qkey))) {
            if (slot) *slot = (int64_t)h;
            return q->tab[h];
        }
        h = (h + 1) & (uint64_t)(q->tcap - 1);
    }
    if (slot) *slot = (int64_t)h;
    return -1;
}
static int64_t q_insert(orc_q *q, const qkey *k) {
    int64_t slot, idx = q_find(q, k, &slot);
    if (idx >= 0) return idx;
    if (q->n == q->cap) {
        q->cap *= 2;
        q->keys = (qkey *)realloc(q->keys, sizeof(qkey) * (size_t)q->cap);
        q->vals = (double *)realloc(q->vals, sizeof(double) * (size_t)q->cap);
    }
    idx = q->n++;
    q->keys[idx] = *k;
    q->vals[idx] = 0.0;
    q->tab[slot] = idx;
    if (q->n * 2 > q->tcap) { /* rehash */
        q->tcap *= 2;
        q->tab = (int64_t *)realloc(q->tab, sizeof(int64_t) * (size_t)q->tcap);
        for (int64_t i = 0; i < q->tcap; i++) q->tab[i] = -1;
        for (int64_t i = 0; i < q->n; i++) {
            uint64_t h = key_hash(&q->keys[i]) & (uint64_t)(q->tcap - 1);
            while (q->tab[h] >= 0) h = (h + 1) & (uint64_t)(q->tcap - 1);
            q->tab[h] = i;
        }
    }
    return idx;
}

/* reference src/algorithms/qlearning.py:34-43 getqvalue + Counter.__getitem__ :645-647: reading inserts 0 */
double orc_q_get(orc_q *q, const orc_state *s, int action) {
    qkey k;
    make_key(&k, s, action);
    int64_t i = q_insert(q, &k); /* may realloc vals: index first */
    return q->vals[i];
}
double orc_q_peek(const orc_q *q, const orc_state *s, int action, int *present) {
    qkey k;
    make_key(&k, s, action);
    int64_t i = q_find(q, &k, NULL);
    if (present) *present = i >= 0;
    return i >= 0 ? q->vals[i] : 0.0;
}
void orc_q_set(orc_q *q, const orc_state *s, int action, double v) {
    qkey k;
    make_key(&k, s, action);
    int64_t i = q_insert(q, &k);
    q->vals[i] = v;
}

/* reference src/algorithms/qlearning.py:60-72 + Counter.argMax :659-667 (first maximum in legal order) */
int orc_q_greedy(orc_q *q, const orc_env *e, const orc_state *s) {
    int legal = orc_env_legal(e, s), best = -1;
    double bv = 0;
    for (int a = 0; a < ORC_ACT; a++) {
        if (!((legal >> a) & 1)) continue;
        double v = orc_q_get(q, s, a);
        if (best < 0 || v > bv) { best = a; bv = v; }
    }
    return best;
}
/* reference src/algorithms/qlearning.py:45-58 */
double orc_q_value(orc_q *q, const orc_env *e, const orc_state *s) {
    int legal = orc_env_legal(e, s), any = 0;
    double bv = 0.0;
    for (int a = 0; a < ORC_ACT; a++) {
        if (!((legal >> a) & 1)) continue;
        double v = orc_q_get(q, s, a);
        if (!any || v > bv) { any = 1; bv = v; }
    }
    return any ? bv : 0.0;
}
/* reference src/algorithms/qlearning.py:107-116.  Python evaluates Q[s,a] (inserting it), then the
 * next-state value (inserting its legal keys), then Q[s,a] again. */
void orc_q_update(orc_q *q, const orc_env *e, const orc_state *s, int a, const orc_state *ns, double r) {
    double cur = orc_q_get(q, s, a);
    double nv = orc_q_value(q, e, ns);
    double out;
    if (q->f32) {
        float al = (float)q->alpha, g = (float)q->discount, c = (float)cur;
        float t = (float)r + g * (float)nv;
        float d = t - c;
        out = (double)(c + al * d);
    } else {
        out = cur + q->alpha * (r + q->discount * nv - cur);
    }
    orc_q_set(q, s, a, out);
}
int64_t orc_q_size(const orc_q *q) { return q->n; }
void orc_q_entry(const orc_q *q, int64_t i, orc_state *s, int *action, double *v) {
    *s = q->keys[i].s;
    *action = q->keys[i].action;
    *v = q->vals[i];
}

/* reference src/utils/utils.py:12-52 give_action_choose_probability; int(episodes * 0.x) in doubles */
double orc_epsilon(int episodes, int episode, int model) {
    if (model == 0) return 0;
    if (model == 1) return 1;
    if (model == 2) return episode < (int)(episodes * 0.4) ? 1 : 0;
    if (model == 3) {
        if (episode < (int)(episodes * 0.4)) return 1;
        if (episode < (int)(episodes * 0.6)) return 0.8;
        if (episode < (int)(episodes * 0.7)) return 0.5;
        if (episode < (int)(episodes * 0.85)) return 0.3;
        if (episode < (int)(episodes * 0.99)) return 0.15;
        return 0;
    }
    return 0; /* Python returns None for other models */
}

/* Philox4x32, Salmon/Moraes/Dror/Shaw SC'11 (Random123 constants), any round count */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4], int rounds) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* the generator of the batch semantics: the round count the CUDA library is built with (env_core.cuh) */
#ifndef MDPP_PHILOX_ROUNDS
#define MDPP_PHILOX_ROUNDS 7
#endif
void orc_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { orc_philox4x32(ctr, key, out, MDPP_PHILOX_ROUNDS); }
int orc_philox_rounds(void) { return MDPP_PHILOX_ROUNDS; }

/* index-level access for the batched driver (mdpp_oracle_batch.c) */
int64_t orc_q_index(orc_q *q, const orc_state *s, int action) {
    qkey k;
    make_key(&k, s, action);
    return q_insert(q, &k);
}
double orc_q_at(const orc_q *q, int64_t i) { return q->vals[i]; }
void orc_q_set_at(orc_q *q, int64_t i, double v) { q->vals[i] = v; }

/* reference src/algorithms/dqn_open_source.py:245-313 RL.convert: 16*total_cars values in {0,1}.
 * Per listed car 8 values: box NUMBER (floor dropped) as 5 bits MSB first, then the heading as 3 bits
 * with N,W,E,S = 0,1,2,3 (:262-268) and "no heading" (the other cars, listed by box) = -1 = 111.
 * Sources of the primary car and the others in string order, missing cars padded with ones (:303),
 * then the destinations the same way. */
void orc_convert(const orc_state *s, int total_cars, uint8_t *out) {
    static const int dir_code[4] = {0, 2, 3, 1}; /* our N,E,S,W -> the reference's N=0 W=1 E=2 S=3 */
    int n = 1 + s->n_others, k = 0;
    for (int half = 0; half < 2; half++) {
        for (int i = 0; i < total_cars; i++) {
            if (i >= n) {
                for (int b = 0; b < 8; b++) out[k++] = 1;
                continue;
            }
            int number, dir;
            if (i == 0) {
                int node = half ? s->dnode : s->pnode;
                number = node_boxnum(node);
                dir = dir_code[node_head(node)];
            } else {
                number = box18_num(s->others[i - 1][half]);
                dir = 7; /* np.binary_repr(-1, width=3) == '111' */
            }
            for (int b = 4; b >= 0; b--) out[k++] = (uint8_t)((number >> b) & 1);
            for (int b = 2; b >= 0; b--) out[k++] = (uint8_t)((dir >> b) & 1);
        }
    }
}
