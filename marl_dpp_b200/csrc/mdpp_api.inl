// mdpp_api.inl -- host side of the C ABI declared in include/mdpp.h (included by mdpp_kernels.cu).
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace mdpp {

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define MDPP_CUDA(call)                                                                         \
    do {                                                                                        \
        cudaError_t err__ = (call);                                                             \
        if (err__ != cudaSuccess)                                                               \
            return fail(err__ == cudaErrorNoDevice || err__ == cudaErrorInsufficientDriver      \
                            ? MDPP_ENODEVICE : MDPP_ECUDA,                                      \
                        "%s failed: %s", #call, cudaGetErrorString(err__));                     \
    } while (0)

// marks of the dense modes are indexed by the compact key  state * 5 + action  (n_marks = n_states * 5)
constexpr size_t kMaxByteMapMarks = 200 * 1024;   // shared-memory BYTE map of a persistent env CTA
constexpr size_t kMaxDenseMarks = 8 * 200 * 1024; // shared-memory BIT map (atomic OR) up to here; scan mode beyond
constexpr uint32_t kMaxEnvCtas = 160;             // per-CTA bitmap slices per region: sub-step kernels, one CTA per SM
constexpr uint32_t kMaxRoundCtas = 320;           // ... fused round kernels, up to two CTAs per SM (byte-map tables only)
constexpr int kMaxPureRounds = 8;                 // pure-exploration rounds per q_rounds_pure_kernel launch

// a library kernel launch: counted per handle (mdpp_launch_count), error-checked
#define MDPP_LAUNCH(h, call)                                                                     \
    do {                                                                                        \
        MDPP_CUDA(call);                                                                        \
        (h)->launches += 1;                                                                     \
    } while (0)

struct Carve {
    size_t tables, fast, rng, bad, recs, sinfo, qalt, marks, flags, gbits, bits, visited, control, barrier, stage, dl, total;
    uint32_t scan_words;    // scan mode: words of one touched bitmap (a multiple of 4; compact keys in bit planes, scan_slot)
    uint32_t scan_regions;  // ... bitmaps per bank: sub-step i of a launch marks region i of the launch's bank.  THREE banks in
                            // rotation: the pull of a sub-step clears the bitmap of the sub-step BEFORE it (its last reader)
                            // while the next env kernel, which may already run, marks into the third
    uint32_t words_per_cta; // words of one per-CTA touched bitmap (dense mode)
    uint32_t cta_slots, regions; // layout of the bitmap array [parity][region][cta_slots][words_per_cta]
    bool dense;        // small table: shared-memory marks + whole-table ping-pong pull; else global bitmaps + mirror table + scanning pull
    bool scan;         // the global touched bitmaps of the scan mode exist: every non-dense table, and the mid-size dense ones
                       // (shared-memory BIT maps, 3 cars), whose pure-exploration rounds take the scan mode's fused env pass
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static Carve carve(const mdpp_config &cfg, int64_t n_envs) {
    Carve c{};
    size_t off = 0;
    const size_t keys = (size_t)cfg.n_states * MDPP_Q_ROW;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    c.tables = take(sizeof(SmemTables));
    c.fast = take(sizeof(FastTabs));
    c.rng = take((size_t)n_envs * 4 * 3);
    c.bad = take((((size_t)cfg.n_box_states + 31) / 32) * 4 + 4);
    c.recs = take((size_t)n_envs * 16);
    c.sinfo = take((size_t)cfg.n_states * 4);
    // dense mode: the mark byte map of a persistent CTA lives in shared memory (one byte per key) and the pull
    // kernel sweeps the whole table once per sub-step (copy + substitute the marked keys) -- right while
    // the table is a few hundred KB.  Larger tables keep touched bitmaps in global memory and a mirror of the table
    // (scan mode, q_pull_scan_kernel).
    const size_t marks = (size_t)cfg.n_states * MDPP_N_ACT;
    c.dense = marks <= kMaxDenseMarks;
    if (const char *env = getenv("MDPP_DENSE")) c.dense = c.dense && atoi(env) != 0;
    c.words_per_cta = (uint32_t)(((marks + 31) / 32 + 7) & ~(size_t)7); // whole blocks of 8 words (slice_index)
    c.cta_slots = marks <= kMaxByteMapMarks ? kMaxRoundCtas : kMaxEnvCtas;
    // sub-step kernels use region 0; fused round: K0 map 0, K0 map 1, K1 map 1; several pure-exploration rounds per
    // launch (1-2 cars): region r * n_cars + c for sub-step (r, c)
    c.regions = marks <= kMaxByteMapMarks ? (cfg.n_cars <= 2 ? (uint32_t)std::max(3, kMaxPureRounds * cfg.n_cars) : 3u) : 1u;
    if (c.dense) {
        c.qalt = take(keys * 4);
        c.marks = take((size_t)2 * c.regions * c.cta_slots * c.words_per_cta * 4); // [parity][region][cta][word]
        c.flags = take((size_t)2 * c.cta_slots * 4);
        c.gbits = take((size_t)c.regions * c.words_per_cta * 4); // reduced marks of a multi-round launch, one bitmap per sub-step
        c.bits = 0;
    } else {
        c.marks = c.flags = c.gbits = 0;
        c.qalt = take(keys * 4); // the mirror table
    }
    c.scan = !c.dense || marks > kMaxByteMapMarks;
    if (const char *env = getenv("MDPP_HYBRID")) c.scan = c.scan && (!c.dense || atoi(env) != 0);
    if (c.scan) {
        c.scan_words = (uint32_t)(((marks + 31) / 32 + 3) & ~(size_t)3);
        // pure-exploration rounds share a launch (q_rounds_pure_scan_kernel): as many rounds as fit 34 MB of bitmaps per bank
        // (4 cars: 8 rounds of 4 x 1.03 MB; 4 rounds per launch measured 49.4 us per round, 8 rounds 46.4, 1 round 59.6)
        const size_t per_round = (size_t)cfg.n_cars * c.scan_words * 4;
        size_t bank_mb = 34;
        if (const char *env = getenv("MDPP_SCAN_BANK_MB")) bank_mb = (size_t)std::max(1, atoi(env)); // A/B runs
        const int rounds = (int)std::max<size_t>(1, std::min<size_t>((size_t)kMaxPureRounds, (bank_mb << 20) / per_round));
        c.scan_regions = (uint32_t)(rounds * cfg.n_cars);
        c.bits = take((size_t)3 * c.scan_regions * c.scan_words * 4);
    }
    c.visited = take((size_t)cfg.n_states * 4);
    c.control = take(sizeof(Control));
    c.barrier = take(64); // grid barrier of the cooperative kernels: word 0 arrival counter (only ever grows), word 1 its value at the next kernel's start; zero at create
    c.stage = take((size_t)n_envs * 16);
    c.dl = take((size_t)n_envs * 4);
    c.total = off;
    return c;
}

} // namespace mdpp

struct mdpp_handle {
    mdpp_config cfg;
    mdpp::DevCfg dev;
    mdpp::Carve cv;
    int64_t n_envs;
    uint32_t env_id_base;
    char *ws;
    int sm_count;
    uint32_t parity;        // sub-step parity of the marks (dense mode)
    // scan mode: the mirror table (workspace) equals the caller's table except on the keys of one bitmap, where it is
    // one update behind (q_pull_scan_kernel); between API calls it is either a complete copy or invalid
    uint32_t *scan_stale;   // that bitmap, or NULL
    uint32_t scan_bank;     // bank 0..2 the next scan-mode launch marks into
    bool mirror_ok;         // the workspace copy and the caller's table mirror_q are such a pair (see mdpp_q_table_written)
    const float *mirror_q;
    uint64_t rng_round;     // round whose Philox words the car-0 launch published
    uint32_t rng_seed;
    bool pdl;               // programmatic dependent launch between learner kernels (MDPP_PDL=0 disables)
    bool k1_rounds;         // MDPP_K1_ROUNDS=1: rounds with greedy decisions through the fused K0 / K1 pair as well
    bool mixed_rounds;      // rounds with greedy decisions: several per cooperative launch (MDPP_MIXED_ROUNDS=0: sub-step kernels)
    int coop_ok;            // -1 unknown, 0 no cooperative launch / not co-resident, 1 ok
    bool multi_round;       // several pure-exploration rounds per launch (MDPP_MULTI_ROUND=0 disables)
    bool multi_pull;        // ... followed by one reduce + one cluster pull launch (MDPP_MULTI_PULL=0: dense pulls)
    bool pure_scan;         // large tables: all cars of pure-exploration rounds in one env pass (MDPP_PURE_SCAN=0: sub-step kernels)
    bool chain;             // the previous launch on the stream was one of our learner kernels
    bool fast_rollout;      // the table-driven rollout kernel applies (BFS lengths fit its move table)
    uint64_t launches;      // learner kernels launched through this handle (bench.py's gpu_launches)
    uint32_t smem_configured; // env-kernel variants whose dynamic shared-memory limit has been raised
    uint32_t round_parity;    // fused rounds alternate between two sets of mark regions
    // Upper bound on every env's episode counter at the current point of the stream.  While it is below the
    // end of an epsilon = 1 schedule segment no car can act greedily, so the fused round needs no catch-up
    // kernel.  Grows by one per round launched; tightened to the device's own maximum at every stats fetch
    // as long as only learner kernels (which maintain that maximum) have advanced episodes since the last reset.
    int64_t ep_bound;
    bool ep_device_max_valid;
    mdpp_stats *stat_slots; // pinned host staging for the privatised device counters
    mdpp::HostStats *host_stats;  // pinned, device-visible: totals published by stats_publish_kernel (end-to-end calls)
    unsigned long long host_seq;
    // replica averaging (q_sync.cuh): every rank's exchange buffer as seen from this process
    int sync_world, sync_rank, sync_interval;
    uint32_t sync_epoch;    // calls so far; the flags of call k carry k (every rank makes the same calls)
    uint64_t sync_count;
    void *sync_bufs[MDPP_SYNC_MAX_RANKS];
    // the exchange runs on a stream of its own, beside the env kernel that follows it on the caller's stream
    cudaStream_t sync_stream;
    cudaEvent_t sync_fork, sync_join;
    bool sync_pending;      // an exchange is in flight: the caller's stream must wait for sync_join before the table is touched
    // scan mode, pure-exploration rounds: the pulls of a group of rounds run on a stream of the handle's own, BESIDE the
    // env pass of the next group on the caller's stream (q_pure_scan_rounds_launch)
    bool scan_overlap;      // MDPP_SCAN_OVERLAP=0 keeps everything on the caller's stream
    cudaStream_t pull_stream;
    cudaEvent_t ev_env[2], ev_pull[2];
    uint32_t ov_groups;     // groups launched since the streams last joined
    bool pulls_pending;     // pulls are in flight on pull_stream: the caller's stream waits for ev_pull[(ov_groups - 1) & 1] first
};

using namespace mdpp;

static int validate_config(const mdpp_config *cfg) {
    if (!cfg) return fail(MDPP_EINVAL, "cfg is NULL");
    if (cfg->abi_version != MDPP_ABI_VERSION)
        return fail(MDPP_EINVAL, "abi_version %d != %d", cfg->abi_version, MDPP_ABI_VERSION);
    if (cfg->n_cars < 1 || cfg->n_cars > MDPP_MAX_CARS) return fail(MDPP_EINVAL, "n_cars %d out of 1..%d", cfg->n_cars, MDPP_MAX_CARS);
    if (cfg->n_local_nodes != 36 && cfg->n_local_nodes != 72) return fail(MDPP_EINVAL, "n_local_nodes must be 36 or 72");
    if (cfg->n_src_boxes != 9 && cfg->n_src_boxes != 18) return fail(MDPP_EINVAL, "n_src_boxes must be 9 or 18");
    if (cfg->n_dest_nodes < 1 || cfg->n_dest_nodes > MDPP_MAX_DEST || cfg->n_dest_boxes < 1 || cfg->n_dest_boxes > MDPP_MAX_DEST)
        return fail(MDPP_EINVAL, "destination counts out of range");
    if (cfg->tcdi != 0 && cfg->tcdi != 1) return fail(MDPP_EINVAL, "training_complete_destination_index must be 0 or 1");
    if (cfg->n_states <= 0 || cfg->n_states * MDPP_Q_ROW >= (int64_t)1 << 32)
        return fail(MDPP_EINVAL, "n_states %lld does not fit 32-bit keys", (long long)cfg->n_states);
    if (cfg->step_cap < 0 || cfg->step_cap > 65535) return fail(MDPP_EINVAL, "step_cap out of 0..65535");
    return MDPP_OK;
}

// The per-env episode counter is 16 bits of the record (rec.z & 0xFFFF): a larger budget would wrap it, restart
// the epsilon schedule and never freeze the env.
static int validate_learner(const mdpp_learner *lp) {
    if (!lp) return fail(MDPP_EINVAL, "learner is NULL");
    if (lp->episodes < 0 || lp->episodes > 65535)
        return fail(MDPP_EINVAL, "learner.episodes %d out of 0..65535 (16-bit per-env episode counter)", lp->episodes);
    return MDPP_OK;
}

static int find_base(const uint8_t *tab, int n) {
    for (int i = 0; i < n; i++)
        if (tab[i] == 0) return i;
    return 0;
}

static void make_devcfg(const mdpp_config &c, DevCfg &d) {
    memset(&d, 0, sizeof d);
    d.n_cars = c.n_cars;
    d.n_local_nodes = c.n_local_nodes;
    d.n_dest_nodes = c.n_dest_nodes;
    d.n_dest_boxes = c.n_dest_boxes;
    d.n_src_boxes = c.n_src_boxes;
    d.node_base = find_base(c.node_local, MDPP_NODES);
    d.box_base = find_base(c.src_box_idx, MDPP_BOXES);
    d.parking_node = c.parking_node;
    d.tcdi = c.tcdi;
    d.bad_on = c.bad_states_on;
    d.step_cap = c.step_cap;
    for (int i = 0; i < c.n_cars; i++) {
        for (int di = 0; di < 2; di++) {
            uint32_t dn = c.car_dest[i][di];
            d.car_dnidx[i][di] = dn;
            d.car_dnode[i][di] = c.dest_node[dn];
            d.car_dbox[i][di] = c.dest_node_box[dn];
            d.car_dbox18[i][di] = c.dest_box[c.dest_node_box[dn]];
            d.car_sel[i][di] = c.sel_mask[c.dest_node_box[dn]];
        }
        // ghost of a finished car: [final box, final box] (reference src/env_gym.py:313-317)
        uint32_t fb = d.car_dbox18[i][c.tcdi];
        d.car_ghost_box18[i] = fb;
        d.car_ghost_code[i] = 1u + (fb - d.box_base) * c.n_dest_boxes + d.car_dbox[i][c.tcdi];
        d.start_box18[i] = c.start_box[i];
        d.start_di[i] = c.start_di[i];
    }
    for (int i = 0; i < c.n_dest_nodes; i++) {
        d.dn_node[i] = c.dest_node[i];
        d.dn_dbox[i] = c.dest_node_box[i];
    }
    for (int i = 0; i < c.n_dest_boxes; i++) d.dbox_sel[i] = c.sel_mask[i];
}

// Table image of the fused round kernel (q_round.cuh): everything the reference derives from a car's
// (node, destination_index), per car, as the primary and as an "other".
static bool build_fast_tabs(const mdpp_config &c, FastTabs &t) {
    memset(&t, 0, sizeof t);
    bool ok = true; // every BFS length fits the 6 bits of the move table
    for (int car = 0; car < c.n_cars; car++) {
        for (int di = 0; di < 2; di++) {
            const uint32_t dn = c.car_dest[car][di], dnode = c.dest_node[dn], dbox = c.dest_node_box[dn];
            const uint32_t sel = c.sel_mask[dbox];
            for (int node = 0; node < MDPP_NODES; node++) {
                const int idx = node | (di << 7);
                const uint32_t box18 = (uint32_t)node >> 2;
                if (c.node_local[node] != MDPP_NONE) { // as the primary
                    const uint32_t fwd = c.fwd[node];
                    uint32_t y = 1u << (box18 % 9);
                    if (fwd != MDPP_NONE) {
                        y |= 1u << (9 + (fwd >> 2) % 9);
                        if ((sel >> (fwd >> 2)) & 1u) y |= 1u << 18;
                    }
                    t.own[car][idx].x = (dn * (uint32_t)c.n_local_nodes + c.node_local[node]) | (dnode << 10) |
                                        ((fwd == MDPP_NONE ? 127u : fwd) << 17) | ((uint32_t)c.base_legal[node] << 24) |
                                        (dbox << 29);
                    t.own[car][idx].y = y;
                    t.own[car][idx].z = sel;
                    t.own[car][idx].w = c.bfs_len[dn][node];
                }
                if (c.src_box_idx[box18] != MDPP_NONE) { // as a listed other: (box, destination box)
                    const uint32_t code = 1u + (uint32_t)c.src_box_idx[box18] * (uint32_t)c.n_dest_boxes + dbox;
                    t.oth[car][idx].x = (1u << box18) | (code << 18);
                    t.oth[car][idx].y = 1u << (box18 % 9);
                }
            }
        }
        for (int di = 0; di < 2; di++) { // the move itself: successor + update_car, per action
            const uint32_t dn = c.car_dest[car][di], dnode = c.dest_node[dn];
            for (int node = 0; node < MDPP_NODES; node++)
                for (int a = 0; a < MDPP_N_ACT; a++) {
                    static const int rot[MDPP_N_ACT] = {0, 3, 1, 0, 2}; // S, L (heading - 1), R (+ 1), F, T (- 2)
                    uint32_t nn = a == MDPP_A_F ? c.fwd[node] : (uint32_t)((node & ~3) | ((node + rot[a]) & 3));
                    uint32_t ndi = (uint32_t)di, done = 0, reached = 0;
                    if (nn == MDPP_NONE) nn = (uint32_t)node; // F off the graph is never legal
                    const uint32_t bfs = c.bfs_len[dn][nn];
                    if (bfs > 63) ok = false;                  // does not fit the move table: rollout falls back
                    if (nn == dnode) {
                        reached = 1;
                        if (di == c.tcdi) {
                            done = 1;
                            nn = (uint32_t)c.parking_node;
                        }
                        ndi ^= 1u;
                    }
                    t.next[car][node | (di << 7)][a] = (uint16_t)((nn & 127u) | (ndi << 7) | (done << 8) | (reached << 9) | ((bfs & 63u) << 10));
                }
        }
        // ghost of a finished car: [final box, final box] (reference src/env_gym.py:313-317)
        const uint32_t gdbox = c.dest_node_box[c.car_dest[car][c.tcdi]], gbox18 = c.dest_box[gdbox];
        if (c.src_box_idx[gbox18] != MDPP_NONE) {
            const uint32_t code = 1u + (uint32_t)c.src_box_idx[gbox18] * (uint32_t)c.n_dest_boxes + gdbox;
            t.ghost[car].x = (1u << gbox18) | (code << 18);
            t.ghost[car].y = 1u << (gbox18 % 9);
        }
    }
    for (int m = 0; m < 32; m++) {
        uint32_t packed = 0, n = 0;
        for (int k = 0; k < 5; k++)
            if ((m >> k) & 1) packed |= (uint32_t)k << (3 * n++);
        t.nth[m] = packed;
    }
    return ok;
}

extern "C" {

int mdpp_abi_version(void) { return MDPP_ABI_VERSION; }

int64_t mdpp_fast_tables(const mdpp_config *cfg, void *out_host, int64_t bytes) {
    if (validate_config(cfg)) return -MDPP_EINVAL;
    if (!out_host) return (int64_t)sizeof(FastTabs);
    if (bytes < (int64_t)sizeof(FastTabs)) return -fail(MDPP_EINVAL, "buffer smaller than the table image");
    return build_fast_tabs(*cfg, *(FastTabs *)out_host) ? (int64_t)sizeof(FastTabs) : 0;
}
const char *mdpp_last_error_string(void) { return g_err; }

int64_t mdpp_workspace_bytes(const mdpp_config *cfg, int64_t n_envs) {
    if (validate_config(cfg) != MDPP_OK || n_envs <= 0) return -1;
    return (int64_t)carve(*cfg, n_envs).total;
}

int mdpp_create(const mdpp_config *cfg, int64_t n_envs, int64_t env_id_base, void *workspace,
                int64_t workspace_bytes, const uint32_t *bad_bitmap_host, void *stream, mdpp_handle **out) {
    if (!out) return fail(MDPP_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = validate_config(cfg);
    if (rc) return rc;
    if (n_envs <= 0 || n_envs >= (int64_t)1 << 31) return fail(MDPP_EINVAL, "n_envs out of range");
    if (env_id_base < 0 || env_id_base + n_envs > (int64_t)1 << 32) return fail(MDPP_EINVAL, "env ids must fit 32 bits");
    if (cfg->bad_states_on && !bad_bitmap_host) return fail(MDPP_EINVAL, "bad_states_on without a bitmap");
    int ndev = 0;
    MDPP_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return fail(MDPP_ENODEVICE, "no CUDA device: this library has no CPU fallback");
    Carve cv = carve(*cfg, n_envs);
    if (!workspace || workspace_bytes < (int64_t)cv.total)
        return fail(MDPP_EWORKSPACE, "workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)cv.total);
    if ((uintptr_t)workspace & 255) return fail(MDPP_EWORKSPACE, "workspace must be 256-byte aligned");
    mdpp_handle *h = new (std::nothrow) mdpp_handle();
    if (!h) return fail(MDPP_EINVAL, "out of host memory");
    h->cfg = *cfg;
    h->cv = cv;
    h->n_envs = n_envs;
    h->env_id_base = (uint32_t)env_id_base;
    h->ws = (char *)workspace;
    make_devcfg(*cfg, h->dev);
    h->rng_round = ~0ull;
    h->pdl = !(getenv("MDPP_PDL") && atoi(getenv("MDPP_PDL")) == 0);
    h->k1_rounds = getenv("MDPP_K1_ROUNDS") && atoi(getenv("MDPP_K1_ROUNDS")) != 0;
    h->mixed_rounds = !(getenv("MDPP_MIXED_ROUNDS") && atoi(getenv("MDPP_MIXED_ROUNDS")) == 0);
    h->coop_ok = -1;
    h->multi_round = !(getenv("MDPP_MULTI_ROUND") && atoi(getenv("MDPP_MULTI_ROUND")) == 0);
    h->multi_pull = !(getenv("MDPP_MULTI_PULL") && atoi(getenv("MDPP_MULTI_PULL")) == 0);
    h->pure_scan = !(getenv("MDPP_PURE_SCAN") && atoi(getenv("MDPP_PURE_SCAN")) == 0);
    h->scan_overlap = !(getenv("MDPP_SCAN_OVERLAP") && atoi(getenv("MDPP_SCAN_OVERLAP")) == 0);
    h->chain = false;
    h->smem_configured = 0;
    h->round_parity = 0;
    h->launches = 0;
    h->ep_bound = 0; // the workspace is zeroed below: every episode counter starts at 0
    h->ep_device_max_valid = true;
    h->dev.tables = (const uint32_t *)(h->ws + cv.tables);
    h->dev.bad_bitmap = (const uint32_t *)(h->ws + cv.bad);
    int devid = 0;
    cudaGetDevice(&devid);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, devid);
    cudaStream_t s = (cudaStream_t)stream;
    SmemTables img;
    memset(&img, 0, sizeof img);
    for (int n = 0; n < MDPP_NODES; n++) img.nodeinfo[n] = (uint16_t)(cfg->fwd[n] | (cfg->base_legal[n] << 8));
    memcpy(img.bfs, cfg->bfs_len, sizeof img.bfs);
    FastTabs *ft = new (std::nothrow) FastTabs();
    if (!ft) { delete h; return fail(MDPP_EINVAL, "out of host memory"); }
    h->fast_rollout = build_fast_tabs(*cfg, *ft) && !(getenv("MDPP_FAST_ROLLOUT") && atoi(getenv("MDPP_FAST_ROLLOUT")) == 0);
    cudaError_t e = cudaMemsetAsync(h->ws, 0, cv.total, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->ws + cv.fast, ft, sizeof(FastTabs), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->ws + cv.dl, 0xFF, (size_t)n_envs * 4, s); // no deadlock state yet
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->ws + cv.tables, &img, sizeof img, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && bad_bitmap_host)
        e = cudaMemcpyAsync(h->ws + cv.bad, bad_bitmap_host, (((size_t)cfg->n_box_states + 31) / 32) * 4,
                            cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) { // per-state info words of the pull kernels
        uint32_t *sinfo = (uint32_t *)(h->ws + cv.sinfo);
        const unsigned g = (unsigned)((cfg->n_states + kThreads - 1) / kThreads);
        switch (cfg->n_cars) {
        case 1: state_info_kernel<1><<<g, kThreads, 0, s>>>(h->dev, sinfo, cfg->n_states); break;
        case 2: state_info_kernel<2><<<g, kThreads, 0, s>>>(h->dev, sinfo, cfg->n_states); break;
        case 3: state_info_kernel<3><<<g, kThreads, 0, s>>>(h->dev, sinfo, cfg->n_states); break;
        case 4: state_info_kernel<4><<<g, kThreads, 0, s>>>(h->dev, sinfo, cfg->n_states); break;
        default: state_info_kernel<5><<<g, kThreads, 0, s>>>(h->dev, sinfo, cfg->n_states); break;
        }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s); // host images go out of scope
    delete ft;
    if (e == cudaSuccess) e = cudaMallocHost((void **)&h->stat_slots, sizeof(Control));
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&h->host_stats, sizeof(HostStats), cudaHostAllocMapped);
    if (e == cudaSuccess) memset(h->host_stats, 0, sizeof(HostStats));
    if (e != cudaSuccess) {
        delete h;
        return fail(MDPP_ECUDA, "workspace initialisation failed: %s", cudaGetErrorString(e));
    }
    *out = h;
    return MDPP_OK;
}

int mdpp_destroy(mdpp_handle *h) {
    if (h && h->stat_slots) cudaFreeHost(h->stat_slots);
    if (h && h->host_stats) cudaFreeHost(h->host_stats);
    if (h && h->pull_stream) {
        cudaStreamSynchronize(h->pull_stream);
        for (int i = 0; i < 2; i++) {
            cudaEventDestroy(h->ev_env[i]);
            cudaEventDestroy(h->ev_pull[i]);
        }
        cudaStreamDestroy(h->pull_stream);
    }
    if (h && h->sync_stream) {
        cudaStreamSynchronize(h->sync_stream);
        cudaEventDestroy(h->sync_fork);
        cudaEventDestroy(h->sync_join);
        cudaStreamDestroy(h->sync_stream);
    }
    delete h;
    return MDPP_OK;
}

void *mdpp_env_records(mdpp_handle *h) { return h ? h->ws + h->cv.recs : nullptr; }
void *mdpp_visited_bits(mdpp_handle *h) { return h ? h->ws + h->cv.visited : nullptr; }
void *mdpp_deadlock_states(mdpp_handle *h) { return h ? h->ws + h->cv.dl : nullptr; }
uint64_t mdpp_launch_count(mdpp_handle *h) { return h ? h->launches : 0; }

#define MDPP_DISPATCH_NC(h, KERNEL, grid, s, ...)                                               \
    switch ((h)->cfg.n_cars) {                                                                  \
    case 1: KERNEL<1><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;                            \
    case 2: KERNEL<2><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;                            \
    case 3: KERNEL<3><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;                            \
    case 4: KERNEL<4><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;                            \
    default: KERNEL<5><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;                           \
    }

// kernels templated on <n_cars, car slot>
#define MDPP_DISPATCH_C(NCV, c, KERNEL, grid, s, ...)                                             \
    switch (c) {                                                                                \
    case 0: KERNEL<NCV, 0><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;                       \
    case 1: KERNEL<NCV, (NCV > 1 ? 1 : 0)><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;       \
    case 2: KERNEL<NCV, (NCV > 2 ? 2 : 0)><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;       \
    case 3: KERNEL<NCV, (NCV > 3 ? 3 : 0)><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;       \
    default: KERNEL<NCV, (NCV > 4 ? 4 : 0)><<<grid, kThreads, 0, s>>>(__VA_ARGS__); break;      \
    }
#define MDPP_DISPATCH_NC_C(h, c, KERNEL, grid, s, ...)                                            \
    switch ((h)->cfg.n_cars) {                                                                  \
    case 1: MDPP_DISPATCH_C(1, c, KERNEL, grid, s, __VA_ARGS__) break;                          \
    case 2: MDPP_DISPATCH_C(2, c, KERNEL, grid, s, __VA_ARGS__) break;                          \
    case 3: MDPP_DISPATCH_C(3, c, KERNEL, grid, s, __VA_ARGS__) break;                          \
    case 4: MDPP_DISPATCH_C(4, c, KERNEL, grid, s, __VA_ARGS__) break;                          \
    default: MDPP_DISPATCH_C(5, c, KERNEL, grid, s, __VA_ARGS__) break;                         \
    }

static inline unsigned grid_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }
static inline Control *control_of(mdpp_handle *h) { return (Control *)(h->ws + h->cv.control); }
static inline uint4 *recs_of(mdpp_handle *h) { return (uint4 *)(h->ws + h->cv.recs); }

int mdpp_reset(mdpp_handle *h, const uint8_t *env_mask, const uint8_t *headings, uint32_t seed, int zero_counters,
               void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    MDPP_DISPATCH_NC(h, reset_kernel, grid_for(h->n_envs), s, h->dev, recs_of(h), h->n_envs, h->env_id_base,
                     env_mask, headings, seed, zero_counters);
    MDPP_CUDA(cudaGetLastError());
    if (zero_counters && !env_mask) {
        MDPP_CUDA(cudaMemsetAsync(h->ws + h->cv.dl, 0xFF, (size_t)h->n_envs * 4, s));
        MDPP_CUDA(cudaMemsetAsync(control_of(h)->max_episode, 0, sizeof(uint32_t) * kStatSlots, s));
        h->ep_bound = 0;
        h->ep_device_max_valid = true;
    }
    return MDPP_OK;
}

int mdpp_step_car(mdpp_handle *h, int car_slot, int deadlock_id, int reward_kind, const uint8_t *actions,
                  uint32_t *state_idx, uint32_t *next_idx, float *reward, uint8_t *legal, uint8_t *status,
                  uint8_t *all_done_out, void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (car_slot < 0 || car_slot >= h->cfg.n_cars) return fail(MDPP_EINVAL, "car_slot %d out of range", car_slot);
    if (!actions) return fail(MDPP_EINVAL, "actions is NULL");
    if (deadlock_id != 0 && deadlock_id != 1) return fail(MDPP_EINVAL, "deadlock_id must be 0 or 1");
    if (reward_kind != MDPP_REWARD_Q && reward_kind != MDPP_REWARD_DQN) return fail(MDPP_EINVAL, "unknown reward kind");
    cudaStream_t s = (cudaStream_t)stream;
    MDPP_DISPATCH_NC(h, step_car_kernel, grid_for(h->n_envs), s, h->dev, recs_of(h), h->n_envs, car_slot,
                     deadlock_id, reward_kind, actions, state_idx, next_idx, reward, legal, status, all_done_out,
                     control_of(h));
    MDPP_CUDA(cudaGetLastError());
    return MDPP_OK;
}

int mdpp_step_car_host(mdpp_handle *h, int car_slot, int deadlock_id, int reward_kind, const uint8_t *actions,
                       uint32_t *state_idx, uint32_t *next_idx, float *reward, uint8_t *legal, uint8_t *status,
                       uint8_t *all_done_out, void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (!actions) return fail(MDPP_EINVAL, "actions is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)h->n_envs;
    char *st = h->ws + h->cv.stage; // [state 4n | next 4n | reward 4n | actions n | legal n | status n | done n]
    uint32_t *d_state = (uint32_t *)st, *d_next = (uint32_t *)(st + 4 * n);
    float *d_reward = (float *)(st + 8 * n);
    uint8_t *d_act = (uint8_t *)(st + 12 * n), *d_legal = d_act + n, *d_status = d_act + 2 * n, *d_done = d_act + 3 * n;
    MDPP_CUDA(cudaMemcpyAsync(d_act, actions, n, cudaMemcpyHostToDevice, s));
    int rc = mdpp_step_car(h, car_slot, deadlock_id, reward_kind, d_act, state_idx ? d_state : nullptr,
                           next_idx ? d_next : nullptr, reward ? d_reward : nullptr, legal ? d_legal : nullptr,
                           status ? d_status : nullptr, all_done_out ? d_done : nullptr, stream);
    if (rc) return rc;
    if (state_idx) MDPP_CUDA(cudaMemcpyAsync(state_idx, d_state, 4 * n, cudaMemcpyDeviceToHost, s));
    if (next_idx) MDPP_CUDA(cudaMemcpyAsync(next_idx, d_next, 4 * n, cudaMemcpyDeviceToHost, s));
    if (reward) MDPP_CUDA(cudaMemcpyAsync(reward, d_reward, 4 * n, cudaMemcpyDeviceToHost, s));
    if (legal) MDPP_CUDA(cudaMemcpyAsync(legal, d_legal, n, cudaMemcpyDeviceToHost, s));
    if (status) MDPP_CUDA(cudaMemcpyAsync(status, d_status, n, cudaMemcpyDeviceToHost, s));
    if (all_done_out) MDPP_CUDA(cudaMemcpyAsync(all_done_out, d_done, n, cudaMemcpyDeviceToHost, s));
    MDPP_CUDA(cudaStreamSynchronize(s));
    return MDPP_OK;
}

int mdpp_rollout_random(mdpp_handle *h, int rounds, uint64_t first_round, uint32_t seed, int reward_kind,
                        uint8_t *action, float *reward, uint8_t *done, uint32_t *next_idx, void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (rounds <= 0) return fail(MDPP_EINVAL, "rounds must be positive");
    if (reward_kind != MDPP_REWARD_Q && reward_kind != MDPP_REWARD_DQN) return fail(MDPP_EINVAL, "unknown reward kind");
    cudaStream_t s = (cudaStream_t)stream;
    if (h->fast_rollout) { // table-driven, persistent CTAs (q_round.cuh)
        const RolloutOut out{action, reward, done, next_idx};
        const FastTabs *tabs = (const FastTabs *)(h->ws + h->cv.fast);
        const uint32_t stride = (uint32_t)(h->cfg.n_dest_nodes * h->cfg.n_local_nodes);
        const int64_t tiles = (h->n_envs + 31) / 32;
        const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)h->sm_count * 4, (tiles + 7) / 8));
        const bool lean = !(getenv("MDPP_LEAN_ROLLOUT") && atoi(getenv("MDPP_LEAN_ROLLOUT")) == 0);
#define MDPP_ROLL(NCV)                                                                                           \
        do {                                                                                                     \
            const size_t smem = lean ? sizeof(LeanTabs<NCV>) : sizeof(RoundSmem<NCV>);                           \
            if (!((h->smem_configured >> 27) & 1u)) {                                                            \
                MDPP_CUDA(cudaFuncSetAttribute(rollout_fast_kernel<NCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RoundSmem<NCV>))); \
                MDPP_CUDA(cudaFuncSetAttribute(rollout_lean_kernel<NCV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LeanTabs<NCV>))); \
                MDPP_CUDA(cudaFuncSetAttribute(rollout_lean_kernel<NCV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LeanTabs<NCV>))); \
                h->smem_configured |= 1u << 27;                                                                  \
            }                                                                                                    \
            if (lean && action && reward && done && next_idx)                                                    \
                rollout_lean_kernel<NCV, true><<<grid, kRolloutThreads, smem, s>>>(h->dev, recs_of(h), (uint32_t)h->n_envs, \
                        h->env_id_base, rounds, first_round, seed, reward_kind, out, tabs, stride, control_of(h)); \
            else if (lean)                                                                                       \
                rollout_lean_kernel<NCV, false><<<grid, kRolloutThreads, smem, s>>>(h->dev, recs_of(h), (uint32_t)h->n_envs, \
                        h->env_id_base, rounds, first_round, seed, reward_kind, out, tabs, stride, control_of(h)); \
            else                                                                                                 \
                rollout_fast_kernel<NCV><<<grid, kRolloutThreads, smem, s>>>(h->dev, recs_of(h), (uint32_t)h->n_envs, \
                        h->env_id_base, rounds, first_round, seed, reward_kind, out, tabs, stride, control_of(h)); \
        } while (0)
        switch (h->cfg.n_cars) {
        case 1: MDPP_ROLL(1); break;
        case 2: MDPP_ROLL(2); break;
        case 3: MDPP_ROLL(3); break;
        case 4: MDPP_ROLL(4); break;
        default: MDPP_ROLL(5); break;
        }
#undef MDPP_ROLL
    } else {
        MDPP_DISPATCH_NC(h, rollout_random_kernel, grid_for(h->n_envs), s, h->dev, recs_of(h), h->n_envs,
                         h->env_id_base, rounds, first_round, seed, reward_kind, action, reward, done, next_idx,
                         control_of(h));
    }
    MDPP_CUDA(cudaGetLastError());
    h->ep_bound += rounds; // auto-reset starts at most one new episode per round
    h->ep_device_max_valid = false;
    return MDPP_OK;
}

static unsigned env_ctas(mdpp_handle *h) { // persistent env CTAs (dense mode): one per SM
    const int64_t tiles = (h->n_envs + 31) / 32;
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(h->sm_count, kMaxEnvCtas), tiles));
}

static uint32_t *mark_region(mdpp_handle *h, uint32_t parity, uint32_t region) {
    return (uint32_t *)(h->ws + h->cv.marks) + ((size_t)parity * h->cv.regions + region) * h->cv.cta_slots * h->cv.words_per_cta;
}

static uint32_t *scan_region(mdpp_handle *h, uint32_t bank, uint32_t region) { // scan mode: touched bitmap `region` of `bank`
    return (uint32_t *)(h->ws + h->cv.bits) + ((size_t)bank * h->cv.scan_regions + region) * h->cv.scan_words;
}

static unsigned round_ctas(mdpp_handle *h) { // persistent CTAs of the fused round kernels
    const int64_t tiles = (h->n_envs + 31) / 32;
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>((int64_t)h->sm_count * kRoundCtasPerSm, kMaxRoundCtas), tiles));
}

static Marks marks_of(mdpp_handle *h) {
    Marks m{};
    const size_t keys = (size_t)h->cfg.n_states * MDPP_Q_ROW;
    const uint32_t p = h->parity;
    m.dense = h->cv.dense ? 1u : 0u;
    if (h->cv.dense) {
        m.words_per_cta = h->cv.words_per_cta;
        m.cta_slots = h->cv.cta_slots;
        m.n_ctas = env_ctas(h);
        m.cta_bits = mark_region(h, p, 0);
    } else {
        m.bits = scan_region(h, h->scan_bank, 0);
        m.bits_prev = h->scan_stale;
    }
    if (h->cv.scan) {
        m.scan_words = h->cv.scan_words;
        m.scan_magic = (uint32_t)(((uint64_t)1 << 32) / h->cv.scan_words) + 1u;
        m.l1_check = 1u;
        if (const char *env = getenv("MDPP_L1_CHECK")) m.l1_check = atoi(env) != 0 ? 1u : 0u; // A/B runs
    }
    m.visited = (uint32_t *)(h->ws + h->cv.visited);
    m.sinfo = (const uint32_t *)(h->ws + h->cv.sinfo);
    m.n_keys = (uint32_t)keys;
    m.n_marks = (uint32_t)((size_t)h->cfg.n_states * MDPP_N_ACT);
    return m;
}

// Launch configuration of one learner kernel.  Programmatic dependent launch lets a kernel start while
// its predecessor drains; every learner kernel orders itself with griddepcontrol.wait.  The first
// learner kernel after anything else on the stream (reset, a memset, a caller's kernel) is launched
// with plain stream order: its prologue reads records that predecessor may have written.
struct LearnLaunch {
    cudaLaunchAttribute attr;
    cudaLaunchConfig_t cfg;
    LearnLaunch(mdpp_handle *h, unsigned grid, cudaStream_t s, unsigned block = kThreads, size_t smem = 0) {
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = (h->pdl && h->chain) ? 1 : 0;
        cfg = cudaLaunchConfig_t{};
        cfg.gridDim = dim3(grid ? grid : 1u);
        cfg.blockDim = dim3(block);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
    }
};

// An exchange of the replicas may be running on the handle's side stream (q_sync_launch): whatever reads or writes
// the table next on the caller's stream waits for it first.  Kernels that touch records and marks only (the env
// kernels of pure-exploration rounds) do not call this: they run beside the exchange.
static int sync_join(mdpp_handle *h, cudaStream_t s) {
    if (!h->sync_pending) return MDPP_OK;
    MDPP_CUDA(cudaStreamWaitEvent(s, h->sync_join, 0));
    h->sync_pending = false;
    h->chain = false; // the event wait sits between this launch and its predecessor
    return MDPP_OK;
}

static size_t step_tabs_bytes(int n_cars) {
    switch (n_cars) {
    case 1: return sizeof(StepTabs<1>);
    case 2: return sizeof(StepTabs<2>);
    case 3: return sizeof(StepTabs<3>);
    case 4: return sizeof(StepTabs<4>);
    default: return sizeof(StepTabs<5>);
    }
}

static int q_env_launch(mdpp_handle *h, const float *q, const mdpp_learner *lp, int car_slot, uint64_t round,
                        const uint8_t *forced, uint8_t *action_out, uint32_t *state_idx, float *reward,
                        cudaStream_t s) {
    if (int rc = sync_join(h, s)) return rc; // a sub-step kernel may read the table (greedy lanes)
    // Philox hand-over: the car-0 launch of a round publishes the decision words of cars 1..3
    int rng_mode = 0;
    const uint32_t order_slot = (uint32_t)car_slot;
    if (lp->flags & MDPP_LEARN_RANDOM_ORDER) {
        car_slot = 0; // every env draws its own car: the slot-0 kernels carry that dispatch
    } else if (h->cfg.n_cars > 1) {
        if (car_slot == 0) {
            rng_mode = 1;
            h->rng_round = round;
            h->rng_seed = lp->seed;
        } else if (car_slot < 4 && h->rng_round == round && h->rng_seed == lp->seed) {
            rng_mode = 2;
        }
    }
    const bool plain = !forced && !action_out && !state_idx && !reward && lp->flags == 0;
    EnvArgs ea{};
    ea.recs = recs_of(h);
    ea.n_envs = h->n_envs;
    ea.env_id_base = h->env_id_base;
    ea.round = round;
    ea.q = q;
    ea.tabs = (const FastTabs *)(h->ws + h->cv.fast);
    ea.rank_stride = (uint32_t)(h->cfg.n_dest_nodes * h->cfg.n_local_nodes);
    ea.n_states = (uint32_t)h->cfg.n_states;
    ea.ctl = control_of(h);
    ea.rng_words = (uint32_t *)(h->ws + h->cv.rng);
    ea.rng_mode = rng_mode;
    ea.order_slot = order_slot;
    ea.forced = forced;
    ea.action_out = action_out;
    ea.state_out = state_idx;
    ea.reward_out = reward;
    ea.dl_state = (uint32_t *)(h->ws + h->cv.dl);
    const Marks mk = marks_of(h);
    const bool bitmap = (size_t)mk.n_marks > kMaxByteMapMarks;
    const size_t map_bytes = bitmap ? (size_t)mk.words_per_cta * 4 : (((size_t)mk.n_marks + 15) / 16) * 16;
    const size_t tab_bytes = (step_tabs_bytes(h->cfg.n_cars) + 15) & ~(size_t)15;
    const size_t smem_bytes = h->cv.dense ? tab_bytes + map_bytes : 0; // scan mode reads the tables in place
    LearnLaunch ll = h->cv.dense ? LearnLaunch(h, env_ctas(h), s, kEnvSmemThreads, smem_bytes)
                                 : LearnLaunch(h, grid_for(h->n_envs), s);
    ll.cfg.attrs = &ll.attr; // the struct was copied: re-point at this copy's attribute
    const int variant = car_slot * 2 + (plain ? 1 : 0);
#define MDPP_QENV_K(KERNEL)                                                                                      \
    do {                                                                                                         \
        if (h->cv.dense && !((h->smem_configured >> variant) & 1u)) {                                            \
            MDPP_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)); \
            h->smem_configured |= 1u << variant;                                                                 \
        }                                                                                                        \
        MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, KERNEL, h->dev, *lp, ea, mk, control_of(h)));                       \
    } while (0)
#define MDPP_QENV_C(NCV, CV)                                                                                     \
    if (h->cv.dense) {                                                                                           \
        if (bitmap) {                                                                                            \
            if (plain) MDPP_QENV_K((q_env_smem_kernel<NCV, CV, true, true>));                                    \
            else MDPP_QENV_K((q_env_smem_kernel<NCV, CV, false, true>));                                         \
        } else {                                                                                                 \
            if (plain) MDPP_QENV_K((q_env_smem_kernel<NCV, CV, true, false>));                                   \
            else MDPP_QENV_K((q_env_smem_kernel<NCV, CV, false, false>));                                        \
        }                                                                                                        \
    } else {                                                                                                     \
        if (plain) MDPP_QENV_K((q_env_kernel<NCV, CV, true>));                                                   \
        else MDPP_QENV_K((q_env_kernel<NCV, CV, false>));                                                        \
    }
#define MDPP_QENV(NCV)                                                                                           \
    switch (car_slot) {                                                                                          \
    case 0: MDPP_QENV_C(NCV, 0) break;                                                                           \
    case 1: if constexpr (NCV > 1) { MDPP_QENV_C(NCV, (NCV > 1 ? 1 : 0)) } break;                                \
    case 2: if constexpr (NCV > 2) { MDPP_QENV_C(NCV, (NCV > 2 ? 2 : 0)) } break;                                \
    case 3: if constexpr (NCV > 3) { MDPP_QENV_C(NCV, (NCV > 3 ? 3 : 0)) } break;                                \
    default: if constexpr (NCV > 4) { MDPP_QENV_C(NCV, (NCV > 4 ? 4 : 0)) } break;                               \
    }
    switch (h->cfg.n_cars) {
    case 1: MDPP_QENV(1) break;
    case 2: MDPP_QENV(2) break;
    case 3: MDPP_QENV(3) break;
    case 4: MDPP_QENV(4) break;
    default: MDPP_QENV(5) break;
    }
#undef MDPP_QENV_K
#undef MDPP_QENV_C
#undef MDPP_QENV
    // the end-of-round reset starts at most one new episode; in random car order any sub-step may end one
    if (car_slot == h->cfg.n_cars - 1 || (lp->flags & MDPP_LEARN_RANDOM_ORDER)) h->ep_bound += 1;
    if (lp->flags & MDPP_LEARN_DETECT) { // a flagged env jumps to the end of its episode budget
        h->ep_bound = (int64_t)1 << 40;
        h->ep_device_max_valid = false;
    }
    h->chain = true;
    return MDPP_OK;
}

// qsrc -> qdst.  dense: the whole table, marked keys substituted.  scan: the keys marked in this and in the previous
// sub-step (q_pull_scan_kernel); the caller swaps the tables afterwards in both modes.
static unsigned scan_grid(const mdpp_handle *h) {
    return (unsigned)(((int64_t)h->cv.scan_words + kThreads - 1) / kThreads); // a bitmap word per thread
}
static int q_scan_pull_launch(mdpp_handle *h, float *qsrc, float *qdst, const mdpp_learner *lp, cudaStream_t s, const Marks &mk) {
    LearnLaunch ll(h, scan_grid(h), s);
    MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_pull_scan_kernel, (const float *)qsrc, qdst, mk, lp->alpha, lp->discount, control_of(h)));
    h->scan_stale = mk.bits; // qsrc is now one update behind on these keys
    h->chain = true;
    return MDPP_OK;
}
static int q_pull_launch(mdpp_handle *h, float *qsrc, float *qdst, const mdpp_learner *lp, cudaStream_t s,
                         const Marks *marks = nullptr, int early_trigger = 0) {
    const Marks mk = marks ? *marks : marks_of(h);
    if (int rc = sync_join(h, s)) return rc;
    if (h->cv.dense) {
        LearnLaunch ll(h, grid_for(mk.n_marks), s);
        MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_pull_dense_kernel, (const float *)qsrc, qdst, mk, lp->alpha,
                                     lp->discount, early_trigger, control_of(h)));
        h->mirror_ok = false; // (mid-size tables) qsrc is behind on keys no bitmap of the scan mode describes
        if (!marks) h->parity ^= 1u;
    } else {
        if (int rc = q_scan_pull_launch(h, qsrc, qdst, lp, s, mk)) return rc;
        if (!marks) h->scan_bank = (h->scan_bank + 1u) % 3u;
    }
    h->chain = true;
    return MDPP_OK;
}

// scan mode, before its first pull: `alt` must equal `cur` except on the keys of h->scan_stale.  The pair (caller's
// table, workspace copy) stays like that from one learner call to the next (scan_settle) unless the caller has
// written the table since (mdpp_q_table_written), passes another table, or a dense pull ran in between (mid-size
// tables): then `cur` is copied once.
static int scan_begin(mdpp_handle *h, const float *q, const float *cur, float *alt, cudaStream_t s) {
    if (!h->cv.scan || (h->mirror_ok && h->mirror_q == q)) return MDPP_OK;
    MDPP_CUDA(cudaMemcpyAsync(alt, cur, (size_t)h->cfg.n_states * MDPP_Q_ROW * 4, cudaMemcpyDeviceToDevice, s));
    h->mirror_ok = true;
    h->mirror_q = q;
    h->scan_stale = nullptr; // (never set here: every path that leaves the scan mode settles first)
    h->chain = false;
    return MDPP_OK;
}
// scan mode: `cur` holds the table; the other copy is behind on the keys of the last sub-step's bitmap.  One
// copy-only pass makes both complete (and clears that bitmap), after which either may be taken as the table: the
// caller's pointers are reset to (q, workspace copy).
static int scan_settle(mdpp_handle *h, float *q, float *&cur, float *&alt, cudaStream_t s) {
    if (!h->scan_stale) return MDPP_OK;
    Marks mk = marks_of(h);
    mk.bits = nullptr;
    mk.bits_prev = h->scan_stale;
    LearnLaunch ll(h, scan_grid(h), s);
    MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_pull_scan_kernel, (const float *)cur, alt, mk, 0.f, 0.f, control_of(h)));
    h->scan_stale = nullptr;
    h->chain = true;
    cur = q;
    alt = (float *)(h->ws + h->cv.qalt);
    return MDPP_OK;
}

// ---- fused round (q_round.cuh): 1- and 2-car tables whose byte maps fit shared memory ------------------
static size_t round_smem_bytes(int n_cars, size_t map_bytes, int maps) {
    const size_t tabs = n_cars == 1 ? sizeof(RoundSmem<1>) : sizeof(RoundSmem<2>);
    return ((tabs + 15) & ~(size_t)15) + (size_t)maps * map_bytes;
}
static bool round_eligible(const mdpp_handle *h) {
    if (!h->cv.dense || h->cfg.n_cars > 2) return false;
    const size_t marks = (size_t)h->cfg.n_states * MDPP_N_ACT;
    if (marks > kMaxByteMapMarks) return false;
    const size_t map_bytes = ((marks + 15) / 16) * 16;
    const size_t need = h->cfg.n_cars == 1 ? round_smem_bytes(1, map_bytes, 1) : round_smem_bytes(2, map_bytes, 2);
    if (getenv("MDPP_FUSED") && atoi(getenv("MDPP_FUSED")) == 0) return false;
    return need <= (size_t)220 * 1024;
}

static int q_round_launch(mdpp_handle *h, float *&cur, float *&alt, const mdpp_learner *lp, uint64_t round, cudaStream_t s) {
    const uint32_t P = h->round_parity;
    h->round_parity ^= 1u;
    // every env still in an epsilon = 1 segment: every coin says "explore", no lane can be parked
    const bool pure = lp->eps_q16[0] >= 65536u && h->ep_bound < (int64_t)lp->eps_threshold[0];
    h->ep_bound += 1;
    RoundArgs ra{};
    ra.recs = recs_of(h);
    ra.n_envs = (uint32_t)h->n_envs;
    ra.env_id_base = h->env_id_base;
    ra.round = round;
    ra.rng_words = (uint32_t *)(h->ws + h->cv.rng);
    ra.tabs = (const FastTabs *)(h->ws + h->cv.fast);
    ra.defer_flags = (uint32_t *)(h->ws + h->cv.flags) + (size_t)P * h->cv.cta_slots;
    ra.words_per_cta = h->cv.words_per_cta;
    ra.cta_slots = h->cv.cta_slots;
    ra.map_bytes = (uint32_t)((((size_t)h->cfg.n_states * MDPP_N_ACT + 15) / 16) * 16);
    ra.rank_stride = (uint32_t)(h->cfg.n_dest_nodes * h->cfg.n_local_nodes);
    ra.n_states = (uint32_t)h->cfg.n_states;
    const unsigned grid = round_ctas(h);
    ra.tiles_q = (uint32_t)(((h->n_envs + 31) / 32) / grid);
    ra.tiles_r = (uint32_t)(((h->n_envs + 31) / 32) % grid);
    Marks mk = marks_of(h);
    mk.n_ctas = grid;
    if (!pure) // greedy lanes of the round kernel read the table: an exchange in flight comes first
        if (int rc = sync_join(h, s)) return rc;
#define MDPP_ROUND_K(KERNEL, SMEM, BIT)                                                                          \
    do {                                                                                                         \
        if (!((h->smem_configured >> (BIT)) & 1u)) {                                                             \
            MDPP_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)));    \
            h->smem_configured |= 1u << (BIT);                                                                   \
        }                                                                                                        \
        LearnLaunch ll(h, grid, s, kRoundThreads, (SMEM));                                                       \
        MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, KERNEL, h->dev, *lp, ra, control_of(h)));                           \
        h->chain = true;                                                                                         \
    } while (0)
    if (h->cfg.n_cars == 1) {
        ra.q = cur;
        ra.bits_a = mark_region(h, P, 0);
        MDPP_ROUND_K((q_round_kernel<1, 0>), round_smem_bytes(1, ra.map_bytes, 1), 28);
        mk.cta_bits = mark_region(h, P, 0);
        int rc = q_pull_launch(h, cur, alt, lp, s, &mk);
        if (rc) return rc;
        std::swap(cur, alt);
        return MDPP_OK;
    }
    // K0: car 0 of every env, car 1 where it explores
    ra.q = cur;
    ra.bits_a = mark_region(h, P, 0);
    ra.bits_b = mark_region(h, P, 1);
    MDPP_ROUND_K((q_round_kernel<2, 0>), round_smem_bytes(2, ra.map_bytes, 2), 29);
    mk.cta_bits = mark_region(h, P, 0);
    int rc = q_pull_launch(h, cur, alt, lp, s, &mk, 1); // followed by K1 or the second pull: both wait first
    if (rc) return rc;
    std::swap(cur, alt);
    mk.cta_bits = mark_region(h, P, 1);
    if (!pure) { // K1: the lanes whose car 1 acts greedily, against the table after the pull of sub-step 0
        ra.q = cur;
        ra.bits_a = mark_region(h, P, 2);
        ra.bits_b = nullptr;
        MDPP_ROUND_K((q_round_kernel<2, 1>), round_smem_bytes(2, ra.map_bytes, 1), 30);
        mk.cta_bits2 = mark_region(h, P, 2);
        mk.flags2 = ra.defer_flags;
    }
    rc = q_pull_launch(h, cur, alt, lp, s, &mk);
    if (rc) return rc;
    std::swap(cur, alt);
#undef MDPP_ROUND_K
    return MDPP_OK;
}

// n pure-exploration rounds in one launch (q_rounds_pure_kernel), then the pulls of their 2n (n) sub-steps in order.
// Returns the number of rounds it ran: 0 when the next n rounds are not certainly inside the epsilon = 1 segment.
static int pure_rounds_fit(const mdpp_handle *h, const mdpp_learner *lp, int want) {
    if (!h->multi_round || lp->eps_q16[0] < 65536u) return 0;
    const int64_t room = (int64_t)lp->eps_threshold[0] - h->ep_bound; // rounds before the bound leaves the segment
    int n = (int)std::min<int64_t>(std::min<int64_t>(want, kMaxPureRounds), room);
    const size_t map_bytes = (((size_t)h->cfg.n_states * MDPP_N_ACT + 15) / 16) * 16;
    const size_t tabs = ((h->cfg.n_cars == 1 ? sizeof(RoundSmem<1>) : sizeof(RoundSmem<2>)) + 15) & ~(size_t)15;
    if (tabs + 2 * (size_t)h->cfg.n_cars * map_bytes > (size_t)220 * 1024) return 0;
    return n >= 2 ? n : 0;
}
static int q_pure_rounds_launch(mdpp_handle *h, float *&cur, float *&alt, const mdpp_learner *lp, uint64_t round, int n,
                                cudaStream_t s, bool env_only = false) {
    const uint32_t P = h->round_parity;
    h->round_parity ^= 1u;
    h->ep_bound += n;
    const int nc = h->cfg.n_cars;
    RoundArgs ra{};
    ra.recs = recs_of(h);
    ra.n_envs = (uint32_t)h->n_envs;
    ra.env_id_base = h->env_id_base;
    ra.round = round;
    ra.q = cur;
    ra.rng_words = (uint32_t *)(h->ws + h->cv.rng);
    ra.tabs = (const FastTabs *)(h->ws + h->cv.fast);
    ra.words_per_cta = h->cv.words_per_cta;
    ra.cta_slots = h->cv.cta_slots;
    ra.map_bytes = (uint32_t)((((size_t)h->cfg.n_states * MDPP_N_ACT + 15) / 16) * 16);
    ra.rank_stride = (uint32_t)(h->cfg.n_dest_nodes * h->cfg.n_local_nodes);
    ra.n_states = (uint32_t)h->cfg.n_states;
    const unsigned grid = round_ctas(h);
    ra.tiles_q = (uint32_t)(((h->n_envs + 31) / 32) / grid);
    ra.tiles_r = (uint32_t)(((h->n_envs + 31) / 32) % grid);
    ra.n_rounds = (uint32_t)n;
    ra.region_words = h->cv.cta_slots * h->cv.words_per_cta;
    ra.bits_a = mark_region(h, P, 0);
    const size_t tabs = ((nc == 1 ? sizeof(RoundSmem<1>) : sizeof(RoundSmem<2>)) + 15) & ~(size_t)15;
    const size_t smem = tabs + 2 * (size_t)nc * ra.map_bytes;
    {
        LearnLaunch ll(h, grid, s, kRoundThreads, smem);
        if (nc == 1) {
            if (!((h->smem_configured >> 25) & 1u)) {
                MDPP_CUDA(cudaFuncSetAttribute(q_rounds_pure_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                h->smem_configured |= 1u << 25;
            }
            MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_rounds_pure_kernel<1>, h->dev, *lp, ra, control_of(h)));
        } else {
            if (!((h->smem_configured >> 26) & 1u)) {
                MDPP_CUDA(cudaFuncSetAttribute(q_rounds_pure_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                h->smem_configured |= 1u << 26;
            }
            MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_rounds_pure_kernel<2>, h->dev, *lp, ra, control_of(h)));
        }
        h->chain = true;
    }
    if (env_only) return MDPP_OK; // profiling hook: the marks are never pulled
    Marks mk = marks_of(h);
    mk.n_ctas = grid;
    // the pulls of all n * nc sub-steps in two launches when one cluster can hold the table (q_round.cuh)
    if (h->multi_pull && (size_t)h->cfg.n_states <= (size_t)kMultiPullThreads * kMultiPullCtas) {
        const uint32_t n_sub = (uint32_t)(n * nc), words = h->cv.words_per_cta;
        uint32_t *gbits = (uint32_t *)(h->ws + h->cv.gbits);
        {
            const unsigned items = n_sub * words;
            LearnLaunch ll(h, (items + kThreads / 32 - 1) / (kThreads / 32), s);
            MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_marks_reduce_kernel, (const uint32_t *)mark_region(h, P, 0),
                                              ra.region_words, n_sub, words, words, h->cv.cta_slots, (uint32_t)grid, gbits));
        }
        if (int rc = sync_join(h, s)) return rc; // the cluster pull reads the table (the reduce above only marks)
        MultiPullArgs pa{};
        pa.q = cur;
        pa.sinfo = mk.sinfo;
        pa.visited = mk.visited;
        pa.gbits = gbits;
        pa.n_substeps = n_sub;
        pa.words = words;
        pa.n_states = (uint32_t)h->cfg.n_states;
        pa.alpha = lp->alpha;
        pa.discount = lp->discount;
        if (!((h->smem_configured >> 24) & 1u)) {
            MDPP_CUDA(cudaFuncSetAttribute(q_multi_pull_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            h->smem_configured |= 1u << 24;
        }
        cudaLaunchAttribute attrs[2];
        attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[0].val.programmaticStreamSerializationAllowed = (h->pdl && h->chain) ? 1 : 0;
        attrs[1].id = cudaLaunchAttributeClusterDimension;
        attrs[1].val.clusterDim.x = kMultiPullCtas;
        attrs[1].val.clusterDim.y = 1;
        attrs[1].val.clusterDim.z = 1;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(kMultiPullCtas);
        cfg.blockDim = dim3(kMultiPullThreads);
        cfg.stream = s;
        cfg.attrs = attrs;
        cfg.numAttrs = 2;
        MDPP_LAUNCH(h, cudaLaunchKernelEx(&cfg, q_multi_pull_kernel, pa, control_of(h)));
        return MDPP_OK;
    }
    for (int i = 0; i < n * nc; i++) {
        mk.cta_bits = mark_region(h, P, (uint32_t)i);
        // every pull but the last is followed by another pull, which waits before it touches anything
        int rc = q_pull_launch(h, cur, alt, lp, s, &mk, i + 1 < n * nc ? 1 : 0);
        if (rc) return rc;
        std::swap(cur, alt);
    }
    return MDPP_OK;
}

// scan mode (large tables): n pure-exploration rounds, all cars, in one env pass (q_rounds_pure_scan_kernel), then the
// pulls of their n * n_cars sub-steps in order.  0 when the next rounds are not certainly inside the epsilon = 1 segment.
static int pure_scan_rounds_fit(const mdpp_handle *h, const mdpp_learner *lp, int want) {
    if (!h->cv.scan || !h->multi_round || !h->pure_scan || lp->eps_q16[0] < 65536u) return 0;
    const int64_t room = (int64_t)lp->eps_threshold[0] - h->ep_bound; // rounds before the bound leaves the segment
    const int cap = (int)(h->cv.scan_regions / (uint32_t)h->cfg.n_cars);
    const int n = (int)std::min<int64_t>(std::min<int64_t>(want, cap), room);
    return n >= 1 ? n : 0;
}
extern "C++" {
template <int NC>
static int pure_scan_kernel_launch(mdpp_handle *h, const mdpp_learner *lp, const PureScanArgs &pa, cudaStream_t s, bool beside) {
    // beside: pulls run next to this kernel.  Four of its CTAs (256 threads x 64 registers each) would take every
    // register of an SM for the whole launch, so the shared-memory request is padded until only three fit, which
    // leaves registers and shared memory for two CTAs of q_pull_scan_kernel per SM.
    const size_t pad = (size_t)58 * 1024, smem = beside ? std::max(sizeof(LeanTabs<NC>), pad) : sizeof(LeanTabs<NC>);
    if (!((h->smem_configured >> 31) & 1u)) {
        MDPP_CUDA(cudaFuncSetAttribute(q_rounds_pure_scan_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)std::max(sizeof(LeanTabs<NC>), pad)));
        // the largest shared-memory carve-out for both kernels: the pull's CTAs must fit the SM configuration the
        // env kernel runs under (an SM does not change its carve-out while CTAs are resident)
        MDPP_CUDA(cudaFuncSetAttribute(q_rounds_pure_scan_kernel<NC>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
        MDPP_CUDA(cudaFuncSetAttribute(q_pull_scan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
        h->smem_configured |= 1u << 31;
    }
    const int64_t tiles = (h->n_envs + 31) / 32;
    const int per_sm = (beside || sizeof(LeanTabs<NC>) > pad) ? 3 : 4;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)h->sm_count * per_sm, (tiles + 7) / 8));
    LearnLaunch ll(h, grid, s, kRolloutThreads, smem);
    MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_rounds_pure_scan_kernel<NC>, h->dev, *lp, pa, control_of(h)));
    h->chain = true;
    return MDPP_OK;
}
} // extern "C++"
// The caller's stream waits for the pulls in flight on the handle's pull stream (no-op when there are none).
static int overlap_join(mdpp_handle *h, cudaStream_t s) {
    if (!h->pulls_pending) return MDPP_OK;
    MDPP_CUDA(cudaStreamWaitEvent(s, h->ev_pull[(h->ov_groups - 1u) & 1u], 0));
    h->pulls_pending = false;
    h->ov_groups = 0;
    h->chain = false;
    return MDPP_OK;
}
static int q_pure_scan_rounds_launch(mdpp_handle *h, float *q, float *&cur, float *&alt, const mdpp_learner *lp, uint64_t round,
                                     int n, cudaStream_t s) {
    const int nc = h->cfg.n_cars;
    if (int rc = scan_begin(h, q, cur, alt, s)) return rc; // (mid-size tables after a dense pull: one copy)
    // While every env explores, the env pass of a group of rounds reads no table: the pulls of group g run on a stream
    // of the handle's own BESIDE the env pass of group g + 1.  Env pass g marks bank g % 3; pulls g read it, and clear
    // the last bitmap of bank (g - 1) % 3 first; env pass g + 1 marks bank (g + 1) % 3, last read by the pulls of group
    // g - 2 -- which env pass g + 1 therefore waits for (ev_pull), as pulls g wait for env pass g (ev_env).
    const bool beside = h->scan_overlap && h->sync_world <= 1;
    cudaStream_t ps = s;
    uint32_t g = 0;
    if (beside) {
        if (!h->pull_stream) {
            int lo = 0, hi = 0;
            MDPP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            MDPP_CUDA(cudaStreamCreateWithPriority(&h->pull_stream, cudaStreamNonBlocking, hi)); // its CTAs first when an SM has room
            for (int i = 0; i < 2; i++) {
                MDPP_CUDA(cudaEventCreateWithFlags(&h->ev_env[i], cudaEventDisableTiming));
                MDPP_CUDA(cudaEventCreateWithFlags(&h->ev_pull[i], cudaEventDisableTiming));
            }
        }
        ps = h->pull_stream;
        g = h->ov_groups++;
        if (g >= 2) MDPP_CUDA(cudaStreamWaitEvent(s, h->ev_pull[g & 1u], 0)); // the pulls of group g - 2
        h->chain = false; // plain stream order behind the previous env pass (whose records this one reads)
    }
    const uint32_t bank = h->scan_bank;
    h->scan_bank = (h->scan_bank + 1u) % 3u;
    h->ep_bound += n;
    PureScanArgs pa{};
    pa.recs = recs_of(h);
    pa.n_envs = (uint32_t)h->n_envs;
    pa.env_id_base = h->env_id_base;
    pa.rounds = (uint32_t)n;
    pa.first_round = round;
    pa.tabs = (const FastTabs *)(h->ws + h->cv.fast);
    pa.rank_stride = (uint32_t)(h->cfg.n_dest_nodes * h->cfg.n_local_nodes);
    pa.n_states = (uint32_t)h->cfg.n_states;
    pa.bits0 = scan_region(h, bank, 0);
    pa.region_words = h->cv.scan_words;
    pa.scan_magic = (uint32_t)(((uint64_t)1 << 32) / h->cv.scan_words) + 1u;
    Marks mk = marks_of(h);
    pa.l1_check = mk.l1_check;
    int rc;
    switch (nc) { // (the env pass touches records and marks only: an exchange in flight need not be joined first)
    case 1: rc = pure_scan_kernel_launch<1>(h, lp, pa, s, beside); break;
    case 2: rc = pure_scan_kernel_launch<2>(h, lp, pa, s, beside); break;
    case 3: rc = pure_scan_kernel_launch<3>(h, lp, pa, s, beside); break;
    case 4: rc = pure_scan_kernel_launch<4>(h, lp, pa, s, beside); break;
    default: rc = pure_scan_kernel_launch<5>(h, lp, pa, s, beside); break;
    }
    if (rc) return rc;
    if (beside) {
        MDPP_CUDA(cudaEventRecord(h->ev_env[g & 1u], s));
        MDPP_CUDA(cudaStreamWaitEvent(ps, h->ev_env[g & 1u], 0));
        h->chain = false; // the first pull of the group follows an event wait
    }
    for (int i = 0; i < n * nc; i++) {
        mk.bits = scan_region(h, bank, (uint32_t)i);
        mk.bits_prev = h->scan_stale;
        if ((rc = sync_join(h, ps))) return rc;
        if ((rc = q_scan_pull_launch(h, cur, alt, lp, ps, mk))) return rc;
        std::swap(cur, alt);
    }
    if (beside) {
        MDPP_CUDA(cudaEventRecord(h->ev_pull[g & 1u], ps));
        h->pulls_pending = true;
        h->chain = false;
    }
    return MDPP_OK;
}

// n rounds WITH greedy decisions in one cooperative launch (q_rounds_mixed_kernel).  The table ends where it began
// for 2 cars (an even number of pulls); for 1 car and an odd n it ends in the other copy.
static size_t mixed_smem_bytes(const mdpp_handle *h) {
    const size_t tabs = ((h->cfg.n_cars == 1 ? sizeof(RoundSmem<1>) : sizeof(RoundSmem<2>)) + 15) & ~(size_t)15;
    return tabs + (((size_t)h->cfg.n_states * MDPP_N_ACT + 15) / 16) * 16;
}
static int mixed_rounds_fit(mdpp_handle *h, int want) {
    if (!h->mixed_rounds || want < 1) return 0;
    if (h->coop_ok < 0) { // first use: cooperative launch available, and the whole grid co-resident?
        h->coop_ok = 0;
        int dev = 0, coop = 0, per_sm = 0;
        const void *fn = h->cfg.n_cars == 1 ? (const void *)q_rounds_mixed_kernel<1> : (const void *)q_rounds_mixed_kernel<2>;
        const size_t smem = mixed_smem_bytes(h);
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess &&
            coop && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kRoundThreads, smem) == cudaSuccess &&
            (int64_t)per_sm * h->sm_count >= (int64_t)round_ctas(h))
            h->coop_ok = 1;
        else
            cudaGetLastError(); // clear: the sub-step kernels take over
    }
    return h->coop_ok == 1 ? std::min(want, 16) : 0;
}
static int q_mixed_rounds_launch(mdpp_handle *h, float *&cur, float *&alt, const mdpp_learner *lp, uint64_t round, int n,
                                 cudaStream_t s) {
    if (int rc = sync_join(h, s)) return rc;
    const int nc = h->cfg.n_cars;
    h->ep_bound += n;
    RoundArgs ra{};
    ra.recs = recs_of(h);
    ra.n_envs = (uint32_t)h->n_envs;
    ra.env_id_base = h->env_id_base;
    ra.round = round;
    ra.rng_words = (uint32_t *)(h->ws + h->cv.rng);
    ra.tabs = (const FastTabs *)(h->ws + h->cv.fast);
    ra.words_per_cta = h->cv.words_per_cta;
    ra.cta_slots = h->cv.cta_slots;
    ra.map_bytes = (uint32_t)((((size_t)h->cfg.n_states * MDPP_N_ACT + 15) / 16) * 16);
    ra.rank_stride = (uint32_t)(h->cfg.n_dest_nodes * h->cfg.n_local_nodes);
    ra.n_states = (uint32_t)h->cfg.n_states;
    const unsigned grid = round_ctas(h);
    ra.tiles_q = (uint32_t)(((h->n_envs + 31) / 32) / grid);
    ra.tiles_r = (uint32_t)(((h->n_envs + 31) / 32) % grid);
    ra.n_rounds = (uint32_t)n;
    ra.bits_a = mark_region(h, 0, 0);
    MixedArgs mx{};
    mx.barrier = (uint32_t *)(h->ws + h->cv.barrier);
    mx.q0 = cur;
    mx.q1 = alt;
    const Marks mk = marks_of(h);
    mx.sinfo = mk.sinfo;
    mx.visited = mk.visited;
    mx.n_marks = (uint32_t)((size_t)h->cfg.n_states * MDPP_N_ACT);
    mx.n_ctas = grid;
    mx.alpha = lp->alpha;
    mx.discount = lp->discount;
    const size_t smem = mixed_smem_bytes(h);
    mdpp_learner lpv = *lp;
    Control *ctl = control_of(h);
    void *args[] = {(void *)&h->dev, (void *)&lpv, (void *)&ra, (void *)&mx, (void *)&ctl};
    const void *fn = nc == 1 ? (const void *)q_rounds_mixed_kernel<1> : (const void *)q_rounds_mixed_kernel<2>;
    MDPP_LAUNCH(h, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kRoundThreads), args, smem, s));
    h->chain = false; // launched in plain stream order; what follows must not overlap it either
    if ((n * nc) & 1) std::swap(cur, alt);
    return MDPP_OK;
}

static int q_copy_launch(mdpp_handle *h, const float *src, float *dst, cudaStream_t s) {
    if (int rc = sync_join(h, s)) return rc;
    const uint32_t n4 = (uint32_t)((size_t)h->cfg.n_states * MDPP_Q_ROW / 4);
    q_copy_kernel<<<grid_for(n4), kThreads, 0, s>>>((const float4 *)src, (float4 *)dst, n4);
    MDPP_CUDA(cudaGetLastError());
    h->launches += 1;
    h->chain = false;
    return MDPP_OK;
}

// ---- replica averaging (q_sync.cuh) -----------------------------------------------------------------------
static size_t sync_table_bytes(const mdpp_config &c) { return align_up((size_t)c.n_states * MDPP_Q_ROW * 4); }
static bool sync_due(const mdpp_handle *h, uint64_t round) {
    return h->sync_world > 1 && h->sync_interval > 0 && round > 0 && round % (uint64_t)h->sync_interval == 0;
}
// rounds that may share a launch from `round` on without crossing an averaging point
static int sync_room(const mdpp_handle *h, uint64_t round, int want) {
    if (h->sync_world <= 1 || h->sync_interval <= 0) return want;
    const uint64_t left = (uint64_t)h->sync_interval - round % (uint64_t)h->sync_interval;
    return (int)std::min<uint64_t>((uint64_t)want, left);
}
// chunks per CTA: one while the table fits 148 CTAs that way (every float4 one NVLink round trip), up to kSyncPer
static uint32_t sync_per(const mdpp_handle *h) {
    const uint32_t n4 = (uint32_t)((size_t)h->cfg.n_states * MDPP_Q_ROW / 4);
    const uint32_t per = (n4 + kSyncThreads * kSyncMaxCtas - 1) / (kSyncThreads * kSyncMaxCtas);
    return std::max(1u, std::min(per, (uint32_t)kSyncPer));
}
static unsigned sync_ctas(const mdpp_handle *h) {
    const uint32_t n4 = (uint32_t)((size_t)h->cfg.n_states * MDPP_Q_ROW / 4), chunk = kSyncThreads * sync_per(h);
    return (n4 + chunk - 1) / chunk;
}
static SyncArgs sync_args(mdpp_handle *h, uint32_t epoch) {
    SyncArgs a{};
    const size_t tb = sync_table_bytes(h->cfg);
    for (int r = 0; r < h->sync_world; r++) {
        a.table[r] = (float *)h->sync_bufs[r];
        a.flags[r] = (uint32_t *)((char *)h->sync_bufs[r] + tb);
    }
    a.rank = (uint32_t)h->sync_rank;
    a.world = (uint32_t)h->sync_world;
    a.epoch = epoch;
    a.n4 = (uint32_t)((size_t)h->cfg.n_states * MDPP_Q_ROW / 4);
    a.per = sync_per(h);
    a.timeout_ns = 2000000000ull;
    a.errors = (unsigned long long *)&control_of(h)->stats[0].internal_errors;
    a.mode = getenv("MDPP_SYNC_MODE") ? (uint32_t)atoi(getenv("MDPP_SYNC_MODE")) : 0u;
    return a;
}
// beside: the exchange goes to the handle's side stream, forked from and later joined into the caller's stream
// (sync_join), so that the env kernel launched next on the caller's stream runs beside it.  As a member of the
// caller's stream -- even with programmatic dependent launch on both sides -- the env kernel did not start before
// the exchange had finished (one GPU, no peer to wait for: 31.1 us per round at interval 1 against 21.9 with an
// empty exchange kernel and 20.1 without any).
static int q_sync_launch(mdpp_handle *h, float *q, cudaStream_t s, bool beside = false) {
    if (h->sync_world <= 1) return MDPP_OK;
    if ((void *)q != h->sync_bufs[h->sync_rank])
        return fail(MDPP_EINVAL, "replica averaging: q must be the table of this rank's exchange buffer");
    if (int rc = sync_join(h, s)) return rc; // (never pending here: every round's pulls join first)
    const SyncArgs a = sync_args(h, ++h->sync_epoch);
    if (getenv("MDPP_SYNC_BESIDE") && atoi(getenv("MDPP_SYNC_BESIDE")) == 0) beside = false; // A/B runs
    cudaStream_t main_stream = s;
    if (beside) {
        if (!h->sync_stream) {
            MDPP_CUDA(cudaStreamCreateWithFlags(&h->sync_stream, cudaStreamNonBlocking));
            MDPP_CUDA(cudaEventCreateWithFlags(&h->sync_fork, cudaEventDisableTiming));
            MDPP_CUDA(cudaEventCreateWithFlags(&h->sync_join, cudaEventDisableTiming));
        }
        MDPP_CUDA(cudaEventRecord(h->sync_fork, main_stream)); // after the last pull: the own table is final
        MDPP_CUDA(cudaStreamWaitEvent(h->sync_stream, h->sync_fork, 0));
        s = h->sync_stream;
    }
    const bool saved_chain = h->chain;
    const bool no_pdl = beside; // plain launch on the side stream
    if (no_pdl) h->chain = false;
    LearnLaunch ll(h, sync_ctas(h), s, kSyncThreads);
    switch (h->sync_world) {
    case 2: MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_sync_kernel<2>, a)); break;
    case 3: MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_sync_kernel<3>, a)); break;
    case 4: MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_sync_kernel<4>, a)); break;
    case 5: MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_sync_kernel<5>, a)); break;
    case 6: MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_sync_kernel<6>, a)); break;
    case 7: MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_sync_kernel<7>, a)); break;
    default: MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_sync_kernel<8>, a)); break;
    }
    h->sync_count += 1;
    if (beside) {
        MDPP_CUDA(cudaEventRecord(h->sync_join, h->sync_stream));
        h->sync_pending = true;
        h->chain = saved_chain; // the caller's stream goes on where it was
    } else {
        h->chain = !no_pdl;
    }
    return MDPP_OK;
}

// All ranks of a group attached in this process: one cooperative launch (q_sync_group_kernel).
int mdpp_sync_emulate(mdpp_handle *const *handles, int world, void *stream) {
    if (!handles || world < 2 || world > MDPP_SYNC_MAX_RANKS) return fail(MDPP_EINVAL, "mdpp_sync_emulate: bad arguments");
    SyncGroupArgs g{};
    for (int r = 0; r < world; r++) {
        mdpp_handle *h = handles[r];
        if (!h || h->sync_world != world || h->sync_rank != r) return fail(MDPP_EINVAL, "handle %d is not attached as rank %d of %d", r, r, world);
        if (h->cfg.n_states != handles[0]->cfg.n_states) return fail(MDPP_EINVAL, "the ranks' tables differ in size");
        g.r[r] = sync_args(h, ++h->sync_epoch);
        h->sync_count += 1;
        h->launches += 1;
        h->chain = false;
    }
    g.ctas_per_rank = sync_ctas(handles[0]);
    const unsigned grid = g.ctas_per_rank * (unsigned)world;
    const void *fn = nullptr;
    switch (world) {
    case 2: fn = (const void *)q_sync_group_kernel<2>; break;
    case 3: fn = (const void *)q_sync_group_kernel<3>; break;
    case 4: fn = (const void *)q_sync_group_kernel<4>; break;
    case 5: fn = (const void *)q_sync_group_kernel<5>; break;
    case 6: fn = (const void *)q_sync_group_kernel<6>; break;
    case 7: fn = (const void *)q_sync_group_kernel<7>; break;
    default: fn = (const void *)q_sync_group_kernel<8>; break;
    }
    int dev = 0, per_sm = 0, sms = 0;
    MDPP_CUDA(cudaGetDevice(&dev));
    MDPP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MDPP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kSyncThreads, 0));
    if ((unsigned)(per_sm * sms) < grid) return fail(MDPP_EINVAL, "the group's %u CTAs cannot be co-resident on this device", grid);
    void *args[] = {(void *)&g};
    MDPP_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kSyncThreads), args, 0, (cudaStream_t)stream));
    return MDPP_OK;
}

int64_t mdpp_sync_buffer_bytes(const mdpp_config *cfg) {
    if (validate_config(cfg) != MDPP_OK) return -1;
    return (int64_t)(sync_table_bytes(*cfg) + align_up(kSyncFlagBytes));
}
int mdpp_sync_alloc(int64_t bytes, void **dev_ptr_out, uint8_t *ipc_handle_out) {
    if (!dev_ptr_out || bytes <= 0) return fail(MDPP_EINVAL, "mdpp_sync_alloc: bad arguments");
    void *p = nullptr;
    MDPP_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess && ipc_handle_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) == MDPP_IPC_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t hd;
        e = cudaIpcGetMemHandle(&hd, p);
        if (e == cudaSuccess) memcpy(ipc_handle_out, &hd, sizeof hd);
    }
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(MDPP_ECUDA, "mdpp_sync_alloc: %s", cudaGetErrorString(e));
    }
    *dev_ptr_out = p;
    return MDPP_OK;
}
int mdpp_sync_free(void *dev_ptr) {
    if (dev_ptr) MDPP_CUDA(cudaFree(dev_ptr));
    return MDPP_OK;
}
int mdpp_sync_open(const uint8_t *ipc_handle, void **peer_ptr_out) {
    if (!ipc_handle || !peer_ptr_out) return fail(MDPP_EINVAL, "mdpp_sync_open: bad arguments");
    cudaIpcMemHandle_t hd;
    memcpy(&hd, ipc_handle, sizeof hd);
    MDPP_CUDA(cudaIpcOpenMemHandle(peer_ptr_out, hd, cudaIpcMemLazyEnablePeerAccess));
    return MDPP_OK;
}
int mdpp_sync_close(void *peer_ptr) {
    if (peer_ptr) MDPP_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return MDPP_OK;
}
int mdpp_sync_attach(mdpp_handle *h, int rank, int world, void *const *bufs, int interval) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (world < 1 || world > MDPP_SYNC_MAX_RANKS || rank < 0 || rank >= world)
        return fail(MDPP_EINVAL, "rank %d / world %d out of range (at most %d ranks)", rank, world, MDPP_SYNC_MAX_RANKS);
    if (interval < 0) return fail(MDPP_EINVAL, "interval must be >= 0");
    if ((size_t)h->cfg.n_states * MDPP_Q_ROW * 4 > (size_t)MDPP_SYNC_MAX_TABLE_BYTES)
        return fail(MDPP_EINVAL, "table of %lld states is too large for the one-shot exchange", (long long)h->cfg.n_states);
    if (world > 1 && !bufs) return fail(MDPP_EINVAL, "bufs is NULL");
    for (int r = 0; r < world; r++) {
        if (world > 1 && (!bufs[r] || ((uintptr_t)bufs[r] & 255))) return fail(MDPP_EINVAL, "bufs[%d] is NULL or not 256-byte aligned", r);
        h->sync_bufs[r] = world > 1 ? bufs[r] : nullptr;
    }
    h->sync_world = world;
    h->sync_rank = rank;
    h->sync_interval = interval;
    return MDPP_OK;
}
int mdpp_q_sync_replicas(mdpp_handle *h, float *q, void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (!q) return fail(MDPP_EINVAL, "q is NULL");
    h->chain = false; // the caller may have put anything on the stream since our last launch
    int rc = q_sync_launch(h, q, (cudaStream_t)stream);
    h->chain = false;
    h->mirror_ok = false; // (scan mode) the exchange rewrites the table
    return rc;
}
uint64_t mdpp_sync_count(mdpp_handle *h) { return h ? h->sync_count : 0; }

int mdpp_q_table_written(mdpp_handle *h) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    h->mirror_ok = false; // scan mode: the next learner call copies the table into the mirror first
    return MDPP_OK;
}

int mdpp_q_substep(mdpp_handle *h, float *q, const mdpp_learner *lp, int car_slot, uint64_t round,
                   const uint8_t *forced, uint8_t *action_out, uint32_t *state_idx, float *reward, void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (!q || !lp) return fail(MDPP_EINVAL, "q / learner is NULL");
    if (car_slot < 0 || car_slot >= h->cfg.n_cars) return fail(MDPP_EINVAL, "car_slot %d out of range", car_slot);
    if ((uintptr_t)q & 31) return fail(MDPP_EINVAL, "q must be 32-byte aligned");
    if (int rc = validate_learner(lp)) return rc;
    h->chain = false; // the caller may have put anything on the stream since our last launch
    return q_env_launch(h, q, lp, car_slot, round, forced, action_out, state_idx, reward, (cudaStream_t)stream);
}

int mdpp_q_finalize(mdpp_handle *h, float *q, const mdpp_learner *lp, void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (!q || !lp) return fail(MDPP_EINVAL, "q / learner is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    h->chain = false;
    float *alt = (float *)(h->ws + h->cv.qalt);
    int rc = h->cv.dense ? MDPP_OK : scan_begin(h, q, q, alt, s);
    if (!rc) rc = q_pull_launch(h, q, alt, lp, s);
    if (rc) return rc;
    float *cur = alt, *other = q;
    rc = h->cv.dense ? q_copy_launch(h, alt, q, s) : scan_settle(h, q, cur, other, s);
    h->chain = false;
    return rc;
}

int mdpp_q_train_rounds(mdpp_handle *h, float *q, const mdpp_learner *lp, uint64_t first_round, int rounds,
                        void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (!q || !lp) return fail(MDPP_EINVAL, "q / learner is NULL");
    if ((uintptr_t)q & 31) return fail(MDPP_EINVAL, "q must be 32-byte aligned");
    if (lp->flags & MDPP_LEARN_OBSERVED) return fail(MDPP_EINVAL, "MDPP_LEARN_OBSERVED is a per-sub-step flag");
    if (int rc = validate_learner(lp)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    h->chain = false;
    float *cur = q, *alt = (float *)(h->ws + h->cv.qalt); // the table ping-pongs (dense: whole copies; scan: a mirror)
    const bool frozen = (lp->flags & MDPP_LEARN_FROZEN) != 0;
    const bool fused = lp->flags == 0 && round_eligible(h);
    if (!frozen && !h->cv.dense)
        if (int rc = scan_begin(h, q, cur, alt, s)) return rc;
    for (int r = 0; r < rounds; r++) {
        if (sync_due(h, first_round + (uint64_t)r)) { // replica averaging before this round (q_sync.cuh)
            if (h->scan_stale) { // scan mode: both copies complete, the caller's table current
                int rc = scan_settle(h, q, cur, alt, s);
                if (rc) return rc;
            } else if (cur != q) { // an odd number of pulls so far: the exchange runs on the caller's table
                int rc = q_copy_launch(h, cur, q, s);
                if (rc) return rc;
                std::swap(cur, alt);
            }
            int rc = q_sync_launch(h, q, s, h->cv.dense);
            if (rc) return rc;
            if (h->sync_world > 1) h->mirror_ok = false; // the exchange rewrites the table
            if (!h->cv.dense && h->sync_world > 1 && !frozen) { // ... and the mirror follows
                if ((rc = sync_join(h, s))) return rc;
                if ((rc = scan_begin(h, q, cur, alt, s))) return rc;
            }
        }
        const int room = sync_room(h, first_round + (uint64_t)r, rounds - r); // rounds up to the next averaging
        if (h->cv.scan && lp->flags == 0) { // 3-5 cars: all cars of several pure-exploration rounds in one env pass
            if (const int n = pure_scan_rounds_fit(h, lp, room)) {
                int rc = q_pure_scan_rounds_launch(h, q, cur, alt, lp, first_round + (uint64_t)r, n, s);
                if (rc) return rc;
                r += n - 1;
                continue;
            }
        }
        if (int rc = overlap_join(h, s)) return rc; // whatever follows reads the table or the marks
        if (h->cv.dense && h->scan_stale) // mid-size tables leaving the scan mode: the dense pulls know nothing of its bitmaps
            if (int rc = scan_settle(h, q, cur, alt, s)) return rc;
        if (fused) { // training fast path: a whole round per launch (q_round.cuh), several while every env explores
            if (const int n = pure_rounds_fit(h, lp, room)) {
                int rc = q_pure_rounds_launch(h, cur, alt, lp, first_round + (uint64_t)r, n, s);
                if (rc) return rc;
                r += n - 1;
                continue;
            }
            // A round with greedy decisions needs two env passes either way.  The fused pair K0 / K1 steps car 1
            // in both (explorers in K0, parked lanes in K1): 35-39 us per round at 2^20 envs for every epsilon
            // < 1 of the schedule, against 31-32 us for one sub-step kernel + pull per car (tools/exp_mixed.py).
            // So the round kernel is taken while every env certainly explores, the sub-step kernels otherwise.
            const bool pure = lp->eps_q16[0] >= 65536u && h->ep_bound < (int64_t)lp->eps_threshold[0];
            if (pure || h->k1_rounds) {
                int rc = q_round_launch(h, cur, alt, lp, first_round + (uint64_t)r, s);
                if (rc) return rc;
                continue;
            }
            if (const int n = mixed_rounds_fit(h, room)) { // several such rounds in one cooperative launch
                int rc = q_mixed_rounds_launch(h, cur, alt, lp, first_round + (uint64_t)r, n, s);
                if (rc) return rc;
                r += n - 1;
                continue;
            }
        }
        for (int c = 0; c < h->cfg.n_cars; c++) {
            int rc = q_env_launch(h, cur, lp, c, first_round + (uint64_t)r, nullptr, nullptr, nullptr, nullptr, s);
            if (rc) return rc;
            if (frozen) { // nothing was marked; back-to-back env kernels must not overlap (records)
                h->chain = false;
                continue;
            }
            rc = q_pull_launch(h, cur, alt, lp, s);
            if (rc) return rc;
            std::swap(cur, alt);
        }
    }
    int rc = overlap_join(h, s);
    if (!rc && h->scan_stale) rc = scan_settle(h, q, cur, alt, s); // both copies complete: the caller's table is current
    else if (cur != q) rc = q_copy_launch(h, cur, q, s); // odd number of sub-steps: bring the table home
    if (!rc) rc = sync_join(h, s); // (a frozen learner has no pull that would have joined)
    h->chain = false;
    return rc;
}

int mdpp_query_transitions(mdpp_handle *h, int car_slot, int reward_kind, const uint32_t *state_idx,
                           const uint8_t *action, int64_t n, uint8_t *legal, uint32_t *next_idx, float *reward,
                           uint8_t *next_legal, void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (!state_idx) return fail(MDPP_EINVAL, "state_idx is NULL");
    if (n <= 0 || n > h->n_envs) return fail(MDPP_EINVAL, "n must be in 1..n_envs (entry i reads env i's cars)");
    if (car_slot < 0 || car_slot >= h->cfg.n_cars) return fail(MDPP_EINVAL, "car_slot %d out of range", car_slot);
    if (reward_kind != MDPP_REWARD_Q && reward_kind != MDPP_REWARD_DQN) return fail(MDPP_EINVAL, "unknown reward kind");
    cudaStream_t s = (cudaStream_t)stream;
    MDPP_DISPATCH_NC(h, query_kernel, grid_for(n), s, h->dev, (const uint4 *)recs_of(h), n, car_slot, reward_kind,
                     state_idx, action, legal, next_idx, reward, next_legal);
    MDPP_CUDA(cudaGetLastError());
    return MDPP_OK;
}

int mdpp_update_car(mdpp_handle *h, int car_slot, const uint8_t *nodes, void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (!nodes) return fail(MDPP_EINVAL, "nodes is NULL");
    if (car_slot < 0 || car_slot >= h->cfg.n_cars) return fail(MDPP_EINVAL, "car_slot %d out of range", car_slot);
    cudaStream_t s = (cudaStream_t)stream;
    MDPP_DISPATCH_NC(h, update_car_kernel, grid_for(h->n_envs), s, h->dev, recs_of(h), h->n_envs, car_slot, nodes);
    MDPP_CUDA(cudaGetLastError());
    return MDPP_OK;
}

int mdpp_dqn_observe(mdpp_handle *h, int car_slot, uint8_t *features, uint8_t *legal, uint32_t *state_idx, void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (car_slot < 0 || car_slot >= h->cfg.n_cars) return fail(MDPP_EINVAL, "car_slot %d out of range", car_slot);
    if (features && ((uintptr_t)features & 15)) return fail(MDPP_EINVAL, "features must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const FastTabs *tabs = (const FastTabs *)(h->ws + h->cv.fast);
    const uint32_t stride = (uint32_t)(h->cfg.n_dest_nodes * h->cfg.n_local_nodes);
    MDPP_DISPATCH_NC_C(h, car_slot, dqn_observe_kernel, grid_for(h->n_envs), s, h->dev, recs_of(h), h->n_envs, tabs, stride,
                       features, legal, state_idx);
    MDPP_CUDA(cudaGetLastError());
    return MDPP_OK;
}

int mdpp_dqn_act(mdpp_handle *h, int car_slot, int reward_kind, const uint8_t *actions, uint8_t *next_features,
                 float *reward, uint8_t *next_legal, uint8_t *done, uint8_t *status, int auto_reset, uint64_t round,
                 uint32_t seed, void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (!actions) return fail(MDPP_EINVAL, "actions is NULL");
    if (car_slot < 0 || car_slot >= h->cfg.n_cars) return fail(MDPP_EINVAL, "car_slot %d out of range", car_slot);
    if (reward_kind != MDPP_REWARD_Q && reward_kind != MDPP_REWARD_DQN) return fail(MDPP_EINVAL, "unknown reward kind");
    if (next_features && ((uintptr_t)next_features & 15)) return fail(MDPP_EINVAL, "next_features must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const FastTabs *tabs = (const FastTabs *)(h->ws + h->cv.fast);
    const uint32_t stride = (uint32_t)(h->cfg.n_dest_nodes * h->cfg.n_local_nodes);
    MDPP_DISPATCH_NC_C(h, car_slot, dqn_act_kernel, grid_for(h->n_envs), s, h->dev, recs_of(h), h->n_envs, h->env_id_base,
                       reward_kind, tabs, stride, actions, next_features, reward, next_legal, done, status, auto_reset, round,
                       seed, control_of(h));
    MDPP_CUDA(cudaGetLastError());
    if (auto_reset && car_slot == h->cfg.n_cars - 1) {
        h->ep_bound += 1;
        h->ep_device_max_valid = false;
    }
    return MDPP_OK;
}

int64_t mdpp_joint_state_count(const mdpp_config *cfg) {
    if (validate_config(cfg)) return -1;
    const uint64_t V = 2ull * (uint64_t)cfg->n_local_nodes + 1ull;
    uint64_t n = 1;
    for (int c = 0; c < cfg->n_cars; c++) {
        if (n > ((uint64_t)1 << 40) / V) return -fail(MDPP_EINVAL, "joint state space too large");
        n *= V;
    }
    return (int64_t)n;
}

int mdpp_deadlock_reachability(mdpp_handle *h, uint32_t *reach_bitmap, int32_t *sweeps_out, void *stream) {
    if (!h || !reach_bitmap) return fail(MDPP_EINVAL, "handle / bitmap is NULL");
    const int64_t n = mdpp_joint_state_count(&h->cfg);
    if (n <= 0) return MDPP_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t *changed = (uint32_t *)(h->ws + h->cv.stage); // scratch word in the staging area
    const size_t words = ((size_t)n + 31) / 32;
    MDPP_CUDA(cudaMemsetAsync(reach_bitmap, 0, words * 4, s));
    int sweeps = 0;
    for (;;) {
        MDPP_CUDA(cudaMemsetAsync(changed, 0, 4, s));
        MDPP_DISPATCH_NC(h, reach_sweep_kernel, grid_for(n), s, h->dev, reach_bitmap, (uint64_t)0, (uint64_t)n, changed);
        MDPP_CUDA(cudaGetLastError());
        uint32_t flag = 0;
        MDPP_CUDA(cudaMemcpyAsync(&flag, changed, 4, cudaMemcpyDeviceToHost, s));
        MDPP_CUDA(cudaStreamSynchronize(s));
        sweeps++;
        if (!flag) break;
        if (sweeps > 100000) return fail(MDPP_ECUDA, "reachability did not converge");
    }
    if (sweeps_out) *sweeps_out = sweeps;
    return MDPP_OK;
}

// One forward sweep over the joint states [first_state, first_state + n_range): the building block of the search sharded
// over ranks (parallel.sharded_reachability): every rank sweeps its own range of the replicated bitmap, the ranks
// exchange their ranges, until a sweep changes nothing anywhere.
int mdpp_deadlock_sweep_range(mdpp_handle *h, uint32_t *reach_bitmap, int64_t first_state, int64_t n_range,
                              uint32_t *changed, void *stream) {
    if (!h || !reach_bitmap || !changed) return fail(MDPP_EINVAL, "handle / bitmap / changed is NULL");
    const int64_t n = mdpp_joint_state_count(&h->cfg);
    if (n <= 0) return MDPP_EINVAL;
    if (first_state < 0 || (first_state & 31) || n_range < 0)
        return fail(MDPP_EINVAL, "first_state must be a non-negative multiple of 32 (a rank owns whole bitmap words), n_range >= 0");
    const int64_t end = std::min<int64_t>(n, first_state + n_range);
    if (end <= first_state) return MDPP_OK; // an empty shard
    cudaStream_t s = (cudaStream_t)stream;
    MDPP_DISPATCH_NC(h, reach_sweep_kernel, grid_for(end - first_state), s, h->dev, reach_bitmap, (uint64_t)first_state,
                     (uint64_t)end, changed);
    MDPP_CUDA(cudaGetLastError());
    return MDPP_OK;
}

// ---- the same bitmap by a backward breadth-first search in one cooperative launch (reach.cuh) ------------------
int64_t mdpp_deadlock_scratch_bytes(const mdpp_config *cfg) {
    const int64_t n = mdpp_joint_state_count(cfg);
    if (n <= 0 || n >= (int64_t)1 << 32) return 0; // 32-bit joint indices only: the sweeps cover the rest
    return (int64_t)(2 * (((size_t)n + 31) / 32) * 4 + 256);
}

extern "C++" {
template <int NC>
static int reach_bfs_launch(mdpp_handle *h, ReachArgs &ra, cudaStream_t s, unsigned *grid_out) {
    const size_t smem = sizeof(ReachSmem<NC>);
    const void *fn = (const void *)reach_bfs_kernel<NC>;
    int per_sm = 0, coop = 0, dev = 0;
    MDPP_CUDA(cudaGetDevice(&dev));
    MDPP_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return fail(MDPP_ECUDA, "cooperative launch is not available on this device");
    MDPP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MDPP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kReachThreads, smem));
    if (per_sm < 1) return fail(MDPP_ECUDA, "reach_bfs_kernel does not fit an SM");
    const unsigned chunks = (ra.words + kReachChunkWords - 1) / kReachChunkWords;
    const unsigned grid = std::max(1u, std::min((unsigned)(per_sm * h->sm_count), chunks));
    void *args[] = {(void *)&ra};
    MDPP_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kReachThreads), args, smem, s));
    *grid_out = grid;
    return MDPP_OK;
}
} // extern "C++"

int mdpp_deadlock_reachability_bfs(mdpp_handle *h, uint32_t *reach_bitmap, void *scratch, int64_t scratch_bytes,
                                   int32_t *levels_out, void *stream) {
    if (!h || !reach_bitmap || !scratch) return fail(MDPP_EINVAL, "handle / bitmap / scratch is NULL");
    const int64_t need = mdpp_deadlock_scratch_bytes(&h->cfg);
    if (need == 0) return fail(MDPP_EINVAL, "joint state space too large for 32-bit indices: use mdpp_deadlock_reachability");
    if (scratch_bytes < need || ((uintptr_t)scratch & 15)) return fail(MDPP_EWORKSPACE, "scratch of %lld B (16-byte aligned) required", (long long)need);
    const int64_t n = mdpp_joint_state_count(&h->cfg);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t words = ((size_t)n + 31) / 32;
    ReachArgs ra{};
    ra.reach = reach_bitmap;
    ra.front[0] = (uint32_t *)scratch;
    ra.front[1] = ra.front[0] + words;
    ra.counts = ra.front[1] + words;
    ra.levels_out = ra.counts + 4;
    ra.errors = ra.counts + 5;
    ra.barrier = (uint32_t *)(h->ws + h->cv.barrier);
    ra.words = (uint32_t)words;
    ra.n_joint = (uint32_t)n;
    ra.V = 2u * (uint32_t)h->cfg.n_local_nodes + 1u;
    ra.node_base = (uint32_t)h->dev.node_base;
    ra.parking_node = (uint32_t)h->cfg.parking_node;
    ra.rank_stride = (uint32_t)(h->cfg.n_dest_nodes * h->cfg.n_local_nodes);
    ra.tabs = (const FastTabs *)(h->ws + h->cv.fast);
    MDPP_CUDA(cudaMemsetAsync(reach_bitmap, 0, words * 4, s));
    MDPP_CUDA(cudaMemsetAsync(scratch, 0, (size_t)need, s));
    reach_seed_kernel<<<1, 32, 0, s>>>(reach_bitmap, ra.front[0], (uint32_t)(n - 1));
    MDPP_CUDA(cudaGetLastError());
    unsigned grid = 0;
    int rc;
    switch (h->cfg.n_cars) {
    case 1: rc = reach_bfs_launch<1>(h, ra, s, &grid); break;
    case 2: rc = reach_bfs_launch<2>(h, ra, s, &grid); break;
    case 3: rc = reach_bfs_launch<3>(h, ra, s, &grid); break;
    case 4: rc = reach_bfs_launch<4>(h, ra, s, &grid); break;
    default: rc = reach_bfs_launch<5>(h, ra, s, &grid); break;
    }
    if (rc) return rc;
    uint32_t res[2] = {0u, 0u}; // levels, reverse-table overflows
    MDPP_CUDA(cudaMemcpyAsync(res, ra.levels_out, 8, cudaMemcpyDeviceToHost, s));
    MDPP_CUDA(cudaStreamSynchronize(s));
    h->chain = false;
    if (res[1]) return fail(MDPP_ECUDA, "reverse move table overflow (%u values): raise kRevSlots", res[1]);
    if (levels_out) *levels_out = (int32_t)res[0];
    return MDPP_OK;
}

// ---- DQN replay memory and Double-DQN target (dqn_replay.cuh); no handle: plain device arrays ------------------
static ReplayMem replay_mem(int64_t n_keys, uint32_t *present, uint32_t *m_next, float *m_reward, uint8_t *m_mask,
                            uint8_t *m_features, uint64_t *counters) {
    ReplayMem m{};
    m.present = present;
    m.next = m_next;
    m.reward = m_reward;
    m.mask = m_mask;
    m.features = m_features;
    m.counters = (unsigned long long *)counters;
    m.n_keys = (uint32_t)n_keys;
    return m;
}
int mdpp_replay_insert(int64_t n, const uint32_t *state_idx, const uint8_t *action, const uint32_t *next_idx,
                       const float *reward, const uint8_t *next_mask, const uint8_t *status, const uint8_t *features,
                       const uint8_t *next_features, int64_t n_keys, uint32_t *present, uint32_t *m_next, float *m_reward,
                       uint8_t *m_mask, uint8_t *m_features, uint64_t *counters, void *stream) {
    if (n <= 0) return MDPP_OK;
    if (!state_idx || !action || !next_idx || !reward || !next_mask || !present || !m_next || !m_reward || !m_mask || !counters)
        return fail(MDPP_EINVAL, "mdpp_replay_insert: NULL argument");
    if (n_keys <= 0 || n_keys >= (int64_t)1 << 32 || (n_keys & 31)) return fail(MDPP_EINVAL, "n_keys must be a positive multiple of 32 below 2^32");
    if (m_features && (!features || !next_features)) return fail(MDPP_EINVAL, "feature memory without feature rows");
    if (m_features && (((uintptr_t)m_features | (uintptr_t)features | (uintptr_t)next_features) & 15))
        return fail(MDPP_EINVAL, "feature arrays must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const ReplayMem m = replay_mem(n_keys, present, m_next, m_reward, m_mask, m_features, counters);
    // duplicates of keys stored by earlier calls are compared with the stored payload first (the payload of a key this
    // call stores is only complete when the call's kernel has finished)
    replay_check_kernel<<<grid_for(n), kThreads, 0, s>>>(n, state_idx, action, next_idx, reward, status, m);
    replay_insert_kernel<<<grid_for(n), kThreads, 0, s>>>(n, state_idx, action, next_idx, reward, next_mask, status,
                                                         features, next_features, m);
    MDPP_CUDA(cudaGetLastError());
    return MDPP_OK;
}
int mdpp_replay_compact(int64_t n_keys, const uint32_t *present, uint32_t *list, uint32_t *count, void *stream) {
    if (!present || !list || !count) return fail(MDPP_EINVAL, "mdpp_replay_compact: NULL argument");
    if (n_keys <= 0 || n_keys >= (int64_t)1 << 32 || (n_keys & 31)) return fail(MDPP_EINVAL, "n_keys must be a positive multiple of 32 below 2^32");
    replay_compact_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(present, (uint32_t)(n_keys / 32), (uint32_t)n_keys, list, count);
    MDPP_CUDA(cudaGetLastError());
    return MDPP_OK;
}
int mdpp_dqn_target(int64_t n, const float *q_online_next, const float *q_target_next, const uint8_t *next_mask,
                    const float *reward, const int64_t *action, const float *prediction, double discount, int index_mode,
                    float *y, float *target_out, uint64_t *counters, void *stream) {
    if (n <= 0) return MDPP_OK;
    if (!q_online_next || !q_target_next || !next_mask || !reward || !action || !prediction || !y)
        return fail(MDPP_EINVAL, "mdpp_dqn_target: NULL argument");
    if (index_mode != 0 && index_mode != 1) return fail(MDPP_EINVAL, "index_mode must be 0 (reference) or 1 (action index)");
    dqn_target_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(n, q_online_next, q_target_next, next_mask, reward,
                                                                         action, prediction, discount, index_mode, y,
                                                                         target_out, (unsigned long long *)counters);
    MDPP_CUDA(cudaGetLastError());
    return MDPP_OK;
}

int mdpp_get_stats_host(mdpp_handle *h, mdpp_stats *out_host, int reset, void *stream) {
    if (!h || !out_host) return fail(MDPP_EINVAL, "handle / out is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    // one copy of the whole control block: kStatSlots privatised counter sets (summed here) + the episode maxima
    Control *host_ctl = (Control *)h->stat_slots;
    MDPP_CUDA(cudaMemcpyAsync(host_ctl, control_of(h), sizeof(Control), cudaMemcpyDeviceToHost, s));
    if (reset) MDPP_CUDA(cudaMemsetAsync(control_of(h)->stats, 0, sizeof(mdpp_stats) * kStatSlots, s));
    MDPP_CUDA(cudaStreamSynchronize(s));
    const mdpp_stats *slots = host_ctl->stats;
    const uint32_t *max_ep = host_ctl->max_episode;
    if (h->ep_device_max_valid) { // everything launched so far has completed: the device's maximum is exact
        uint32_t m = 0;
        for (int i = 0; i < kStatSlots; i++) m = std::max(m, max_ep[i]);
        h->ep_bound = std::min<int64_t>(h->ep_bound, (int64_t)m);
    }
    mdpp_stats t;
    memset(&t, 0, sizeof t);
    for (int i = 0; i < kStatSlots; i++) {
        t.agent_steps += slots[i].agent_steps;
        t.q_updates += slots[i].q_updates;
        t.episodes_done += slots[i].episodes_done;
        t.episodes_capped += slots[i].episodes_capped;
        t.reward_sum += slots[i].reward_sum;
        t.explore_steps += slots[i].explore_steps;
        t.touched_keys += slots[i].touched_keys;
        t.internal_errors += slots[i].internal_errors;
    }
    *out_host = t;
    return MDPP_OK;
}

int mdpp_q_profile_env_kernel(mdpp_handle *h, float *q, const mdpp_learner *lp, uint64_t round, void *stream) {
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (!q || !lp) return fail(MDPP_EINVAL, "q / learner is NULL");
    if (lp->flags != 0 || !round_eligible(h)) return fail(MDPP_EINVAL, "no fused round kernel for this table / these flags");
    cudaStream_t s = (cudaStream_t)stream;
    h->chain = false;
    RoundArgs ra{};
    ra.recs = recs_of(h);
    ra.n_envs = (uint32_t)h->n_envs;
    ra.env_id_base = h->env_id_base;
    ra.round = round;
    ra.q = q;
    ra.rng_words = (uint32_t *)(h->ws + h->cv.rng);
    ra.tabs = (const FastTabs *)(h->ws + h->cv.fast);
    ra.defer_flags = (uint32_t *)(h->ws + h->cv.flags) + (size_t)h->round_parity * h->cv.cta_slots;
    ra.bits_a = mark_region(h, h->round_parity, 0);
    ra.bits_b = mark_region(h, h->round_parity, 1);
    ra.words_per_cta = h->cv.words_per_cta;
    ra.cta_slots = h->cv.cta_slots;
    ra.map_bytes = (uint32_t)((((size_t)h->cfg.n_states * MDPP_N_ACT + 15) / 16) * 16);
    ra.rank_stride = (uint32_t)(h->cfg.n_dest_nodes * h->cfg.n_local_nodes);
    ra.n_states = (uint32_t)h->cfg.n_states;
    ra.tiles_q = (uint32_t)(((h->n_envs + 31) / 32) / round_ctas(h));
    ra.tiles_r = (uint32_t)(((h->n_envs + 31) / 32) % round_ctas(h));
    const size_t smem = round_smem_bytes(h->cfg.n_cars, ra.map_bytes, h->cfg.n_cars);
    LearnLaunch ll(h, round_ctas(h), s, kRoundThreads, smem);
    if (h->cfg.n_cars == 1) {
        MDPP_CUDA(cudaFuncSetAttribute(q_round_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_round_kernel<1, 0>, h->dev, *lp, ra, control_of(h)));
    } else {
        MDPP_CUDA(cudaFuncSetAttribute(q_round_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MDPP_LAUNCH(h, cudaLaunchKernelEx(&ll.cfg, q_round_kernel<2, 0>, h->dev, *lp, ra, control_of(h)));
    }
    h->ep_bound += 1;
    return MDPP_OK;
}

int mdpp_q_profile_env_rounds(mdpp_handle *h, float *q, const mdpp_learner *lp, uint64_t round, int rounds, void *stream) {
    if (rounds == 1) return mdpp_q_profile_env_kernel(h, q, lp, round, stream);
    if (!h) return fail(MDPP_EINVAL, "handle is NULL");
    if (!q || !lp) return fail(MDPP_EINVAL, "q / learner is NULL");
    if (lp->flags != 0 || !round_eligible(h)) return fail(MDPP_EINVAL, "no fused round kernel for this table / these flags");
    if (pure_rounds_fit(h, lp, rounds) != rounds)
        return fail(MDPP_EINVAL, "%d rounds do not fit one pure-exploration launch here (2..%d, every env in the epsilon = 1 segment)", rounds, kMaxPureRounds);
    h->chain = false;
    float *cur = q, *alt = q;
    int rc = q_pure_rounds_launch(h, cur, alt, lp, round, rounds, (cudaStream_t)stream, true);
    h->chain = false;
    return rc;
}

int mdpp_q_train_rounds_host(mdpp_handle *h, float *q, const mdpp_learner *lp_host, uint64_t first_round, int rounds,
                             mdpp_stats *stats_host, void *stream) {
    int rc = mdpp_q_train_rounds(h, q, lp_host, first_round, rounds, stream);
    if (rc) return rc;
    if (!stats_host) return fail(MDPP_EINVAL, "stats_host is NULL");
    if (getenv("MDPP_HOST_STATS_COPY") && atoi(getenv("MDPP_HOST_STATS_COPY")) != 0)
        return mdpp_get_stats_host(h, stats_host, 0, stream); // the copy + synchronise form (A/B runs)
    // totals + sequence number stored into pinned host memory by one small kernel; this thread polls the word
    cudaStream_t s = (cudaStream_t)stream;
    HostStats *dev_view = nullptr;
    MDPP_CUDA(cudaHostGetDevicePointer((void **)&dev_view, h->host_stats, 0));
    const unsigned long long seq = ++h->host_seq;
    stats_publish_kernel<<<1, 128, 0, s>>>(control_of(h), dev_view, seq);
    MDPP_CUDA(cudaGetLastError());
    unsigned spins = 0;
    while (h->host_stats->seq != seq) {
        if ((++spins & 0xFFFFu) == 0u) { // a failed launch must not turn into an endless poll
            cudaError_t e = cudaStreamQuery(s);
            if (e != cudaSuccess && e != cudaErrorNotReady)
                return fail(MDPP_ECUDA, "stream failed while waiting for the counters: %s", cudaGetErrorString(e));
            if (e == cudaSuccess && h->host_stats->seq != seq)
                return fail(MDPP_ECUDA, "the counters never arrived in host memory");
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    *stats_host = h->host_stats->totals;
    if (h->ep_device_max_valid) // everything launched so far has completed: the device's maximum is exact
        h->ep_bound = std::min<int64_t>(h->ep_bound, (int64_t)h->host_stats->max_episode);
    return MDPP_OK;
}

// ---- state-id codec: pure host arithmetic, usable without a device --------------------------------
static uint64_t binom_u64(uint64_t n, int j) {
    if (n < (uint64_t)j) return 0;
    uint64_t r = 1;
    for (int i = 1; i <= j; i++) r = r * (n - (uint64_t)j + (uint64_t)i) / (uint64_t)i;
    return r;
}
// ordering of the reference's ascend_array (src/env_gym.py:267-279): (box number, [src, dst] as strings)
static bool other_before(const int32_t *a, const int32_t *b) {
    int ka[4] = {a[0] % 9, a[0] < 9, a[1] < 9, a[1] % 9}; // 'D' (floor 1) sorts before 'U' (floor 0)
    int kb[4] = {b[0] % 9, b[0] < 9, b[1] < 9, b[1] % 9};
    for (int i = 0; i < 4; i++)
        if (ka[i] != kb[i]) return ka[i] < kb[i];
    return false;
}

int mdpp_decode_state(const mdpp_config *cfg, int64_t idx, int32_t *pnode, int32_t *dnode, int32_t *others) {
    if (validate_config(cfg)) return -MDPP_EINVAL;
    if (idx < 0 || idx >= cfg->n_states) return -fail(MDPP_EINVAL, "state index out of range");
    const int node_base = find_base(cfg->node_local, MDPP_NODES), box_base = find_base(cfg->src_box_idx, MDPP_BOXES);
    int64_t nl = idx % cfg->n_local_nodes, t = idx / cfg->n_local_nodes;
    int64_t dn = t % cfg->n_dest_nodes;
    uint64_t rank = (uint64_t)(t / cfg->n_dest_nodes);
    *pnode = (int32_t)nl + node_base;
    *dnode = cfg->dest_node[dn];
    int n = 0;
    for (int j = cfg->n_cars - 1; j >= 1; j--) { // unrank the multiset, largest element first
        uint64_t d = (uint64_t)j - 1;
        while (binom_u64(d + 1, j) <= rank) d++;
        rank -= binom_u64(d, j);
        uint64_t code = d - (uint64_t)j + 1;
        if (code == 0) continue;
        code -= 1;
        others[2 * n] = (int32_t)(code / (uint64_t)cfg->n_dest_boxes) + box_base;
        others[2 * n + 1] = cfg->dest_box[code % (uint64_t)cfg->n_dest_boxes];
        n++;
    }
    for (int i = 1; i < n; i++) // insertion sort into the reference's order
        for (int k = i; k > 0 && other_before(others + 2 * k, others + 2 * (k - 1)); k--) {
            std::swap(others[2 * k], others[2 * k - 2]);
            std::swap(others[2 * k + 1], others[2 * k - 1]);
        }
    return n;
}

int64_t mdpp_encode_state(const mdpp_config *cfg, int32_t pnode, int32_t dnode, int32_t n_others, const int32_t *others) {
    if (validate_config(cfg)) return -MDPP_EINVAL;
    if (n_others < 0 || n_others > cfg->n_cars - 1) return -fail(MDPP_EINVAL, "too many others for n_cars");
    if (pnode < 0 || pnode >= MDPP_NODES || cfg->node_local[pnode] == MDPP_NONE) return -fail(MDPP_EINVAL, "primary node outside the universe");
    int dn = -1;
    for (int i = 0; i < cfg->n_dest_nodes; i++)
        if (cfg->dest_node[i] == dnode) dn = i;
    if (dn < 0) return -fail(MDPP_EINVAL, "destination node is not a destination");
    uint64_t codes[MDPP_MAX_CARS] = {0, 0, 0, 0, 0};
    const int k = cfg->n_cars - 1;
    for (int i = 0; i < n_others; i++) {
        int sb = others[2 * i], db = -1;
        if (sb < 0 || sb >= MDPP_BOXES || cfg->src_box_idx[sb] == MDPP_NONE) return -fail(MDPP_EINVAL, "source box outside the universe");
        for (int j = 0; j < cfg->n_dest_boxes; j++)
            if (cfg->dest_box[j] == others[2 * i + 1]) db = j;
        if (db < 0) return -fail(MDPP_EINVAL, "destination box is not a destination");
        codes[i] = 1 + (uint64_t)cfg->src_box_idx[sb] * (uint64_t)cfg->n_dest_boxes + (uint64_t)db;
    }
    std::sort(codes, codes + k);
    uint64_t rank = 0;
    for (int j = 1; j <= k; j++) rank += binom_u64(codes[j - 1] + (uint64_t)j - 1, j);
    return (int64_t)((rank * (uint64_t)cfg->n_dest_nodes + (uint64_t)dn) * (uint64_t)cfg->n_local_nodes + cfg->node_local[pnode]);
}

} // extern "C"
