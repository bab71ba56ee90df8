/*
 * fjsp_oracle.c — CPU restatement of the reference environment step.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product path (libfjsp_b200.so) never does and has no CPU fallback.
 *
 * What it restates (reference = /root/reference, unmodified, Python + SimPy):
 *   FJSPSimulation.step              FJSPSimulation.py:144-242
 *   FJSPSimulation.reset/generate    FJSPSimulation.py:89-131,286-323
 *   _check_order_completions         FJSPSimulation.py:245-258
 *   add_tray_to_packaging            FJSPSimulation.py:402-430
 *   PickupStationAgent               agents/PickupStationAgent.py:58-232
 *   AGVAgent                         agents/AGVAgent.py:53-403
 *   MachineAgent (+Small/Big)        agents/MachineAgent.py:62-169
 *   PackagingAgent                   agents/PackagingAgent.py:54-153
 *   Tray / Storage                   models/Tray.py:16-58, models/Storage.py:16-36
 *   RewardModel                      utils/RewardModel.py:12-110
 * plus the SimPy 4 event ordering the reference relies on (simpy>=4.0.0, requirements.txt:5; not
 * vendored): heap key (time, priority, eid), URGENT Initialize before NORMAL timeouts at equal time,
 * run(until=t) leaves NORMAL events due exactly at t for the next run, one grant per _trigger_put
 * (SURVEY.md Appendix B).  The representation is deliberately object-like (products, trays, FIFOs,
 * one process record per SimPy process) and unlike the packed state of the CUDA path.
 *
 * Pinning: the reference ships no tests or golden vectors, so this restatement is
 * pinned against (a) golden trajectories generated here by executing the unmodified reference
 * (oracle/gen_golden.py -> tests/golden/, tests/test_oracle_golden.py) and (b) live differential runs
 * against the reference in this container (tests/test_oracle_vs_reference.py).  In both, the
 * unmodified reference runs on oracle/shims/simpy, a restatement of the SimPy 4 core from its
 * documented semantics: SimPy itself is not vendored by the reference and not installable here.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "../include/fjsp_b200.h"

/* Two builds of this file: the default one (episodes of <= 32 orders and <= 253 steps: small envs, thousands of them for
 * the CPU baseline) and -DFJSP_ORACLE_LONG (long order streams: up to 4095 orders, 1000 trays, order arrivals). */
#ifdef FJSP_ORACLE_LONG
#define ORC_MAX_ORDERS FJSP_LONG_MAX_ORDERS
#define MAX_TRAYS 1100 /* min(num_trays, 1000) trays ever reach the pickup station (FJSPSimulation.py:96) */
#define LIST_CAP 1100
#else
#define ORC_MAX_ORDERS FJSP_MAX_ORDERS
#define MAX_TRAYS 256 /* one tray per LOAD at most, episodes are <= 253 steps */
#define LIST_CAP 320
#endif
#define MAX_PRODUCTS (ORC_MAX_ORDERS * FJSP_MAX_ORDER_PRODUCTS)

enum { LOC_PICKUP = 0, LOC_BIG = 1, LOC_SMALL = 2, LOC_STORAGE = 3, LOC_PACKAGING = 4, LOC_NONE = -1 };
enum { TYPE_SMALL = 1, TYPE_MEDIUM = 2, TYPE_BIG = 3 };
enum { COL_RED = 1, COL_BLUE = 2, COL_GREEN = 3 };

typedef struct Prod {
    int id, type, colour, order, idx;
    int is_processed, is_packaged;
} Prod;

typedef struct Order {
    int id, n;
    int in_process; /* an AGV has taken the order's first tray (the port's long layout gives it an order slot then) */
    Prod* products[FJSP_MAX_ORDER_PRODUCTS];
    int is_complete, completion_step;
} Order;

typedef struct Tray {
    int id, n, order_id;
    Prod* products[8];
} Tray;

typedef struct TrayList {
    Tray* a[LIST_CAP];
    int n;
} TrayList;

typedef struct Machine {
    int proc_steps;
    int is_busy;
    Tray* current_tray;
    int progress_done; /* processing_progress == 1.0 */
    TrayList queue, ready;
    /* the (single) live SimPy process of _simpy_processing_process */
    Tray* proc_tray;      /* non-NULL while the process exists */
    int proc_created_now; /* created in this step's action phase, Initialize pending */
    int proc_next_idx, proc_next_fire;
} Machine;

typedef struct PkRun {
    Prod* prod;
    int finish_step;
} PkRun;

typedef struct Pack {
    int colour;
    int is_busy;
    Prod* current_product;
    double progress;
    Prod* queue[LIST_CAP];
    int qn;
    int products_completed;
    /* simpy.Resource(capacity): users = granted requests; put_queue = waiting requests */
    int users;
    PkRun running[64];
    int nrunning;
    Prod* put_queue[LIST_CAP];
    int nput;
    /* processes created by START in this step's action phase (Initialize pending) */
    Prod* created[LIST_CAP];
    int ncreated;
} Pack;

typedef struct Cell {
    /* agv */
    int agv_row, agv_col;
    Tray* carrying;
    int is_moving;
    int move_created_now, move_target_loc, move_arrive_step;
    int holds_dock; /* scaled shop: this AGV occupies (or has been granted) the single dock of the pickup station */
    /* machines, storage, packaging */
    Machine machine[2]; /* 0 small, 1 big */
    TrayList storage;
    Pack pack[4];
} Cell;

typedef struct OracleEnv {
    FjspConfig cfg;
    int small_steps, big_steps, pack_steps;
    Prod prods[MAX_PRODUCTS];
    int nprods;
    Order orders[ORC_MAX_ORDERS];
    int norders;       /* orders that exist (arrived) */
    int norders_table; /* orders whose attributes are known (long order streams: arrivals reveal the next one) */
    uint64_t seed, genv; /* arrival stream (long order streams) */
    uint32_t episode;
    Tray trays[MAX_TRAYS];
    int ntrays_total;
    /* pickup station */
    int order_queue_head; /* order_queue = orders[head..norders) */
    Order* current_order;
    int current_order_product_idx;
    int trays_next; /* trays_at_station = ids (ntrays_total-1-trays_next) downwards */
    Tray* current_tray;
    TrayList ps_ready;
    /* cells: each has its AGV, machines, storage and packaging stations (one cell = the reference shop) */
    struct Cell cells[FJSP_MAX_CELLS];
    int ncells;
    /* shared floor (include/fjsp_b200.h): nagv AGVs on the stations of cells[0]; AGV j's own fields (position, tray,
     * movement) live in cells[j], nothing else of cells[1..] is used.  nagv = 1: the reference. */
    int nagv;
    /* tracking */
    int current_step, completed_orders, total_products_packaged, fault;
} OracleEnv;

void fjsp_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* ------------------------------------------------------------------ helpers */
static void tl_push(TrayList* l, Tray* t) {
    if (l->n < LIST_CAP) l->a[l->n++] = t;
}
static Tray* tl_pop0(TrayList* l) {
    if (l->n == 0) return NULL;
    Tray* t = l->a[0];
    memmove(&l->a[0], &l->a[1], (size_t)(l->n - 1) * sizeof(Tray*));
    l->n--;
    return t;
}
/* models/Tray.py:44-51 / :35-42 */
static int tray_needs_processing(const Tray* t) {
    for (int i = 0; i < t->n; i++)
        if (!t->products[i]->is_processed) return 1;
    return 0;
}
static int tray_needs_packaging(const Tray* t) {
    for (int i = 0; i < t->n; i++)
        if (!t->products[i]->is_packaged) return 1;
    return 0;
}
static int tray_type(const Tray* t) { return t->n ? t->products[0]->type : 0; }

/* AGVAgent._get_current_location (AGVAgent.py:398-403): first LocationType whose cell matches */
static int agv_location(const OracleEnv* e, const Cell* c) {
    for (int l = 0; l < FJSP_NUM_LOCATIONS; l++)
        if (e->cfg.pos[l][0] == c->agv_row && e->cfg.pos[l][1] == c->agv_col) return l;
    return LOC_NONE;
}

int fjsp_oracle_default_config(FjspConfig* c) {
    memset(c, 0, sizeof(*c));
    c->struct_size = (int32_t)sizeof(FjspConfig);
    /* constants.py:5-11 */
    c->pos[LOC_PICKUP][0] = 0, c->pos[LOC_PICKUP][1] = 0;
    c->pos[LOC_BIG][0] = 0, c->pos[LOC_BIG][1] = 3;
    c->pos[LOC_SMALL][0] = 2, c->pos[LOC_SMALL][1] = 3;
    c->pos[LOC_STORAGE][0] = 3, c->pos[LOC_STORAGE][1] = 0;
    c->pos[LOC_PACKAGING][0] = 3, c->pos[LOC_PACKAGING][1] = 5;
    c->grid_rows = 4, c->grid_cols = 6;
    c->proc_small = 60, c->proc_big = 120, c->proc_pack = 30; /* constants.py:14-18 */
    c->step_size = 10, c->agv_speed = 1, c->max_episode_steps = 200;
    c->storage_capacity = 100, c->pack_capacity = 20, c->tray_capacity = 5, c->num_trays = 1000;
    c->num_cells = 1;
    return 0;
}

OracleEnv* fjsp_oracle_create(const FjspConfig* cfg) {
    OracleEnv* e = (OracleEnv*)calloc(1, sizeof(OracleEnv));
    if (!e) return NULL;
    if (cfg) e->cfg = *cfg; else fjsp_oracle_default_config(&e->cfg);
    e->small_steps = e->cfg.proc_small / e->cfg.step_size;
    e->big_steps = e->cfg.proc_big / e->cfg.step_size;
    e->pack_steps = e->cfg.proc_pack / e->cfg.step_size;
    return e;
}
void fjsp_oracle_destroy(OracleEnv* e) { free(e); }

/* FJSPSimulation.reset (FJSPSimulation.py:286-323) with an explicit order table in place of the
 * three NumPy draws of generate_order (:107-112). */
void fjsp_oracle_reset(OracleEnv* e, const FjspOrderRec* orders, int num_orders) {
    FjspConfig cfg = e->cfg;
    int ss = e->small_steps, bs = e->big_steps, ps = e->pack_steps;
    memset(e, 0, sizeof(*e));
    e->cfg = cfg, e->small_steps = ss, e->big_steps = bs, e->pack_steps = ps;
    /* _init_agents (:62-82), once per cell */
    e->ncells = cfg.num_cells < 1 ? 1 : (cfg.num_cells > FJSP_MAX_CELLS ? FJSP_MAX_CELLS : cfg.num_cells);
    e->nagv = 1;
    if (cfg.shared_agvs >= 2) { /* shared floor: one set of stations, shared_agvs AGVs (AGV j's fields in cells[j]) */
        static const int start_loc[FJSP_MAX_SHARED_AGVS] = {LOC_PICKUP, LOC_STORAGE, LOC_SMALL, LOC_BIG};
        e->ncells = 1;
        e->nagv = cfg.shared_agvs > FJSP_MAX_SHARED_AGVS ? FJSP_MAX_SHARED_AGVS : cfg.shared_agvs;
        for (int j = 1; j < e->nagv; j++) e->cells[j].agv_row = cfg.pos[start_loc[j]][0], e->cells[j].agv_col = cfg.pos[start_loc[j]][1];
    }
    for (int ci = 0; ci < e->ncells; ci++) {
        Cell* c = &e->cells[ci];
        /* AGVAgent.py:41: the AGV starts at PICKUP; the dock holds one AGV, so AGVs of further cells start at STORAGE */
        const int start = ci == 0 ? LOC_PICKUP : LOC_STORAGE;
        c->agv_row = cfg.pos[start][0], c->agv_col = cfg.pos[start][1];
        c->holds_dock = ci == 0;
        c->machine[0].proc_steps = ss;
        c->machine[1].proc_steps = bs;
        c->pack[0].colour = COL_BLUE, c->pack[1].colour = COL_BLUE; /* FJSPSimulation.py:68-73 */
        c->pack[2].colour = COL_RED, c->pack[3].colour = COL_GREEN;
    }
    /* _init_trays (:89-98): ids 0..num_trays-1, popped from the END into trays_at_station, at most 1000 */
    e->ntrays_total = cfg.num_trays < 1000 ? cfg.num_trays : 1000; /* tray i (allocation order) has id num_trays-1-i */
    /* generate_order x num_orders (:101-131, :315-318) */
    if (num_orders > ORC_MAX_ORDERS) num_orders = ORC_MAX_ORDERS;
    for (int o = 0; o < num_orders; o++) {
        uint32_t r = orders[o];
        int n = (int)(r & 0xff), type = (int)((r >> 8) & 0xff), colour = (int)((r >> 16) & 0xff);
        Order* od = &e->orders[o];
        od->id = o, od->n = n, od->is_complete = 0, od->completion_step = -1;
        for (int i = 0; i < n; i++) {
            Prod* p = &e->prods[e->nprods++];
            p->id = o * 100 + i; /* :115 */
            p->type = type, p->colour = colour, p->order = o, p->idx = i;
            p->is_processed = 0, p->is_packaged = 0;
            od->products[i] = p;
        }
    }
    e->norders = num_orders;
    e->norders_table = num_orders;
}

/* Long order streams with arrivals (builder-defined extension, include/fjsp_b200.h): `orders` holds the attributes of all
 * num_table orders that can ever exist, num_initial of them exist at reset; one more arrives per step with probability
 * arrival_prob_q16 / 65536 from Philox(key = seed, counter = (genv, episode, step, 5)) until arrival_max_orders exist. */
void fjsp_oracle_reset_stream(OracleEnv* e, const FjspOrderRec* orders, int num_table, int num_initial, uint64_t seed, uint64_t genv,
                              uint32_t episode) {
    fjsp_oracle_reset(e, orders, num_table);
    e->norders = num_initial < e->norders_table ? num_initial : e->norders_table;
    e->seed = seed, e->genv = genv, e->episode = episode;
}

/* ------------------------------------------------------------------ observations (R9) */
static int order_queue_len(const OracleEnv* e) { return e->norders - e->order_queue_head; }
static int trays_at_station(const OracleEnv* e) { return e->ntrays_total - e->trays_next; }
static int pack_has_capacity(const OracleEnv* e, const Pack* p) { return p->users < e->cfg.pack_capacity; } /* PackagingAgent.py:149-153 */

/* scaled shop: the pickup station has ONE dock; another cell's AGV standing there or under way to it blocks it */
static int dock_blocked_for(const OracleEnv* e, const Cell* c) {
    for (int ci = 0; ci < e->ncells; ci++)
        if (&e->cells[ci] != c && e->cells[ci].holds_dock) return 1;
    return 0;
}

/* shared floor: station position `loc` is taken by another AGV (standing there, or under way to it) */
static int loc_taken_by_other(const OracleEnv* e, const Cell* c, int loc) {
    for (int j = 0; j < e->nagv; j++) {
        const Cell* o = &e->cells[j];
        if (o == c) continue;
        const int at = (o->is_moving || o->move_created_now) ? o->move_target_loc : agv_location(e, o);
        if (at == loc) return 1;
    }
    return 0;
}

/* one cell's block of the observation: AGV (13) + small/big machine (3 + 3) + four packaging stations (12) = 31 floats,
 * and of the masks: 8 + 3 + 3 + 12 = 26 bytes */
/* the AGV block (13 floats, 8 mask bytes) of AGV `c` working on the stations of `st` (st == c except on the shared floor) */
static void observe_agv(const OracleEnv* e, const Cell* c, const Cell* st, float* obs, int8_t* masks) {
    /* --- AGV: AGVAgent.get_observation (:53-76), keys sorted --- */
    const Machine* sm = &st->machine[0];
    const Machine* bm = &st->machine[1];
    const Tray* tr = c->carrying;
    obs[0] = (float)bm->is_busy;
    obs[1] = (float)bm->ready.n;
    obs[2] = (float)(tr ? 1 : 0);
    obs[3] = (float)e->ps_ready.n;
    obs[4] = (float)c->agv_row, obs[5] = (float)c->agv_col;
    obs[6] = (float)sm->is_busy;
    obs[7] = (float)sm->ready.n;
    obs[8] = (float)st->storage.n;
    obs[9] = (float)(tr && tray_needs_packaging(tr));
    obs[10] = (float)(tr && tray_needs_processing(tr));
    obs[11] = (float)(tr ? tr->n : 0);
    obs[12] = (float)(tr ? tray_type(tr) : 0);
    /* get_action_mask (:79-178) */
    {
        int8_t* m = masks;
        m[0] = 1;
        if (!c->is_moving) {
            int loc = agv_location(e, c);
            static const int move_loc[6] = {-1, LOC_PICKUP, LOC_SMALL, LOC_BIG, LOC_STORAGE, LOC_PACKAGING};
            for (int a = 1; a <= 5; a++) m[a] = (int8_t)(loc != move_loc[a]);
            if (e->nagv > 1) { /* shared floor: every station position holds one AGV */
                for (int a = 1; a <= 5; a++)
                    if (m[a] && loc_taken_by_other(e, c, move_loc[a])) m[a] = 0;
            } else if (m[1] && dock_blocked_for(e, c)) m[1] = 0; /* scaled shop only: the dock is taken */
            if (tr == NULL && loc != LOC_NONE) {
                if (loc == LOC_PICKUP) m[6] = (int8_t)(e->ps_ready.n > 0);
                else if (loc == LOC_SMALL) m[6] = (int8_t)(sm->ready.n > 0);
                else if (loc == LOC_BIG) m[6] = (int8_t)(bm->ready.n > 0);
                else if (loc == LOC_STORAGE) m[6] = (int8_t)(st->storage.n > 0);
            } else if (tr != NULL && loc != LOC_NONE) {
                int ty = tray_type(tr);
                if (loc == LOC_PICKUP) m[7] = (int8_t)(tr->n == 0);
                else if (loc == LOC_SMALL) m[7] = (int8_t)(tray_needs_processing(tr) && (ty == TYPE_SMALL || ty == TYPE_MEDIUM));
                else if (loc == LOC_BIG) m[7] = (int8_t)(tray_needs_processing(tr) && (ty == TYPE_BIG || ty == TYPE_MEDIUM));
                else if (loc == LOC_PACKAGING) m[7] = (int8_t)(tray_needs_packaging(tr) && !tray_needs_processing(tr));
                else if (loc == LOC_STORAGE) m[7] = 1;
            }
        }
    }
}

/* machines (6 floats, 6 mask bytes) and packaging stations (12, 12) of one cell */
static void observe_stations(const OracleEnv* e, const Cell* c, float* obs, int8_t* masks) {
    /* --- machines: MachineAgent.get_observation (:62-70), get_action_mask (:72-97) --- */
    for (int i = 0; i < 2; i++) {
        const Machine* m = &c->machine[i];
        float* o = obs + 3 * i;
        int8_t* k = masks + 3 * i;
        o[0] = (float)m->is_busy;
        o[1] = m->progress_done ? 1.0f : 0.0f;
        o[2] = (float)(int8_t)m->queue.n; /* dtype=np.int8, :67 */
        k[0] = 1;
        k[1] = (int8_t)(m->queue.n > 0 && !m->is_busy);
        k[2] = (int8_t)(!m->is_busy && m->current_tray != NULL);
    }
    /* --- packaging: PackagingAgent.get_observation (:54-62), get_action_mask (:64-89) --- */
    for (int i = 0; i < 4; i++) {
        const Pack* p = &c->pack[i];
        float* o = obs + 6 + 3 * i;
        int8_t* k = masks + 6 + 3 * i;
        o[0] = (float)p->is_busy;
        o[1] = (float)p->progress; /* np.array(double, dtype=np.float32) */
        o[2] = (float)(int8_t)p->qn;
        k[0] = 1;
        k[1] = (int8_t)(p->qn > 0 && !p->is_busy && pack_has_capacity(e, p));
        k[2] = (int8_t)(!p->is_busy && p->current_product != NULL);
    }
}

static void observe_cell(const OracleEnv* e, const Cell* c, float* obs, int8_t* masks) {
    observe_agv(e, c, c, obs, masks);
    observe_stations(e, c, obs + 13, masks + 8);
}

void fjsp_oracle_observe(const OracleEnv* e, float* obs, int8_t* masks) {
    memset(masks, 0, (size_t)(e->nagv > 1 ? FJSP_SHARED_MASK_DIM(e->nagv) : FJSP_MASK_DIM_K(e->ncells)));
    /* --- pickup station: PickupStationAgent.get_observation (:58-98), keys sorted --- */
    int order_size = 0, remaining = 0, next_type = 0, next_colour = 0;
    if (e->current_order) {
        order_size = e->current_order->n;
        remaining = order_size - e->current_order_product_idx;
        if (remaining > 0) {
            const Prod* np_ = e->current_order->products[e->current_order_product_idx];
            next_type = np_->type, next_colour = np_->colour;
        }
    }
    int t_type = 0, t_colour = 0, t_count = 0;
    if (e->current_tray) {
        t_count = e->current_tray->n;
        if (t_count > 0) t_type = e->current_tray->products[0]->type, t_colour = e->current_tray->products[0]->colour;
    }
    obs[0] = (float)t_colour, obs[1] = (float)t_count, obs[2] = (float)t_type;
    obs[3] = (float)next_colour, obs[4] = (float)next_type, obs[5] = (float)order_size, obs[6] = (float)remaining;
    /* get_action_mask (:100-142) */
    {
        int has_order = e->current_order != NULL || order_queue_len(e) > 0;
        int has_tray = e->current_tray != NULL || trays_at_station(e) > 0;
        int tray_not_full = 1;
        if (e->current_tray) tray_not_full = e->current_tray->n < FJSP_TRAY_CAPACITY; /* global CONFIG, :125 */
        int prem = 0;
        if (e->current_order) prem = e->current_order_product_idx < e->current_order->n;
        else if (order_queue_len(e) > 0) prem = 1;
        masks[0] = 1;
        masks[1] = (int8_t)(has_order && has_tray && tray_not_full && prem);
        masks[2] = (int8_t)(e->current_tray && e->current_tray->n > 0);
    }
    if (e->nagv > 1) { /* shared floor: pickup station | AGVs | machines | packaging stations */
        for (int j = 0; j < e->nagv; j++) observe_agv(e, &e->cells[j], &e->cells[0], obs + 7 + 13 * j, masks + 3 + 8 * j);
        observe_stations(e, &e->cells[0], obs + 7 + 13 * e->nagv, masks + 3 + 8 * e->nagv);
        return;
    }
    for (int ci = 0; ci < e->ncells; ci++) observe_cell(e, &e->cells[ci], obs + 7 + 31 * ci, masks + 3 + 26 * ci);
}

/* ------------------------------------------------------------------ action phase */
/* PickupStationAgent.execute_action (:146-232) -> local reward per RewardModel.py:53-60 */
static double act_pickup(OracleEnv* e, int action, uint8_t* res) {
    int loaded = 0, tray_done = 0, idle_orders = 0, success = 0;
    if (action == 0) {
        idle_orders = order_queue_len(e) > 0 || e->current_order != NULL;
        success = 1;
    } else if (action == 1) {
        do {
            if (!e->current_order) {
                if (order_queue_len(e) > 0) {
                    e->current_order = &e->orders[e->order_queue_head++]; /* pop(0), :169 */
                    e->current_order_product_idx = 0;
                } else break;
            }
            if (!e->current_tray) {
                if (trays_at_station(e) > 0) {
                    int ti = e->trays_next++; /* trays_at_station.pop(0), :176 */
                    e->current_tray = &e->trays[ti % MAX_TRAYS];
                    e->current_tray->id = e->cfg.num_trays - 1 - ti;
                    e->current_tray->n = 0;
                    e->current_tray->order_id = e->current_order->id;
                } else break;
            }
            Prod* p = e->current_order->products[e->current_order_product_idx];
            if (e->current_tray->n < e->cfg.tray_capacity) { /* not is_full(), :186 */
                if (e->current_order->id != e->current_tray->order_id) { /* :187-191 (unreachable) */
                    tl_push(&e->ps_ready, e->current_tray);
                    e->current_tray = NULL;
                    tray_done = 1;
                    break;
                }
                e->current_tray->products[e->current_tray->n++] = p;
                e->current_order_product_idx++;
                loaded = 1, success = 1;
                if (e->current_order_product_idx >= e->current_order->n) { /* :201-208 */
                    e->current_order = NULL;
                    e->current_order_product_idx = 0;
                    tl_push(&e->ps_ready, e->current_tray);
                    e->current_tray = NULL;
                    tray_done = 1;
                    break;
                }
                if (e->current_tray->n >= e->cfg.tray_capacity) { /* :211-216 */
                    tl_push(&e->ps_ready, e->current_tray);
                    e->current_tray = NULL;
                    tray_done = 1;
                    break;
                }
            } else { /* :217-223 (unreachable) */
                tl_push(&e->ps_ready, e->current_tray);
                e->current_tray = NULL;
                tray_done = 1;
            }
        } while (0);
    } else if (action == 2) {
        if (e->current_tray && e->current_tray->n > 0) { /* :226-230 */
            tl_push(&e->ps_ready, e->current_tray);
            e->current_tray = NULL;
            success = 1;
        }
    }
    *res = (uint8_t)((success ? FJSP_RES_SUCCESS : 0) | (loaded ? FJSP_RES_PS_LOADED : 0) |
                     (tray_done ? FJSP_RES_PS_TRAY_DONE : 0) | (idle_orders ? FJSP_RES_PS_IDLE_ORDERS : 0));
    double r = 0.0;
    if (loaded) r += 1.0;
    if (tray_done) r += 5.0;
    if (action == 0 && idle_orders) r += -1.0;
    return r;
}

/* FJSPSimulation.add_tray_to_packaging (:402-430) + PackagingAgent.add_tray (:127-131) */
static void add_tray_to_packaging(OracleEnv* e, Cell* c, Tray* t) {
    static const char* station_colour_name[4] = {"blue", "blue", "red", "green"};
    static const char* colour_name[4] = {"", "red", "blue", "green"};
    int colour = t->products[0]->colour;
    for (int s = 0; s < 4; s++) {
        Pack* p = &c->pack[s];
        /* `packaging_color.name.lower() in station.color.name.lower()` — substring test, equal names here */
        if (strstr(station_colour_name[s], colour_name[colour]) && pack_has_capacity(e, p)) {
            for (int i = 0; i < t->n; i++)
                if (t->products[i]->colour == p->colour && p->qn < LIST_CAP) p->queue[p->qn++] = t->products[i];
            return;
        }
    }
    /* no station with capacity: products silently dropped (:426-427) */
}

/* AGVAgent.execute_action (:180-252), _execute_pickup (:254-293), _execute_drop (:295-368) */
/* `st` = the cell whose stations the AGV works on (st == c except on the shared floor, where it is cells[0]) */
static double act_agv(OracleEnv* e, Cell* c, Cell* st, int action, uint8_t* res) {
    int invalid = 0, moved = 0, pickup_ok = 0, drop_ok = 0, to_pack = 0, success = 0;
    if (c->is_moving) {
        invalid = 1; /* :210-212 */
    } else if (action == 0) {
        success = 1;
    } else if (action >= 1 && action <= 5) {
        static const int move_loc[6] = {-1, LOC_PICKUP, LOC_SMALL, LOC_BIG, LOC_STORAGE, LOC_PACKAGING}; /* :218-224 */
        int tl = move_loc[action];
        int d = abs(c->agv_row - e->cfg.pos[tl][0]) + abs(c->agv_col - e->cfg.pos[tl][1]);
        if (e->nagv > 1 && d != 0 && loc_taken_by_other(e, c, tl)) {
            invalid = 1; /* shared floor: the position is occupied by (or granted to) another AGV */
        } else if (e->nagv == 1 && tl == LOC_PICKUP && d != 0 && dock_blocked_for(e, c)) {
            invalid = 1; /* scaled shop only: the single dock of the pickup station is taken by another cell's AGV */
        } else {
        success = 1;
        if (d != 0) {
            if (tl == LOC_PICKUP) c->holds_dock = 1; /* granted now: later AGVs of this step already see it */
            /* env.process(_move_process) (:236): Initialize is URGENT in this step's run */
            c->move_created_now = 1;
            c->move_target_loc = tl;
            /* timeout(d / agv_speed) from t = step*step_size fires in the run of step + floor(d/(speed*step_size)) (R0) */
            c->move_arrive_step = e->current_step + d / (e->cfg.agv_speed * e->cfg.step_size);
            moved = 1;
        }
        }
    } else if (action == 6) {
        int loc = agv_location(e, c);
        if (c->carrying != NULL || loc == LOC_NONE || loc == LOC_PACKAGING) {
            invalid = 1;
        } else {
            Tray* t = NULL;
            if (loc == LOC_PICKUP && e->cfg.long_streams && e->ps_ready.n > 0 && !e->orders[e->ps_ready.a[0]->order_id].in_process) {
                /* capacity of the port's long layout (not reference behaviour): FJSP_LONG_ORDER_SLOTS orders can be in process
                 * at once; an order leaves one step after it completes */
                int active = 0;
                for (int o = 0; o < e->order_queue_head; o++)
                    active += e->orders[o].in_process && !(e->orders[o].is_complete && e->orders[o].completion_step < e->current_step);
                if (active >= FJSP_LONG_ORDER_SLOTS) {
                    e->fault = FJSP_FAULT_ORDER_SLOTS;
                    loc = LOC_NONE; /* -> invalid_action below, the tray stays where it is */
                } else {
                    e->orders[e->ps_ready.a[0]->order_id].in_process = 1;
                }
            }
            if (loc == LOC_PICKUP) t = tl_pop0(&e->ps_ready);
            else if (loc == LOC_SMALL) t = tl_pop0(&st->machine[0].ready);
            else if (loc == LOC_BIG) t = tl_pop0(&st->machine[1].ready);
            else if (loc == LOC_STORAGE) t = tl_pop0(&st->storage);
            if (t) c->carrying = t, success = 1, pickup_ok = 1;
            else invalid = 1;
        }
    } else if (action == 7) {
        int loc = agv_location(e, c);
        Tray* t = c->carrying;
        if (t == NULL || loc == LOC_NONE) {
            invalid = 1;
        } else if (loc == LOC_PICKUP) {
            if (t->n == 0) { /* returning an empty tray (:310-314) — never happens, trays are never emptied */
                drop_ok = 1;
            } else invalid = 1;
        } else if (loc == LOC_SMALL || loc == LOC_BIG) {
            int ty = tray_type(t);
            int compatible = (loc == LOC_SMALL) ? (ty == TYPE_SMALL || ty == TYPE_MEDIUM) : (ty == TYPE_BIG || ty == TYPE_MEDIUM);
            if (tray_needs_processing(t) && compatible) {
                tl_push(&st->machine[loc == LOC_SMALL ? 0 : 1].queue, t); /* add_tray, MachineAgent.py:141-143 */
                drop_ok = 1;
            } else invalid = 1;
        } else if (loc == LOC_STORAGE) {
            if (st->storage.n < e->cfg.storage_capacity) tl_push(&st->storage, t); /* Storage.add_tray (:16-22); False ignored */
            drop_ok = 1;
        } else if (loc == LOC_PACKAGING) {
            if (tray_needs_packaging(t) && !tray_needs_processing(t)) {
                add_tray_to_packaging(e, st, t);
                drop_ok = 1, to_pack = 1;
            } else invalid = 1;
        }
        if (drop_ok) c->carrying = NULL, success = 1;
    } else {
        invalid = 1; /* :249-250 */
    }
    *res = (uint8_t)((success ? FJSP_RES_SUCCESS : 0) | (invalid ? FJSP_RES_AGV_INVALID : 0) | (moved ? FJSP_RES_AGV_MOVED : 0) |
                     (pickup_ok ? FJSP_RES_AGV_PICKUP : 0) | (drop_ok ? FJSP_RES_AGV_DROP : 0) | (to_pack ? FJSP_RES_AGV_TO_PACK : 0));
    double r = 0.0; /* RewardModel.py:62-77, same accumulation order */
    if (pickup_ok) r += 2.0;
    if (drop_ok) r += 2.0;
    if (to_pack) r += 10.0;
    if (moved) r += -0.1;
    if (invalid) r += -5.0;
    return r;
}

/* MachineAgent.execute_action (:99-139) */
static double act_machine(OracleEnv* e, Machine* m, int action, uint8_t* res) {
    (void)e;
    int started = 0, completed = 0, idle_q = 0, success = 0;
    if (action == 0) {
        idle_q = m->queue.n > 0 && !m->is_busy;
        success = 1;
    } else if (action == 1) {
        if (m->queue.n > 0 && !m->is_busy) {
            m->proc_tray = tl_pop0(&m->queue); /* env.process(...) :121 */
            m->proc_created_now = 1;
            started = 1, success = 1;
        }
    } else if (action == 2) {
        if (!m->is_busy && m->current_tray) {
            tl_push(&m->ready, m->current_tray);
            m->current_tray = NULL;
            completed = 1, success = 1;
        }
    }
    *res = (uint8_t)((success ? FJSP_RES_SUCCESS : 0) | (started ? FJSP_RES_M_STARTED : 0) |
                     (completed ? FJSP_RES_M_COMPLETED : 0) | (idle_q ? FJSP_RES_M_IDLE_QUEUE : 0));
    double r = 0.0; /* RewardModel.py:79-86 */
    if (started) r += 1.0;
    if (completed) r += 5.0;
    if (action == 0 && idle_q) r += -2.0;
    return r;
}

/* PackagingAgent.execute_action (:91-125) */
static double act_pack(OracleEnv* e, Pack* p, int action, uint8_t* res) {
    int started = 0, completed = 0, idle_q = 0, success = 0;
    if (action == 0) {
        idle_q = p->qn > 0 && !p->is_busy;
        success = 1;
    } else if (action == 1) {
        if (p->nput > 0 && p->qn > 0) e->fault = FJSP_FAULT_PKG_RESTART_WITH_WAITERS; /* reference raises later (R-PKG-cap-b) */
        for (int i = 0; i < p->qn; i++) { /* one process per queued product, no busy/capacity check (:113-118) */
            p->created[p->ncreated++] = p->queue[i];
            started = 1, success = 1;
            p->progress = (1.0 / (double)p->qn) * 100.0; /* i is never incremented (:112,117) */
        }
    } else if (action == 2) {
        if (!p->is_busy && p->current_product) completed = 1; /* :121-123, success stays False */
    }
    *res = (uint8_t)((success ? FJSP_RES_SUCCESS : 0) | (started ? FJSP_RES_M_STARTED : 0) |
                     (completed ? FJSP_RES_M_COMPLETED : 0) | (idle_q ? FJSP_RES_M_IDLE_QUEUE : 0));
    double r = 0.0; /* RewardModel.py:88-95 */
    if (started) r += 2.0;
    if (completed) r += 20.0;
    if (action == 0 && idle_q) r += -1.0;
    return r;
}

/* ------------------------------------------------------------------ run phase: env.run(until=now+step_size) */
static void pack_begin(OracleEnv* e, Pack* p, Prod* prod) {
    /* body of _simpy_packaging_process after `yield req` (PackagingAgent.py:138-141) */
    p->is_busy = 1;
    p->current_product = prod;
    int found = -1;
    for (int i = 0; i < p->qn; i++)
        if (p->queue[i] == prod) { found = i; break; }
    if (found < 0) {
        e->fault = FJSP_FAULT_PKG_RESTART_WITH_WAITERS; /* list.remove(x): x not in list */
    } else {
        memmove(&p->queue[found], &p->queue[found + 1], (size_t)(p->qn - 1 - found) * sizeof(Prod*));
        p->qn--;
    }
    if (p->nrunning < 64) {
        p->running[p->nrunning].prod = prod;
        p->running[p->nrunning].finish_step = e->current_step + e->pack_steps;
        p->nrunning++;
    }
}

static void run_pack(OracleEnv* e, Pack* p) {
    int s = e->current_step;
    Prod* granted[LIST_CAP];
    int ngranted = 0;
    /* 1. URGENT Initialize events of this step's processes: request() appends to put_queue and
     *    _trigger_put examines the HEAD once (Resource._do_put returns None -> break). */
    for (int i = 0; i < p->ncreated; i++) {
        p->put_queue[p->nput++] = p->created[i];
        if (p->users < e->cfg.pack_capacity) {
            p->users++;
            granted[ngranted++] = p->put_queue[0];
            memmove(&p->put_queue[0], &p->put_queue[1], (size_t)(p->nput - 1) * sizeof(Prod*));
            p->nput--;
        }
    }
    p->ncreated = 0;
    /* 2. NORMAL timeouts due exactly at the boundary (scheduled pack_steps runs ago, oldest eids) */
    int releases = 0, w = 0;
    for (int i = 0; i < p->nrunning; i++) {
        if (p->running[i].finish_step == s) {
            Prod* prod = p->running[i].prod; /* PackagingAgent.py:143-147 */
            prod->is_packaged = 1;
            p->products_completed++;
            e->total_products_packaged++;
            p->is_busy = 0;
            p->users--; /* Request.__exit__ -> release: users.remove now, Release event queued */
            releases++;
        } else {
            p->running[w++] = p->running[i];
        }
    }
    p->nrunning = w;
    /* 3. grant events of step 1 (NORMAL, scheduled after the old timeouts) */
    for (int i = 0; i < ngranted; i++) pack_begin(e, p, granted[i]);
    /* 4. Release events: each callback _trigger_put grants at most the head waiter */
    ngranted = 0;
    for (int i = 0; i < releases; i++) {
        if (p->nput > 0 && p->users < e->cfg.pack_capacity) {
            p->users++;
            granted[ngranted++] = p->put_queue[0];
            memmove(&p->put_queue[0], &p->put_queue[1], (size_t)(p->nput - 1) * sizeof(Prod*));
            p->nput--;
        }
    }
    /* 5. their grant events */
    for (int i = 0; i < ngranted; i++) pack_begin(e, p, granted[i]);
}

static void run_machine(OracleEnv* e, Machine* m) {
    int s = e->current_step;
    if (m->proc_created_now) {
        /* Initialize -> request granted (capacity 1, free) -> MachineAgent.py:159-161 */
        m->proc_created_now = 0;
        m->is_busy = 1;
        m->current_tray = m->proc_tray; /* overwrites an unsignalled finished tray (:160) */
        m->proc_next_idx = 0;
        m->proc_next_fire = s + m->proc_steps;
    } else if (m->proc_tray && m->proc_next_fire == s) {
        m->proc_tray->products[m->proc_next_idx++]->is_processed = 1; /* :165-166 */
        if (m->proc_next_idx >= m->proc_tray->n) {
            m->is_busy = 0; /* :168-169 */
            m->progress_done = 1;
            m->proc_tray = NULL;
        } else {
            m->proc_next_fire = s + m->proc_steps;
        }
    }
}

static void run_agv(OracleEnv* e, Cell* c) {
    if (c->move_created_now) { /* _move_process (:387-396) starts: is_moving = True */
        c->move_created_now = 0;
        c->is_moving = 1;
    }
    if (c->is_moving && c->move_arrive_step == e->current_step) {
        c->agv_row = e->cfg.pos[c->move_target_loc][0];
        c->agv_col = e->cfg.pos[c->move_target_loc][1];
        c->is_moving = 0;
    }
    /* the dock is held while standing at PICKUP or under way to it; it is freed when the AGV starts to leave */
    c->holds_dock = c->is_moving ? (c->move_target_loc == LOC_PICKUP) : (agv_location(e, c) == LOC_PICKUP);
}

/* ------------------------------------------------------------------ step */
void fjsp_oracle_step(OracleEnv* e, const uint8_t* actions, float* obs, int8_t* masks, double* rewards,
                      uint8_t* flags, uint8_t* results) {
    uint8_t res_local[FJSP_ACT_DIM_K(FJSP_MAX_CELLS)];
    uint8_t* res = results ? results : res_local;
    const int K = e->ncells, A = e->nagv > 1 ? FJSP_SHARED_AGENTS(e->nagv) : FJSP_AGENTS_K(K);
    const int act_dim = e->nagv > 1 ? FJSP_SHARED_ACT_DIM(e->nagv) : FJSP_ACT_DIM_K(K);
    int orders_before = e->completed_orders;
    int products_before = e->total_products_packaged;
    double local[FJSP_ACT_DIM_K(FJSP_MAX_CELLS)];
    if (e->current_step > e->cfg.max_episode_steps) {
        /* stepped after the truncation step without a reset: the port's env is inert and says so (the reference would go
         * on simulating; include/fjsp_b200.h FJSP_FAULT_PAST_END) */
        for (int i = 0; i < act_dim; i++) res[i] = 0;
        for (int i = 0; i < A; i++) rewards[i] = 0.0;
        if (obs && masks) fjsp_oracle_observe(e, obs, masks);
        flags[0] = 0, flags[1] = 1, flags[2] = FJSP_FAULT_PAST_END, flags[3] = 0;
        return;
    }
    /* 0. order arrivals (extension, long order streams only) */
    if (e->cfg.long_streams && e->cfg.arrival_prob_q16 > 0 && e->norders < e->cfg.arrival_max_orders && e->norders < e->norders_table) {
        uint32_t ctr[4] = {(uint32_t)e->genv, e->episode, (uint32_t)e->current_step, 5u}, key[2] = {(uint32_t)e->seed, (uint32_t)(e->seed >> 32)}, r[4];
        fjsp_oracle_philox(ctr, key, r);
        if ((r[0] & 0xffffu) < (uint32_t)e->cfg.arrival_prob_q16) e->norders++;
    }
    /* 1. actions in dict order (FJSPSimulation.py:172-174, :76-82): pickup station, then cell by cell agv, small
     *    machine, big machine, four packaging stations (one cell = the reference's order) */
    local[0] = act_pickup(e, actions[0], &res[0]);
    if (e->nagv > 1) { /* shared floor: pickup station, the AGVs in index order, machines, packaging stations */
        Cell* st = &e->cells[0];
        const int n = e->nagv;
        for (int i = A; i < act_dim; i++) res[i] = 0;
        for (int j = 0; j < n; j++) local[1 + j] = act_agv(e, &e->cells[j], st, actions[1 + j], &res[1 + j]);
        local[1 + n] = act_machine(e, &st->machine[0], actions[1 + n], &res[1 + n]);
        local[2 + n] = act_machine(e, &st->machine[1], actions[2 + n], &res[2 + n]);
        for (int i = 0; i < 4; i++) local[3 + n + i] = act_pack(e, &st->pack[i], actions[3 + n + i], &res[3 + n + i]);
        for (int j = 0; j < n; j++) run_agv(e, &e->cells[j]);
        run_machine(e, &st->machine[0]);
        run_machine(e, &st->machine[1]);
        for (int i = 0; i < 4; i++) run_pack(e, &st->pack[i]);
    } else {
    for (int ci = 0; ci < K; ci++) {
        Cell* c = &e->cells[ci];
        const int b = 1 + 7 * ci;
        local[b] = act_agv(e, c, c, actions[b], &res[b]);
        local[b + 1] = act_machine(e, &c->machine[0], actions[b + 1], &res[b + 1]);
        local[b + 2] = act_machine(e, &c->machine[1], actions[b + 2], &res[b + 2]);
        for (int i = 0; i < 4; i++) local[b + 3 + i] = act_pack(e, &c->pack[i], actions[b + 3 + i], &res[b + 3 + i]);
    }
    /* 2. env.run(until=now+step_size) (:183-184). Stations do not interact inside a run. */
    for (int ci = 0; ci < K; ci++) {
        Cell* c = &e->cells[ci];
        run_agv(e, c);
        run_machine(e, &c->machine[0]);
        run_machine(e, &c->machine[1]);
        for (int i = 0; i < 4; i++) run_pack(e, &c->pack[i]);
    }
    }
    /* 3. _check_order_completions (:245-258) */
    for (int o = 0; o < e->norders; o++) {
        Order* od = &e->orders[o];
        if (!od->is_complete) {
            int all = 1;
            for (int i = 0; i < od->n; i++) all = all && od->products[i]->is_packaged;
            if (all) {
                od->is_complete = 1;
                od->completion_step = e->current_step;
                e->completed_orders++;
            }
        }
    }
    /* 4. rewards (:190-209, RewardModel.py:34-44,99-110), same double arithmetic order; num_agents = 1 + 7K */
    int oc = e->completed_orders - orders_before;
    int pp = e->total_products_packaged - products_before;
    double g = 100.0 * (double)oc;
    g += 10.0 * (double)pp;
    g += -0.1 * (double)e->cfg.step_size;
    for (int i = 0; i < A; i++) rewards[i] = g / (double)A + local[i];
    /* 5. observations */
    if (obs && masks) fjsp_oracle_observe(e, obs, masks);
    /* 6. termination / truncation (:216-224), pre-increment current_step */
    int all_done = e->completed_orders == e->norders && e->norders > 0 && order_queue_len(e) == 0;
    if (e->cfg.long_streams && e->cfg.arrival_prob_q16 > 0 && e->norders < e->cfg.arrival_max_orders) all_done = 0;
    int truncated = e->current_step >= e->cfg.max_episode_steps;
    flags[0] = (uint8_t)all_done, flags[1] = (uint8_t)truncated, flags[2] = (uint8_t)e->fault, flags[3] = 0;
    e->current_step++;
}

/* ------------------------------------------------------------------ canonical record S */
static int g_long_entries = 0; /* set by the exporter: FJSP_TRAY_ENTRY_LONG for envs with long_streams */
static int32_t tray_entry(const Tray* t) {
    if (!t) return -1;
    if (t->n == 0) return t->id;
    if (g_long_entries) return FJSP_TRAY_ENTRY_LONG(t->id, t->products[0]->order, t->products[0]->idx, t->n);
    return FJSP_TRAY_ENTRY(t->id, t->products[0]->order, t->products[0]->idx, t->n);
}
static int fill_trays(int32_t* dst, int cap, const TrayList* l) {
    for (int i = 0; i < cap; i++) dst[i] = -1;
    for (int i = 0; i < l->n && i < cap; i++) dst[i] = tray_entry(l->a[i]);
    return l->n;
}
static void order_record(const Order* od, int32_t* r4) {
    int pm = 0, km = 0;
    for (int i = 0; i < od->n; i++) pm |= od->products[i]->is_processed << i, km |= od->products[i]->is_packaged << i;
    r4[0] = km, r4[1] = pm, r4[2] = od->is_complete, r4[3] = od->completion_step;
}
/* orders [first, first + count): {packaged_mask, processed_mask, is_complete, completion_step}; with long_streams the
 * port's conventions of fjsp_export_orders: popped orders that left the ring read "complete, all bits", queued ones zeros */
void fjsp_oracle_export_orders(const OracleEnv* e, int first, int count, int32_t* out4) {
    for (int i = 0; i < count; i++) {
        const int o = first + i;
        int32_t* r = out4 + 4 * i;
        r[0] = r[1] = r[2] = 0, r[3] = -1;
        if (o < 0 || o >= e->norders) continue;
        if (e->cfg.long_streams && o >= e->order_queue_head) continue;
        /* the port gives an order's slot back one step after it completes: from then on it reads "complete, all bits" */
        if (e->cfg.long_streams && e->orders[o].is_complete && e->orders[o].completion_step < e->current_step - 1) {
            r[0] = r[1] = 0x1ff, r[2] = 1;
            continue;
        }
        order_record(&e->orders[o], r);
    }
}

void fjsp_oracle_export_cell(const OracleEnv* e, int cell, FjspCanonState* s) {
    const Cell* agv = &e->cells[cell];                        /* shared floor: `cell` = AGV index, the stations are cells[0]'s */
    const Cell* c = e->nagv > 1 ? &e->cells[0] : agv;
    memset(s, 0, sizeof(*s));
    g_long_entries = e->cfg.long_streams != 0;
    s->current_step = e->current_step, s->num_orders = e->norders, s->fault = e->fault;
    s->agv_row = agv->agv_row, s->agv_col = agv->agv_col, s->agv_carry = tray_entry(agv->carrying), s->agv_is_moving = agv->is_moving;
    s->ps_order_queue_len = order_queue_len(e);
    s->ps_current_order = e->current_order ? e->current_order->id : -1;
    s->ps_product_idx = e->current_order_product_idx;
    s->ps_current_tray = tray_entry(e->current_tray);
    s->ps_trays_at_station = trays_at_station(e);
    s->ps_ready_n = fill_trays(s->ps_ready, FJSP_CANON_PS_READY, &e->ps_ready);
    for (int i = 0; i < 2; i++) {
        const Machine* m = &c->machine[i];
        s->machine[i].is_busy = m->is_busy;
        s->machine[i].current_tray = tray_entry(m->current_tray);
        s->machine[i].progress_done = m->progress_done;
        s->machine[i].queue_n = fill_trays(s->machine[i].queue, FJSP_CANON_MAXQ, &m->queue);
        s->machine[i].ready_n = fill_trays(s->machine[i].ready, FJSP_CANON_MAXQ, &m->ready);
    }
    s->storage_n = fill_trays(s->storage, FJSP_CANON_MAXQ, &c->storage);
    for (int i = 0; i < 4; i++) {
        const Pack* p = &c->pack[i];
        s->pack[i].is_busy = p->is_busy;
        s->pack[i].current_product = p->current_product ? p->current_product->id : -1;
        s->pack[i].progress_L = p->progress != 0.0 ? (int32_t)lround(100.0 / p->progress) : 0;
        s->pack[i].products_completed = p->products_completed;
        s->pack[i].users = p->users;
        s->pack[i].queue_n = p->qn;
        for (int k = 0; k < FJSP_CANON_MAXPQ; k++) s->pack[i].queue[k] = k < p->qn ? p->queue[k]->id : -1;
    }
    for (int o = 0; o < FJSP_MAX_ORDERS; o++) s->order_completion_step[o] = -1;
    /* per-order arrays: orders 0..31; long order streams: the 32 most recently popped orders */
    const int base = e->cfg.long_streams ? (e->order_queue_head > FJSP_MAX_ORDERS ? e->order_queue_head - FJSP_MAX_ORDERS : 0) : 0;
    int32_t rec[4 * FJSP_MAX_ORDERS];
    fjsp_oracle_export_orders(e, base, FJSP_MAX_ORDERS, rec);
    for (int i = 0; i < FJSP_MAX_ORDERS; i++)
        s->packaged_mask[i] = rec[4 * i], s->processed_mask[i] = rec[4 * i + 1], s->order_complete[i] = rec[4 * i + 2],
        s->order_completion_step[i] = rec[4 * i + 3];
    s->total_products_packaged = e->total_products_packaged;
    s->completed_orders = e->completed_orders;
}
void fjsp_oracle_export(const OracleEnv* e, FjspCanonState* s) { fjsp_oracle_export_cell(e, 0, s); }

/* ------------------------------------------------------------------ Philox4x32-10 streams (replayable) */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}
void fjsp_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}
static uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* orders: counter = (global env, episode, order, 0) */
void fjsp_oracle_philox_orders(uint64_t seed, uint64_t genv, uint32_t episode, int num_orders, FjspOrderRec* out) {
    for (int o = 0; o < num_orders; o++) {
        uint32_t r[4];
        philox4x32_10((uint32_t)genv, episode, (uint32_t)o, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        out[o] = FJSP_ORDER_REC(1 + mulhi32(r[0], 9), 1 + mulhi32(r[1], 3), 1 + mulhi32(r[2], 3));
    }
}
/* actions: counter = (global env, t_lo, t_hi, 1); a_j = (u16_j * n_j) >> 16 */
void fjsp_oracle_philox_actions(uint64_t seed, uint64_t genv, uint64_t t, uint8_t* a) {
    static const uint32_t nact[8] = {3, 8, 3, 3, 3, 3, 3, 3};
    uint32_t r[4];
    philox4x32_10((uint32_t)genv, (uint32_t)t, (uint32_t)(t >> 32), 1u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    for (int j = 0; j < 8; j++) {
        uint32_t h = (j & 1) ? (r[j >> 1] >> 16) : (r[j >> 1] & 0xffffu);
        a[j] = (uint8_t)((h * nact[j]) >> 16);
    }
}

/* scaled shop: cell c draws its 8 values with counter word 3 = 1 + 16c; cell 0 also supplies the pickup station */
void fjsp_oracle_philox_actions_k(uint64_t seed, uint64_t genv, uint64_t t, int K, uint8_t* a) {
    static const uint32_t nact[8] = {3, 8, 3, 3, 3, 3, 3, 3};
    for (int c = 0; c < K; c++) {
        uint32_t r[4];
        philox4x32_10((uint32_t)genv, (uint32_t)t, (uint32_t)(t >> 32), 1u + 16u * (uint32_t)c, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        for (int j = (c == 0 ? 0 : 1); j < 8; j++) {
            uint32_t h = (j & 1) ? (r[j >> 1] >> 16) : (r[j >> 1] & 0xffffu);
            a[7 * c + j] = (uint8_t)((h * nact[j]) >> 16);
        }
    }
}

/* shared floor: the reference stream's eight values serve (pickup station, agv_0, machines, stations); AGVs 1.. draw from
 * counter word 3 = 17 (u16 lane j - 1) */
void fjsp_oracle_philox_actions_shared(uint64_t seed, uint64_t genv, uint64_t t, int nagv, uint8_t* a) {
    uint8_t base[8];
    uint32_t r[4];
    fjsp_oracle_philox_actions(seed, genv, t, base);
    philox4x32_10((uint32_t)genv, (uint32_t)t, (uint32_t)(t >> 32), 17u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    for (int i = 0; i < FJSP_SHARED_ACT_DIM(nagv); i++) a[i] = 0;
    a[0] = base[0], a[1] = base[1];
    for (int j = 1; j < nagv; j++) {
        uint32_t h = ((j - 1) & 1) ? (r[(j - 1) >> 1] >> 16) : (r[(j - 1) >> 1] & 0xffffu);
        a[1 + j] = (uint8_t)((h * 8u) >> 16);
    }
    for (int i = 0; i < 6; i++) a[1 + nagv + i] = base[2 + i];
}

/* ------------------------------------------------------------------ batch rollout (CPU baseline + full-size statistics parity)
 * n_envs independent envs, `steps` lockstep steps, Philox actions at t = t0.., Philox orders, auto-reset on
 * terminated|truncated|fault.  stats[8] += {env_steps, episodes, orders_completed, products_packaged, faults,
 * sum(round(10*A*reward)) over all A = 1+7K agents (80*reward for the reference shop: rewards are multiples of 1/(10A)), 0, 0}.  Env state persists in `envs` (array of n_envs OracleEnv*). */
int64_t fjsp_oracle_env_size(void) { return (int64_t)sizeof(OracleEnv); }

typedef struct RolloutJob {
    OracleEnv** envs;
    uint32_t* episodes;
    int64_t lo, hi, first_env;
    int steps, num_orders;
    uint64_t seed, t0;
    uint64_t acc[8];
    int64_t rsum;
} RolloutJob;

static void* rollout_worker(void* arg) {
    RolloutJob* j = (RolloutJob*)arg;
    uint64_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; /* thread-local: the job structs share cache lines */
    int64_t rsum = 0;
    for (int64_t i = j->lo; i < j->hi; i++) {
        OracleEnv* e = j->envs[i];
        uint64_t genv = (uint64_t)(j->first_env + i);
        float obs[FJSP_OBS_DIM_K(FJSP_MAX_CELLS)];
        int8_t masks[FJSP_MASK_DIM_K(FJSP_MAX_CELLS)];
        double rew[FJSP_ACT_DIM_K(FJSP_MAX_CELLS)];
        uint8_t flags[4], act[FJSP_ACT_DIM_K(FJSP_MAX_CELLS)];
        uint32_t episode = j->episodes[i];
        const int K = e->ncells, A = FJSP_AGENTS_K(K);
        for (int k = 0; k < j->steps; k++) {
            int before_o = e->completed_orders, before_p = e->total_products_packaged;
            fjsp_oracle_philox_actions_k(j->seed, genv, j->t0 + (uint64_t)k, K, act);
            fjsp_oracle_step(e, act, obs, masks, rew, flags, NULL);
            acc[0] += 1;
            acc[2] += (uint64_t)(e->completed_orders - before_o);
            acc[3] += (uint64_t)(e->total_products_packaged - before_p);
            for (int a = 0; a < A; a++) rsum += (int64_t)llround(rew[a] * 10.0 * (double)A);
            if (flags[0] | flags[1] | flags[2]) {
                FjspOrderRec orders[FJSP_MAX_ORDERS];
                acc[1] += 1;
                acc[4] += flags[2] ? 1 : 0;
                episode += 1;
                fjsp_oracle_philox_orders(j->seed, genv, episode, j->num_orders, orders);
                fjsp_oracle_reset(e, orders, j->num_orders);
            }
        }
        j->episodes[i] = episode;
    }
    for (int k = 0; k < 8; k++) j->acc[k] = acc[k];
    j->rsum = rsum;
    return NULL;
}

void fjsp_oracle_rollout_random(OracleEnv** envs, uint32_t* episodes, int64_t n_envs, int64_t first_env, int steps,
                                uint64_t seed, uint64_t t0, int num_orders, int nthreads, uint64_t* stats) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > n_envs) nthreads = n_envs > 0 ? (int)n_envs : 1;
    RolloutJob jobs[256];
    pthread_t tids[256];
    for (int t = 0; t < nthreads; t++) {
        RolloutJob* j = &jobs[t];
        memset(j, 0, sizeof(*j));
        j->envs = envs, j->episodes = episodes, j->first_env = first_env;
        j->lo = n_envs * t / nthreads, j->hi = n_envs * (t + 1) / nthreads;
        j->steps = steps, j->num_orders = num_orders, j->seed = seed, j->t0 = t0;
        if (t > 0) pthread_create(&tids[t], NULL, rollout_worker, j);
    }
    rollout_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(tids[t], NULL);
    for (int t = 0; t < nthreads; t++) {
        for (int k = 0; k < 8; k++) stats[k] += jobs[t].acc[k];
        stats[5] += (uint64_t)jobs[t].rsum;
    }
}
