/*
 * fjsp_b200.h — C ABI of the B200-native batched FJSP environment (libfjsp_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of FARIDKH/Multi-agent-RL-for-FJSP: the
 * environment step.  The reference has no FFI; its boundary is the Python class
 * FJSPParallelEnv (FJSPParallelEnvWrapper.py:9-136) over FJSPSimulation (FJSPSimulation.py:27-430).
 * Each entry point below names the reference interface it replaces.  The Python host side
 * (multi_agent_rl_for_fjsp_b200/env.py, FJSPParallelEnvWrapper.py) binds these with ctypes and
 * passes torch tensors' data_ptr(); no torch type appears here.
 *
 * Conventions: every function returns 0 on success, non-zero on error (fjsp_last_error() gives the
 * thread-local message).  No exceptions cross the ABI.  Device pointers are caller-owned.  Calls
 * that take a `stream` enqueue work on that cudaStream_t (passed as void*) and do not synchronise
 * the host unless stated.  A handle belongs to one device and is not thread-safe.
 */
#ifndef FJSP_B200_H
#define FJSP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FJSP_ABI_VERSION 4

/* ---- fixed shape of the problem (reference: 8 agents, FJSPSimulation.py:62-82) ---- */
#define FJSP_NUM_AGENTS 8          /* pickup_station, agv, small_machine, big_machine, packaging_blue_1, _blue_2, _red, _green */
#define FJSP_OBS_DIM 38            /* a2c._flatten_obs order, SURVEY.md §8a-R9 (a2c.py:137-151) */
#define FJSP_MASK_DIM 32           /* 29 used: 3 + 8 + 6*3, padded to 32 */
#define FJSP_MASK_USED 29
#define FJSP_FLAG_DIM 4            /* terminated, truncated, fault, was_reset */
#define FJSP_INFO_DIM 4            /* current_step, orders_completed, total_products_packaged, reserved */
#define FJSP_MAX_ORDERS 32         /* reference reset default is 30 (FJSPSimulation.py:316) */
#define FJSP_MAX_ORDER_PRODUCTS 9  /* np.random.randint(1,10) (FJSPSimulation.py:107) */
#define FJSP_NUM_LOCATIONS 5       /* LocationType order: PICKUP, BIG_MACHINE, SMALL_MACHINE, STORAGE, PACKAGING */
#define FJSP_TRAY_CAPACITY 5       /* constants.py:22; the mask hard-codes the global (PickupStationAgent.py:125) */
#define FJSP_POOL_SLOTS 64         /* trays in transit per env (bound: 1 + max_episode_steps/4 = 51) */
#define FJSP_STATE_WORDS 128       /* packed state: 128 x u32 = 512 B per env */
#define FJSP_TILE_ENVS 64          /* envs per HBM tile (array-of-tiles, each tile word-major [words][64]) */

/* ---- scaled shop (BASELINE configs[4], builder-defined extension): K = num_cells "cells", each with its
 * own AGV, small machine, big machine, four packaging stations, storage and tray pool, sharing ONE pickup station /
 * order stream and its single dock.  K = 1 is the reference shop, bit for bit.  Agents (and the action / reward /
 * result columns): pickup_station, then per cell: agv, small_machine, big_machine, packaging_blue_1, _blue_2, _red,
 * _green.  Row widths below; K = 1 gives the reference's 8 / 38 / 32. */
#define FJSP_MAX_CELLS 4
#define FJSP_AGENTS_K(k) (1 + 7 * (k))
#define FJSP_ACT_DIM_K(k) ((FJSP_AGENTS_K(k) + 7) / 8 * 8)      /* actions u8, rewards f32, results u8 per env */
#define FJSP_OBS_DIM_K(k) (7 + 31 * (k))
#define FJSP_MASK_DIM_K(k) ((3 + 26 * (k) + 31) / 32 * 32)
#define FJSP_STATE_WORDS_K(k) (64 + 64 * (k) + 20 * ((k) - 1))  /* 128 words for K = 1, 380 for K = 4 */

/* ---- shared floor (BASELINE configs[4] "multi-AGV with collision"; FjspConfig.shared_agvs = A in 2..4; a builder-defined
 * extension — the reference has ONE AGV, /root/reference/FJSPSimulation.py:62-82): A AGVs serve ONE set of stations (pickup
 * station, small / big machine, storage, four packaging stations) and are not sealed into cells.
 *   agents   : pickup station, agv_0 .. agv_{A-1}, small machine, big machine, four packaging stations (7 + A), acting in
 *              that order (A = 1 is the reference's order, FJSPSimulation.py:76-82,172-174);
 *   occupancy: each of the five station positions holds at most ONE AGV.  A standing AGV occupies its station; an AGV under
 *              way occupies its TARGET from the moment the move is granted (its origin is free at once).  A move to a position
 *              occupied by another AGV is an invalid action (-5, AGVAgent.py:210-212's penalty) and masked out; two AGVs that
 *              ask for the same free position in one step are served in agent order — the earlier one gets it, the later
 *              one's action is invalid.  A move to the AGV's own position succeeds without moving, as in the reference;
 *   start    : agv_0 at PICKUP (AGVAgent.py:41), agv_1 at STORAGE, agv_2 at SMALL, agv_3 at BIG;
 *   rewards  : r_i = g / (7 + A) + local_i with the reference's g and local terms (RewardModel.py:34-110);
 *   layout   : observation = pickup station (7) | 13 per AGV | small, big machine (3 + 3) | stations (12); masks = 3 | 8 per
 *              AGV | 3 + 3 | 12, padded to 16; actions / rewards / results padded to 8; packed state = the 128 words of the
 *              reference shop + one word per further AGV (132 words = 528 B per env); Philox actions: the reference stream's
 *              eight values for (pickup station, agv_0, machines, stations), further AGVs from counter word 3 = 17.
 * Served by fjsp_reset / fjsp_step / fjsp_random_actions / fjsp_export_state_cell (cell index = AGV index) / fjsp_export_packed /
 * fjsp_export_orders / fjsp_state_save / fjsp_state_load; the wire-row, host-buffer and K-steps-per-launch entry points refuse such a handle. */
#define FJSP_MAX_SHARED_AGVS 4
#define FJSP_SHARED_AGENTS(a) (7 + (a))
#define FJSP_SHARED_ACT_DIM(a) ((FJSP_SHARED_AGENTS(a) + 7) / 8 * 8)
#define FJSP_SHARED_OBS_DIM(a) (25 + 13 * (a))
#define FJSP_SHARED_MASK_DIM(a) ((21 + 8 * (a) + 15) / 16 * 16)
#define FJSP_SHARED_STATE_WORDS 132

/* ---- long order streams (BASELINE configs[4]; FjspConfig.long_streams = 1): a second packed layout for episodes with more
 * than FJSP_MAX_ORDERS orders and / or more than 240 steps, and for order ARRIVALS during the episode.  The reference
 * accepts any num_orders (FJSPSimulation.py:315-318) and max_episode_steps (:223); this layout takes up to
 * FJSP_LONG_MAX_ORDERS orders and FJSP_LONG_MAX_STEPS steps.  Queued orders cost nothing (their attributes come from the
 * Philox stream / the explicit table when the pickup station pops them), trays waiting at the pickup station live in a
 * per-env ready FIFO of FJSP_LONG_READY_FIFO words outside the packed state (the reference hands out at most
 * min(num_trays, 1000) trays per episode, FJSPSimulation.py:96), and an order occupies one of FJSP_LONG_ORDER_SLOTS order
 * slots from the moment an AGV takes its first tray until one step after it completes.  More orders in process at once
 * than slots -> FJSP_FAULT_ORDER_SLOTS.  With long_streams = 0 and at most 32 orders nothing changes: the compact layout
 * is used, bit for bit.
 * Arrivals (builder-defined extension; the reference creates all orders at reset): with arrival_prob_q16 > 0 one more
 * order arrives per step with probability arrival_prob_q16 / 65536 — Philox4x32-10(key = seed, counter = (global env,
 * episode, step, 5)), low 16 bits < arrival_prob_q16 — until arrival_max_orders exist; order i always has the attributes of
 * index i of the order stream.  An episode then terminates only when all arrival_max_orders orders are complete. */
#define FJSP_LONG_ORDER_SLOTS 64
#define FJSP_LONG_READY_FIFO 1024
#define FJSP_LONG_MAX_ORDERS 4095
#define FJSP_LONG_MAX_STEPS 65000
#define FJSP_STATE_WORDS_LONG_K(k) (228 + 64 * (k) + 24 * ((k) - 1))  /* 292 words = 1168 B for K = 1 */

/* ---- wire rows: the compact form in which a step's results cross PCIe on the host-buffer path (fjsp_step_host) and
 * which fjsp_step_wire / fjsp_wire_decode expose.  Everything a step returns is a small integer, so one env's results
 * are bit-fields in FJSP_WIRE_WORDS_K(K) u32 — 32 bytes for K = 1, against 220 bytes of float32 / int8 tensors:
 *   word 0  pickup station: current_tray_color2 | current_tray_count3<<2 | current_tray_type2<<5 | next_product_color2<<7 |
 *           next_product_type2<<9 | order_size4<<11 | products_remaining4<<15 ; flag bits<<19 (terminated, truncated,
 *           fault x3, was_reset) ; its action-mask bits 1,2 <<25 ; its reward code2<<27
 *   word 1  products packaged by this step 10 | orders completed by this step 9<<10 | pickup_ready_trays 13<<19
 *   then five words per cell c (2 + 5c ..):
 *   +0  agv: station index3 (LocationType order; decoded through FjspConfig.pos into BOTH position fields) | carrying1<<3 |
 *       tray_needs_processing1<<4 | tray_product_count3<<5 | tray_type2<<8 | big_machine_ready6<<10 |
 *       small_machine_ready6<<16 | agv mask bits 1..7 <<22 | agv reward code3<<29
 *   +1  storage_tray_count8 | small machine (is_busy1, progress1, queue_length6)<<8 | big machine<<16 |
 *       small mask bits 1,2 <<24 | big<<26 | small reward code2<<28 | big<<30
 *   +2  packaging_blue_1 | packaging_green: progress index8<<21, is_busy1<<29, mask bits 1,2 <<30
 *   +3  packaging_blue_2 | packaging_green: queue_length8<<21, reward code2<<29
 *   +4  packaging_red           with station = is_busy1 | progress index L8<<1 (decoded to float32(100/L), 0 for 0) |
 *                               queue_length8<<9 (the int8 it is) | mask bits 1,2 <<17 | reward code2<<19
 * Mask bit 0 of every agent is the constant 1.  A reward code selects one of the few local rewards an agent kind can get
 * (tenths: pickup {0,-10,+10,+60}, agv {0,-1,+20,+120,-50}, machine {0,-20,+10,+50}, packaging {0,-10,+20,+200};
 * RewardModel.py:46-97): reward_i = (g + A * local_i) / (10 * A) with g = 10 * (100 * orders + 10 * products) - step_size and
 * A = 1 + 7K — the exact integer numerator of every reward.  fjsp_wire_decode turns rows back into exactly the
 * tensors fjsp_step writes (bit for bit). */
#define FJSP_WIRE_WORDS_K(k) (((2 + 5 * (k)) + 1) / 2 * 2)

/* fault codes (flags[2]); the reference has no equivalent — see DESIGN.md "faults" */
#define FJSP_FAULT_NONE 0
#define FJSP_FAULT_PKG_RESTART_WITH_WAITERS 1 /* reference raises ValueError out of env.run (SURVEY R-PKG-cap-b) */
#define FJSP_FAULT_POOL_EXHAUSTED 2           /* >64 trays in transit: impossible when max_episode_steps <= 240 */
#define FJSP_FAULT_BAD_ORDER 3                /* an explicit order record outside n 1..9 / type 1..3 / colour 1..3 (set at reset) */
#define FJSP_FAULT_PAST_END 4                 /* stepped after the truncation step without a reset: the env is inert (the
                                                 reference would simulate on; the packed counters are not sized for it) */
#define FJSP_FAULT_ORDER_SLOTS 5              /* long layout: more than FJSP_LONG_ORDER_SLOTS orders in process at once */

/* Mirrors constants.py:5-32 (LOCATION_POSITIONS, PROCESSING_TIMES, CONFIG). */
typedef struct FjspConfig {
    int32_t struct_size;                   /* sizeof(FjspConfig), for ABI checking */
    int32_t pos[FJSP_NUM_LOCATIONS][2];    /* (row, col) per LocationType, constants.py:5-11 */
    int32_t grid_rows, grid_cols;          /* constants.py:26-27 (observation-space bounds only) */
    int32_t proc_small, proc_big, proc_pack; /* constants.py:14-18, sim time units; multiples of step_size */
    int32_t step_size;                     /* constants.py:29 */
    int32_t agv_speed;                     /* constants.py:28 */
    int32_t max_episode_steps;             /* constants.py:31; <= 240 */
    int32_t storage_capacity;              /* FJSPSimulation.py:87 default 100 */
    int32_t pack_capacity;                 /* PackagingAgent.py:46 simpy.Resource(capacity=20); <= 31 */
    int32_t tray_capacity;                 /* must be 5 */
    int32_t num_trays;                     /* constants.py:21; min(num_trays,1000) reach the pickup station (FJSPSimulation.py:96) */
    int32_t num_cells;                     /* 1 = the reference shop; 2..4 = scaled shop (see above) */
    int32_t long_streams;                  /* 0 = compact layout (<= 32 orders, <= 240 steps); 1 = long order streams (see above) */
    int32_t arrival_prob_q16;              /* long_streams only: P(one order arrives in a step) * 65536; 0 = no arrivals */
    int32_t arrival_max_orders;            /* long_streams only: orders stop arriving once this many exist */
    int32_t shared_agvs;                   /* 0 / 1 = one AGV (the reference); 2..4 = SHARED FLOOR: that many AGVs on one set of
                                              stations with station occupancy (see "shared floor" above); needs num_cells = 1,
                                              long_streams = 0 */
} FjspConfig;

/* One order as the reference generates it (FJSPSimulation.py:101-131): all products of an order
 * share type and colour.  Packed u32: n | type<<8 | colour<<16. */
/* Valid ranges: n 1..FJSP_MAX_ORDER_PRODUCTS, type 1..3 (ProductType), colour 1..3 (PackagingColor); a record
 * outside them is clamped and the env reports FJSP_FAULT_BAD_ORDER with its first step. */
typedef uint32_t FjspOrderRec;
#define FJSP_ORDER_REC(n, type, colour) ((uint32_t)(n) | ((uint32_t)(type) << 8) | ((uint32_t)(colour) << 16))

/* ---- canonical integer state S: the record both sides export for bit-exact diff ----
 * tray entry  = tray_id | order<<16 | first_product_idx<<22 | count<<26      (-1 = none)
 * product id  = order*100 + idx  (FJSPSimulation.py:115)                     (-1 = none)            */
#define FJSP_CANON_MAXQ 64
#define FJSP_CANON_PS_READY 256
#define FJSP_CANON_MAXPQ 256
#define FJSP_TRAY_ENTRY(id, order, first, count) ((int32_t)((id) | ((order) << 16) | ((first) << 22) | ((count) << 26)))

typedef struct FjspCanonMachine {
    int32_t is_busy, current_tray, progress_done;
    int32_t queue_n, queue[FJSP_CANON_MAXQ];
    int32_t ready_n, ready[FJSP_CANON_MAXQ];
} FjspCanonMachine;

typedef struct FjspCanonPack {
    int32_t is_busy, current_product, progress_L, products_completed, users;
    int32_t queue_n, queue[FJSP_CANON_MAXPQ];
} FjspCanonPack;

typedef struct FjspCanonState {
    int32_t current_step, num_orders, fault;
    int32_t agv_row, agv_col, agv_carry, agv_is_moving;
    int32_t ps_order_queue_len, ps_current_order, ps_product_idx, ps_current_tray, ps_trays_at_station;
    int32_t ps_ready_n, ps_ready[FJSP_CANON_PS_READY];
    FjspCanonMachine machine[2];           /* 0 = small_machine, 1 = big_machine */
    int32_t storage_n, storage[FJSP_CANON_MAXQ];
    FjspCanonPack pack[4];                 /* blue_1, blue_2, red, green */
    int32_t processed_mask[FJSP_MAX_ORDERS];  /* bit i = products[i].is_processed */
    int32_t packaged_mask[FJSP_MAX_ORDERS];   /* bit i = products[i].is_packaged */
    int32_t order_complete[FJSP_MAX_ORDERS];
    int32_t order_completion_step[FJSP_MAX_ORDERS]; /* completion_time / step_size - 1, or -1 */
    int32_t total_products_packaged, completed_orders;
} FjspCanonState;

typedef struct FjspHandle FjspHandle;

/* Per-rollout counters written by fjsp_rollout_random (device memory, 8 x u64). */
#define FJSP_STATS_WORDS 8 /* env_steps, episodes, orders_completed, products_packaged, faults, reward_sum_x40, 0, 0 */

const char* fjsp_last_error(void);
int fjsp_abi_version(void);

/* constants.py:5-32 defaults. */
int fjsp_default_config(FjspConfig* cfg);

/* FJSPParallelEnv.__init__ (FJSPParallelEnvWrapper.py:27-33) for `num_envs` independent shops on
 * CUDA device `device`.  `first_env` is the global index of env 0 of this handle (multi-GPU
 * sharding: Philox counters use first_env + local index, so results do not depend on the shard map). */
int fjsp_create(const FjspConfig* cfg, int64_t num_envs, int64_t first_env, int device, FjspHandle** out);
int fjsp_destroy(FjspHandle* h);
int64_t fjsp_num_envs(const FjspHandle* h);
int fjsp_num_cells(const FjspHandle* h);        /* K of the handle's config (1 = the reference shop) */
size_t fjsp_state_bytes(const FjspHandle* h);   /* packed HBM bytes per env (512 for K = 1, 4 * FJSP_STATE_WORDS_K(K)) */
void* fjsp_state_ptr(FjspHandle* h);            /* device pointer of the packed state (tiles) */

/* FJSPParallelEnv.reset (FJSPParallelEnvWrapper.py:43-54) -> FJSPSimulation.reset (:286-323).
 *   env_mask  : device u8[N] (non-zero = reset that env) or NULL = all
 *   orders    : device FjspOrderRec[N][FJSP_MAX_ORDERS] explicit order tables (long layout: [N][W] with W = num_orders, or
 *               arrival_max_orders when arrivals are on — one record per order that can ever exist; copied into the
 *               handle), or NULL = draw
 *               n~U{1..9}, type~U{1..3}, colour~U{1..3} from Philox4x32-10(key=seed,
 *               counter=(global env, episode, order, 0)) — replayable on the host
 *   num_orders: 0..32 (reference default 30); long layout: 0..FJSP_LONG_MAX_ORDERS
 *   obs/masks : device float[N][38] / int8[N][32] initial observations (may be NULL)            */
int fjsp_reset(FjspHandle* h, const uint8_t* env_mask, uint64_t seed, const FjspOrderRec* orders,
               int num_orders, float* obs, int8_t* masks, void* stream);

/* FJSPParallelEnv.step (FJSPParallelEnvWrapper.py:56-69) -> FJSPSimulation.step (:144-242) for all
 * N envs in lockstep; ONE kernel launch.
 *   actions : device u8[N][8], agent order = FJSPSimulation.py:76-82
 *   obs     : device float[N][38]    masks: device int8[N][32]    rewards: device float[N][8]
 *   flags   : device u8[N][4] = terminated, truncated, fault, was_reset
 *   results : device u8[N][8] per-agent action_result bit-fields, or NULL   (see FJSP_RES_*)
 *   infos   : device int32[N][4] = current_step, orders_completed, total_products_packaged, 0, or NULL
 *   autoreset: non-zero = envs that end (terminated|truncated|fault) are re-initialised inside the
 *              step (episode counter + 1, fresh Philox orders) and return the post-reset observation */
int fjsp_step(FjspHandle* h, const uint8_t* actions, float* obs, int8_t* masks, float* rewards,
              uint8_t* flags, uint8_t* results, int32_t* infos, int autoreset, void* stream);

/* Same call with HOST buffers (pinned or pageable): H2D actions, step, D2H outputs, stream
 * synchronised on return.  This is the end-to-end path bench.py reports as `e2e`.  The results cross PCIe as wire
 * rows (see above) in pipelined chunks and are decoded into the caller's buffers by the library's host threads while
 * later chunks are still in flight; what lands in obs/masks/rewards/flags is bit for bit what fjsp_step writes. */
int fjsp_step_host(FjspHandle* h, const uint8_t* actions, float* obs, int8_t* masks, float* rewards,
                   uint8_t* flags, int autoreset, void* stream);

/* The same pipelined host-buffer step WITHOUT the decode: wire = host u32[N][FJSP_WIRE_WORDS_K(K)] (pinned for full
 * speed) receives the wire rows; the caller decodes them (fjsp_wire_decode) where and when it needs the tensors, or
 * consumes the integers as they are.  Synchronised on return. */
int fjsp_step_host_wire(FjspHandle* h, const uint8_t* actions, uint32_t* wire, int autoreset, void* stream);

/* Measurement aid (bench.py e2e.host_stream_write_gbs): one pass of streaming stores over a HOST buffer from `threads`
 * threads — the pattern fjsp_step_host's decode delivers the tensors with, i.e. the box's ceiling for it. */
int fjsp_host_stream_write_probe(void* host_buf, size_t bytes, int threads, double* seconds);

/* Host threads fjsp_step_host may use for the decode (including the caller's); 0 = every CPU the process may run on
 * (default).  Several handles / ranks on one host should share the cores out.  Call before the first fjsp_step_host. */
int fjsp_set_decode_threads(FjspHandle* h, int threads);

/* fjsp_step with the results written as wire rows: wire = device u32[N][FJSP_WIRE_WORDS_K(K)] (16-byte aligned).
 * results / infos as in fjsp_step (may be NULL).  ONE kernel launch. */
int fjsp_step_wire(FjspHandle* h, const uint8_t* actions, uint32_t* wire, uint8_t* results, int32_t* infos, int autoreset,
                   void* stream);
/* HOST function (no GPU needed): decode n wire rows of a `cfg` shop into the tensors fjsp_step writes —
 * obs float[n][OBS], masks int8[n][MASK], rewards float[n][ACT], flags u8[n][4]; any output may be NULL.
 * `threads` <= 1 decodes on the calling thread, otherwise on that many short-lived threads. */
int fjsp_wire_decode(const FjspConfig* cfg, const uint32_t* wire, int64_t n, float* obs, int8_t* masks, float* rewards,
                     uint8_t* flags, int threads);
size_t fjsp_wire_row_bytes(int num_cells);

/* a_i ~ U{0..n_i-1}, n = (3,8,3,3,3,3,3,3), Philox4x32-10(key=seed, counter=(global env, t, 0, 1)). */
int fjsp_random_actions(FjspHandle* h, uint64_t seed, uint64_t t, uint8_t* actions, void* stream);

/* `steps` lockstep steps with in-kernel random actions (same stream as fjsp_random_actions with
 * t = t0 .. t0+steps-1) and autoreset, state kept on-chip between steps; adds to stats[8] (device u64). */
int fjsp_rollout_random(FjspHandle* h, int steps, uint64_t seed, uint64_t t0, uint64_t* stats, void* stream);

/* Canonical record S of one env, decoded on the host (synchronises). Diagnostic / parity path.
 * For a scaled shop the record describes ONE cell (AGV, machines, storage, packaging) plus the shared pickup station
 * and order table; fjsp_export_state is cell 0. */
int fjsp_export_state(FjspHandle* h, int64_t env, FjspCanonState* out);
int fjsp_export_state_cell(FjspHandle* h, int64_t env, int cell, FjspCanonState* out);
/* Raw packed words of one env (FJSP_STATE_WORDS_K(K) x u32; 128 for K = 1) copied to the host (synchronises). */
int fjsp_export_packed(FjspHandle* h, int64_t env, uint32_t* out_words);
/* Per-order record of orders [first, first + count) of one env, decoded on the host (synchronises):
 * out[i] = { packaged_mask, processed_mask, is_complete, completion_step (or -1) }.  Works for both layouts; in the long
 * layout orders whose slot has been given back (one step after completion) are reported as complete with every bit set
 * and completion_step -1, orders no AGV has touched yet as zeros.  In the long layout FjspCanonState's
 * per-order arrays describe the 32 most recently popped orders (order_base = max(0, next_order - 32) is returned
 * here through *order_base when it is not NULL) and tray entries are FJSP_TRAY_ENTRY_LONG. */
int fjsp_export_orders(FjspHandle* h, int64_t env, int first, int count, int32_t* out4, int32_t* order_base);
#define FJSP_TRAY_ENTRY_LONG(id, order, first, count) ((int32_t)((id) | ((order) << 12) | ((first) << 24) | ((count) << 28)))

/* Snapshot / restore of the whole packed state (num_tiles * 32 KB, see fjsp_state_total_bytes) to / from a
 * caller-owned DEVICE buffer: env checkpointing, and bit-for-bit comparison of two handles. */
size_t fjsp_state_total_bytes(const FjspHandle* h);
int fjsp_state_save(FjspHandle* h, void* dst_device, size_t bytes, void* stream);
int fjsp_state_load(FjspHandle* h, const void* src_device, size_t bytes, void* stream);

/* Number of kernels this library has launched since the handle was created. */
int64_t fjsp_launch_count(const FjspHandle* h);

/* ---- fused device ops for a batched A2C trainer, the caller of the path.  Stateless. ----
 * fjsp_a2c_sample: per row, for each of the 8 agents: softmax(logits segment) * mask, renormalise (uniform over valid
 *   actions when the masked mass is 0), Categorical sample, log-prob of the sample (a2c.py:214-247).
 *   logits float[rows][32] and masks int8[rows][32] use the env's mask layout (3 | 8 | 6x3 | pad); actions u8[rows][8];
 *   logp float[rows][8] or NULL.  Uniforms: Philox4x32-10(key=seed, counter=(first_row+row, t_lo, t_hi, 2|3)) with
 *   t = counter[0] + t_off (counter: device u64 or NULL = 0), so a captured CUDA graph draws fresh numbers each replay.
 * fjsp_a2c_gae: returns and GAE per (env, agent) over T steps with the shared critic value
 *   (transition_memory.py:83-105): rewards float[T][N][8], values float[T+1][N] (values[T] = bootstrap),
 *   flags u8[T][N][4] as written by fjsp_step (terminated|truncated|fault ends the episode: bootstrap 0). */
int fjsp_a2c_sample(const float* logits, const int8_t* masks, uint8_t* actions, float* logp, int64_t rows, int64_t first_row,
                    uint64_t seed, const uint64_t* counter, uint64_t t_off, void* stream);
int fjsp_a2c_counter_add(uint64_t* counter, uint64_t inc, void* stream);
int fjsp_a2c_gae(const float* rewards, const float* values, const uint8_t* flags, float* returns, float* advantages, int T, int64_t N,
                 float gamma, float lamb, void* stream);

/* Cell views of a scaled shop for a trainer that shares the reference's 8 networks between the cells: view row
 * env * K + cell holds the reference's own layout — obs float[38] = pickup station's 7 fields + the cell's 31, masks
 * int8[32] = 3 + the cell's 26 (+3 pad), rewards float[8] = pickup station's + the cell's 7, flags u8[4] (copied).  The
 * pickup station acts through the row of cell 0; in the other rows its mask allows action 0 only.
 * fjsp_cells_unpack_views: env tensors (as fjsp_step writes them) -> view tensors [num_envs * K][...].
 * fjsp_cells_pack_actions: view actions u8[num_envs * K][8] -> env actions u8[num_envs][FJSP_ACT_DIM_K(K)].  Stateless. */
int fjsp_cells_pack_actions(const uint8_t* view_actions, uint8_t* actions, int64_t num_envs, int num_cells, void* stream);
int fjsp_cells_unpack_views(const float* obs, const int8_t* masks, const float* rewards, const uint8_t* flags, float* v_obs,
                            int8_t* v_masks, float* v_rewards, uint8_t* v_flags, int64_t num_envs, int num_cells, void* stream);

/* ---- tensor-core GEMM for the trainer's actor / critic layers (networks.py:22-61; a2c.py:168-252,647-731) ----
 * fjsp_a2c_gemm: C[M x N] (+)= A[M x K] * B[N x K]^T for a device-resident TABLE of problems in one launch (the 8 actors
 * and the critic are one grouped launch per layer), computed by tcgen05.mma.kind::tf32 with accumulators in TMEM.
 * passes = 3: every fp32 operand is split into two TF32 terms and three products are accumulated ("3xTF32": fp32-level
 * accuracy, ~2^-21 relative per product); passes = 1: plain TF32.
 * Operand orientation (the same for all problems of a call): FJSP_OP_KC  X(r,k) = X[r*ld + k] with 16-byte loads
 * (ld % 4 == 0, 16-byte aligned pointer, K % 4 == 0), FJSP_OP_KCS the same with scalar loads (any alignment),
 * FJSP_OP_MC  X(r,k) = X[k*ld + r].  Supported (a_op, b_op): (KC,MC) forward y = x W, (KCS,MC) forward from an
 * unaligned observation slice, (KC,KC) dx = dy W^T, (KCS,KCS) the same from unaligned rows, (MC,MC) dW = x^T dy.
 * Epilogue per problem: + bias[n]; ReLU (FJSP_GEMM_RELU); * (mask(m,n) > 0) (mask indexed like C);
 * colsum[n] += column sums of the stored values (bias gradient); FJSP_GEMM_ATOMIC: atomicAdd into C (split-K);
 * rowdot_*: a fused 1-column head on the stored values (the critic's 128 -> 1 layer).
 * max_ctas = max over the problems of ceil(M / 128) * splitk.  N <= 256.
 * FJSP_OP_PK (b_op only; pairs (KC,PK), (KCS,PK)): B is a PACKED IMAGE of a weight matrix made by fjsp_a2c_gemm_pack —
 * per 16-wide K chunk the bytes the tensor core reads from shared memory (hi terms, then lo terms of the 3xTF32 split,
 * four K-quad planes of npad = roundup(N, 16) rows x 16 bytes each): a stage's B tile is a straight copy, no conversion
 * work per use.  Image size: ceil(K / 16) * 32 * npad floats, 16-byte aligned.  Results are bit-identical to the unpacked
 * orientations (same split, same accumulation order).  Measured: the same speed as the on-the-fly split —
 * the kernel's two-stage main loop is latency-bound, not conversion-bound — so the trainer does not use it yet; it is
 * the B half of a bulk-copy-fed main loop. */
#define FJSP_OP_KC 0
#define FJSP_OP_KCS 1
#define FJSP_OP_MC 2
#define FJSP_OP_PK 3
#define FJSP_PACK_IMAGE_FLOATS(n, k) ((((k) + 15) / 16) * 32 * (((n) + 15) / 16 * 16))
typedef struct FjspPackJob {   /* device array; one job = one weight matrix in one orientation */
    const float* src;
    float* dst;                /* FJSP_PACK_IMAGE_FLOATS(N, K) floats */
    int32_t op;                /* FJSP_OP_KC / KCS: B(n,k) = src[n*ld + k]; FJSP_OP_MC: B(n,k) = src[k*ld + n] */
    int32_t ld, N, K;
} FjspPackJob;
#define FJSP_GEMM_RELU 1
#define FJSP_GEMM_ATOMIC 2
typedef struct FjspGemmProb {
    const float* A;
    const float* B;
    float* C;
    const float* bias;
    const float* mask;
    float* colsum;
    int32_t M, N, K;
    int32_t lda, ldb;
    int32_t csm, csn;      /* C(m, n) = C[m * csm + n * csn] */
    int32_t flags;
    int32_t splitk;
    int32_t head_n;          /* 0: rowdot_* is the one-column head; 1..8: a fused head of that many columns, any N <= 256:      */
    int32_t head_ld;         /*   rowdot_out[m * head_ld + j] = sum_n C(m, n) * rowdot_w[n * head_n + j] + rowdot_bias[j]   */
    int32_t reserved;        /*   (an actor's 256 -> 3..8 logits layer in the epilogue of its second layer, fp32 FMAs)      */
    const float* rowdot_w;   /* optional fused head, N <= 128: rowdot_out[m] = sum_n C(m, n) * rowdot_w[n] + rowdot_bias[0] */
    float* rowdot_out;
    const float* rowdot_bias;
    int64_t reserved2;
} FjspGemmProb;
/* fjsp_a2c_loss_grad: the gradients of one A2C update's losses with respect to the actors' pre-softmax outputs and the
 * critic value, analytically (a2c.py:647-731: -mean(adv_n * log q[a]) - entropy_coef * mean(H(softmax)) per actor with
 * q = masked renormalised softmax and Categorical's clamp; MSE of the shared value against every agent's returns).
 * logits/dlogits float[rows][32] and masks int8[rows][32] in the env's mask layout, actions u8[rows][8], adv/returns
 * float[rows][8], values/dvalue float[rows]; adv_mean / adv_rstd float[8] (device): adv_n = (adv - mean) * rstd.
 * sums float[64] (device, += ): [0..31] column sums of dlogits, [32..39] sum(-adv_n log q) per agent, [40..47] sum of
 * entropies per agent, [48] sum of squared value errors, [49] sum of dvalue. */
int fjsp_a2c_loss_grad(const float* logits, const int8_t* masks, const uint8_t* actions, const float* adv, const float* returns,
                       const float* values, const float* adv_mean, const float* adv_rstd, float entropy_coef, int64_t rows, float* dlogits,
                       float* dvalue, float* sums, void* stream);
int fjsp_a2c_gemm_pack(const FjspPackJob* jobs_device, int njobs, void* stream);
/* clip_grad_norm_(max_norm) per network, then Adam (a2c.py:668,686-690), over a table of parameter segments (one per
 * (tensor, network): the six small actors are slices of stacked tensors): three launches instead of torch's ~20.  The
 * arithmetic of torch.optim.Adam (no amsgrad / weight decay) on the optimizer's own state tensors: m, v, and a float step
 * counter per parameter tensor (the segment with bump != 0 increments it; every segment of a tensor reads it).
 * beta1 / beta2 / eps are doubles as in torch (1 - beta is formed in double, then the per-element arithmetic is float).
 * norms_sq: float[16] device scratch, zero before the first call (the call leaves it zero).  max_elems = largest n. */
typedef struct FjspOptSeg {
    float* param;
    float* grad;
    float* m;
    float* v;
    float* step;
    int32_t n, net;
    float lr;
    int32_t bump;
    int64_t reserved;
} FjspOptSeg;
int fjsp_a2c_clip_adam(const FjspOptSeg* segs_device, int nseg, int max_elems, float* norms_sq, float max_norm, double beta1,
                       double beta2, double eps, void* stream);
/* Backward through a narrow head (an actor's 256 -> 3..8 logits layer, the critic's 128 -> 1 value head; a2c.py:647-731) in one
 * pass over the head's post-ReLU input H: dH[m, n] = (sum_j dl[m, j] W[n, j]) * (H[m, n] > 0); gb[n] += sum_m dH[m, n] (the
 * bias gradient of the layer below); gW[n, j] += sum_m H[m, n] dl[m, j].  n % 4 == 0, n <= 256, na <= 8, H / dH rows of n
 * floats, 16-byte aligned.  gW and gb are accumulated into (zero them first).  max_rows = the largest `rows` of the jobs. */
typedef struct FjspHeadBwdJob {
    const float* dl;       /* [rows][ld_dl], columns 0..na-1 */
    const float* W;        /* [n][na] row-major */
    const float* H;        /* [rows][n] */
    float* dH;             /* [rows][n] */
    float* gW;             /* [n][na] */
    float* gb;             /* [n] or NULL */
    int32_t rows, n, na, ld_dl;
} FjspHeadBwdJob;
int fjsp_a2c_head_backward(const FjspHeadBwdJob* jobs_device, int njobs, int max_rows, void* stream);
/* y = relu?(x W + b) with K <= 40, N <= 256 as fp32 FMAs, one job per network: the actors' first layers in a rollout step
 * (networks.py:22-38: Linear(3..13 -> 256) + ReLU).  The caller guarantees k <= 40 and n <= 256 (checked on the host side of
 * the table).  max_rows / max_k = the largest `rows` / `k` of the jobs. */
typedef struct FjspLayer1Job {
    const float* X;        /* [rows][ldx], columns 0..k-1 */
    const float* W;        /* [k][n] row-major */
    const float* bias;     /* [n] or NULL */
    float* Y;              /* [rows][ldy] */
    int32_t rows, k, n, ldx, ldy, relu;
    int32_t reserved[2];
} FjspLayer1Job;
int fjsp_a2c_layer1(const FjspLayer1Job* jobs_device, int njobs, int max_rows, int max_k, void* stream);
/* Weight gradients with a narrow side, G[i*gsi + j*gsj] += sum_b X[b*ldx + i] * Y[b*ldy + j] for i < nx <= 256, j < ny <= 40:
 * the actors' heads, the first layers (transposed) and the critic's value head (a2c.py:647-731 backward of networks.py:22-61).
 * fp32 FMAs, X streamed once; as tensor-core GEMMs these cost as much as a 256 x 256 product each.  max_rows = max B of the
 * jobs, max_ny = max ny.  G must be zeroed (or hold the value to add to) by the caller. */
typedef struct FjspWgradJob {
    const float* X;
    const float* Y;
    float* G;
    int32_t B, nx, ny, ldx, ldy, gsi, gsj;
    int32_t reserved[3];
} FjspWgradJob;
int fjsp_a2c_wgrad_small(const FjspWgradJob* jobs_device, int njobs, int max_rows, int max_ny, void* stream);
int fjsp_a2c_gemm(const FjspGemmProb* probs_device, int nprob, int max_ctas, int a_op, int b_op, int passes, void* stream);

/* action_result bit-fields (results[N][8]); reference dict keys in comments */
#define FJSP_RES_SUCCESS 0x01        /* 'success' */
#define FJSP_RES_PS_LOADED 0x02      /* 'product_loaded'    PickupStationAgent.py:197 */
#define FJSP_RES_PS_TRAY_DONE 0x04   /* 'tray_completed'    PickupStationAgent.py:207,215 */
#define FJSP_RES_PS_IDLE_ORDERS 0x08 /* 'idle_with_orders'  PickupStationAgent.py:163 */
#define FJSP_RES_AGV_INVALID 0x02    /* 'invalid_action'    AGVAgent.py:211 */
#define FJSP_RES_AGV_MOVED 0x04      /* 'moved'             AGVAgent.py:240 */
#define FJSP_RES_AGV_PICKUP 0x08     /* 'pickup_success'    AGVAgent.py:288 */
#define FJSP_RES_AGV_DROP 0x10       /* 'drop_success'      AGVAgent.py:366 */
#define FJSP_RES_AGV_TO_PACK 0x20    /* 'delivered_to_packaging' AGVAgent.py:357 */
#define FJSP_RES_M_STARTED 0x02      /* 'started_processing' / 'started_packaging' */
#define FJSP_RES_M_COMPLETED 0x04    /* 'completed_processing' / 'completed_packaging' */
#define FJSP_RES_M_IDLE_QUEUE 0x08   /* 'idle_with_queue' */

#ifdef __cplusplus
}
#endif
#endif /* FJSP_B200_H */
