// fjsp_core.h — packed per-env state (128 x u32 = 512 B) and the lockstep step function.
//
// This is the device code of the hot path (FJSPSimulation.step, /root/reference/FJSPSimulation.py:144-242,
// and everything it calls in agents/*.py, models/*.py, utils/RewardModel.py).  It is written against an
// abstract word accessor S { u32 ld(int w); void st(int w, u32 v); u32 ld_hot(int w); void st_hot(int w, u32 v); }
// (ld/st: the dynamically indexed words 24..127 — completion steps, orders, tray pool; ld_hot/st_hot: the 24 hot words,
// always addressed with compile-time indices) so the same source runs
//   * in the CUDA kernels with S = one column of a shared-memory tile  (fjsp_kernels.cu), and
//   * in tests/host_harness (g++ build, tests only) with S = a plain array,
// which lets the CPU-only test suite exercise the packed-state logic against the oracle.  The product
// library has no CPU execution path: libfjsp_b200.so only ever launches the kernels.
//
// The algorithm is NOT a translation of the reference's SimPy object graph; it is an integer tick model:
//   - orders live in one u32 each (n, type, colour, tray-cut mask, packaged mask),
//   - trays waiting at the pickup station are implicit (a cursor over the cut masks),
//   - trays in transit live in a 64-slot pool of u32 records linked into FIFOs by 6-bit next indices,
//   - SimPy timers become start/finish step stamps resolved in the run phase of the step they fire in
//     (rule R0 of SURVEY.md §8a: a timeout of T time units started in step k fires in step k + T/step_size).
#ifndef FJSP_CORE_H
#define FJSP_CORE_H

#include <stdint.h>

#include "../../include/fjsp_b200.h"

#if defined(__CUDACC__)
#define FJSP_HD __host__ __device__ __forceinline__
#else
#define FJSP_HD inline
#endif

namespace fjsp {

typedef uint32_t u32;

// ---------------------------------------------------------------------------------------------
// Derived configuration, passed by value to every kernel (constants.py:5-32 after validation).
// ---------------------------------------------------------------------------------------------
struct Params {
    int32_t pos_row[FJSP_NUM_LOCATIONS], pos_col[FJSP_NUM_LOCATIONS];
    int32_t dist[FJSP_NUM_LOCATIONS][FJSP_NUM_LOCATIONS];   // Manhattan distance (AGVAgent.py:229)
    int32_t delay[FJSP_NUM_LOCATIONS][FJSP_NUM_LOCATIONS];  // floor(d / (agv_speed*step_size)) steps until arrival
    int32_t small_steps, big_steps, pack_steps;             // PROCESSING_TIMES / step_size
    int32_t max_episode_steps, storage_capacity, pack_capacity, trays_total, num_trays;
    int32_t step_size;
    int32_t arrival_q16, arrival_max;                       // long order streams: P(arrival per step) * 65536, cap on num_orders
    float progress_tab[256];                                // float32((1/L)*100) (PackagingAgent.py:117)
};

// location indices = LocationType order (enums/LocationType.py:3-8)
enum { LOC_PICKUP = 0, LOC_BIG = 1, LOC_SMALL = 2, LOC_STORAGE = 3, LOC_PACKAGING = 4 };
enum { TYPE_SMALL = 1, TYPE_MEDIUM = 2, TYPE_BIG = 3 };
enum { COL_RED = 1, COL_BLUE = 2, COL_GREEN = 3 };

// ---------------------------------------------------------------------------------------------
// Word map of the packed state of a K-cell shop.  Two layouts share the code below:
//
// COMPACT (LONG = false; K = 1: the reference shop, 128 words = 512 B): at most 32 orders and 240 steps per episode.
//   words 0..23      hot words of the pickup station (shared) and of cell 0
//   words 24..31     order completion steps (u8 each)
//   words 32..63     order words (the whole order table, generated at reset)
//   words 64+64c ..  tray pool of cell c (64 records)                      -> words 24 .. 64+64K are "dynamic"
//   words 64+64K ..  hot words of cells 1..K-1, 20 each (AGV, free bitmap x2, storage, 2 machines x 2, 4 stations x 3)
//
// LONG (long order streams, BASELINE configs[4]; K = 1: 292 words = 1168 B + a 4 KB side FIFO that a step touches in at
// most two words): up to 4095 orders and 65,000 steps per episode, Philox order ARRIVALS.  An order goes through three
// stages and only the last one costs packed state:
//   queued   (not popped yet)                 nothing stored: its attributes come from the Philox stream / the explicit
//                                             table when the pickup station pops it;
//   loaded   (trays waiting at the station)   one word per TRAY in the READY FIFO, a side buffer in HBM outside the tile
//                                             (order12 | first4<<12 | count3<<16, index = the tray's allocation index: the
//                                             reference hands out at most min(num_trays, 1000) trays per episode,
//                                             FJSPSimulation.py:96, so 1024 entries always suffice);
//   in process (an AGV took its first tray)   one of 64 ORDER SLOTS (three words: order word as in the compact layout,
//                                             processed bits + order id, completion step + first tray's allocation index),
//                                             given back one step after the order completes.  More than 64 orders in
//                                             process at once -> FJSP_FAULT_ORDER_SLOTS (an order whose tray was lost to one
//                                             of the reference's quirks never completes and keeps its slot).
// Counters are wider than in the compact layout; lost trays give their pool slot back at once.
//   words 0..23      as above (three shared words re-packed, see load_hot)
//   words 24..27     more shared hot words; 28..31 one more word per packaging station of cell 0 (current product)
//   words 32..35     order-slot bitmaps: free (2), freed by this step (2)
//   words 36..227    order slots: A 36..99, B 100..163, C 164..227
//   words 228+64c .. tray pools                                            -> words 32 .. 228+64K are "dynamic"
//   then             hot words of cells 1..K-1, 24 each
// ---------------------------------------------------------------------------------------------
enum {
    W_CTRL = 0,     // compact: step16 | num_orders6<<16 | fault3<<22 | completed_orders6<<25
                    // long   : step16 | fault3<<16 | cur_tray_count3<<19 | prod_idx4<<22
    W_PS = 1,       // compact: total_packaged9 | next_order6<<9 | cur_order6<<15 (63=none) | prod_idx4<<21 | cur_tray_count3<<25
                    // long   : num_orders12 | completed_orders12<<12 | dock_mask4<<24
    W_PSQ = 2,      // compact: alloc_count8 | ready_count8<<8 | ready_order6<<16 | ready_idx4<<22 | dock_mask4<<26
                    // long   : next_order12 | cur_order13<<12 (0x1fff = none)
    W_AGV = 3,      // loc3 | moving1<<3 | target3<<4 | arrive16<<7 | carry7<<23 (slot+1, 0 = none)
    W_FREE_LO = 4,  // pool free bitmap, slots 0..31 (1 = free)
    W_FREE_HI = 5,  // slots 32..63
    W_EPISODE = 6,  // Philox episode counter
    W_STORAGE = 7,  // head6 | tail6<<6 | len8<<12
    W_MACH = 8,     // 2 machines x 2 words (small, big)
                    //   A: busy1 | has_cur1<<1 | cur6<<2 | start16<<8 | prog1<<24 | qlen6<<25
                    //   B: qhead6 | qtail6<<6 | rhead6<<12 | rtail6<<18 | rlen6<<24
    W_PACK = 12,    // 4 stations x 3 words (blue_1, blue_2, red, green)
                    //   A: qhead6 | qtail6<<6 | qrec6<<12 | fhead6<<18 | ftail6<<24
                    //   B: frec6 | users5<<6 | busy1<<11 | waiters1<<12 | hascur1<<13 | compact: cur_slot5<<14 | cur_idx4<<19 | qcount8<<23
                    //                                                                  | long   : qcount8<<14
                    //   C: compact: completed8 | progL8<<8          long: completed16 | progL8<<16
                    //   D (long only): cur_order12 | cur_idx4<<12
    W_LONG_PSQ = 24,   // long: ready_head12 (trays taken from the ready FIFO so far) | ready_count12<<12
    W_LONG_CNT = 25,   // long: total_packaged16 | alloc_count12<<16
    W_LONG_ACT = 26,   // long: the order whose tray was taken last: order13 (0x1fff = none) | its slot6<<13 ;
                       //       cur_word8<<19 = n4 | type2 | colour2 of the order being loaded
    W_LONG_PACKD = 28, // long: word D of cell 0's four packaging stations
    CELL_HOT_WORDS = 20
};
// tray record (both layouts): order slot7 | first4<<7 | count3<<11 | processed1<<14 | next6<<15 | stamp8<<21 | lost1<<29
// order word / slot word A  : n4 | type2<<4 | colour2<<6 | cut8<<8 | packaged9<<16
// long: slot word B = processed9 | order12<<9;  slot word C = completion step + 1 (16) | first tray's allocation index 12 << 16

template <bool LONG>
struct WM {
    static constexpr int SLOTS = LONG ? 64 : 32;                       // order slots (compact: the whole order table, slot = order id)
    static constexpr int W_DYN0 = LONG ? 32 : 24;                      // first dynamically indexed word
    static constexpr int W_SLOT_FREE = 32, W_SLOT_FREED = 34;          // long only: bitmaps (2 words each)
    static constexpr int W_CSTEP = 24;                                 // compact only: u8 completion step + 1 per order
    static constexpr int W_ORDER = LONG ? 36 : 32;                     // order words / slot words A
    static constexpr int W_ORDER_B = W_ORDER + SLOTS, W_ORDER_C = W_ORDER + 2 * SLOTS;   // long only
    static constexpr int W_POOL = W_ORDER + (LONG ? 3 : 1) * SLOTS;    // 64 / 228
    static constexpr int CELL_HOT = LONG ? 24 : 20;
};
static_assert(WM<false>::W_POOL + 64 == FJSP_STATE_WORDS, "compact state size");
enum { W_CSTEP = WM<false>::W_CSTEP, W_ORDER = WM<false>::W_ORDER, W_POOL = WM<false>::W_POOL, W_TOTAL = 128 };  // compact names (host tools)
constexpr int READY_FIFO_WORDS = FJSP_LONG_READY_FIFO;  // side buffer per env (long layout)

template <int K, bool LONG = false>
struct Lay {
    static constexpr int DYN0 = WM<LONG>::W_DYN0;
    static constexpr int DYN_END = WM<LONG>::W_POOL + 64 * K;       // hot | dynamic (DYN0..DYN_END) | hot words of cells >= 1
    static constexpr int TOTAL = DYN_END + WM<LONG>::CELL_HOT * (K - 1);
    static constexpr int AGENTS = FJSP_AGENTS_K(K), ACT = FJSP_ACT_DIM_K(K), OBS = FJSP_OBS_DIM_K(K), MASK = FJSP_MASK_DIM_K(K);
    static_assert(LONG || TOTAL == FJSP_STATE_WORDS_K(K), "compact state size");
    static_assert(!LONG || TOTAL == FJSP_STATE_WORDS_LONG_K(K), "long state size");
};
// word index of hot word `which` (0 AGV, 1 FREE_LO, 2 FREE_HI, 3 STORAGE, 4..7 MACH, 8..19 PACK A/B/C, long: 20..23 PACK D) of cell c
template <int K, bool LONG>
FJSP_HD constexpr int cell_word(int c, int which) {
    return c == 0 ? (which < 3 ? W_AGV + which : which == 3 ? W_STORAGE : which < 8 ? W_MACH + (which - 4)
                     : which < 20 ? W_PACK + (which - 8) : W_LONG_PACKD + (which - 20))
                  : Lay<K, LONG>::DYN_END + WM<LONG>::CELL_HOT * (c - 1) + which;
}
template <bool LONG>
FJSP_HD constexpr int pool_base(int c) { return WM<LONG>::W_POOL + 64 * c; }
// ready-FIFO entry (long layout)
FJSP_HD u32 make_rq(int order, int first, int count) { return (u32)order | ((u32)first << 12) | ((u32)count << 16); }
FJSP_HD int rq_order(u32 e) { return (int)(e & 0xfffu); }
FJSP_HD int rq_first(u32 e) { return (int)((e >> 12) & 15u); }
FJSP_HD int rq_count(u32 e) { return (int)((e >> 16) & 7u); }

// order word
FJSP_HD int ord_n(u32 w) { return (int)(w & 15u); }
FJSP_HD int ord_type(u32 w) { return (int)((w >> 4) & 3u); }
FJSP_HD int ord_colour(u32 w) { return (int)((w >> 6) & 3u); }
FJSP_HD u32 ord_cut(u32 w) { return (w >> 8) & 0xffu; }
FJSP_HD u32 ord_packaged(u32 w) { return (w >> 16) & 0x1ffu; }
FJSP_HD u32 make_order(int n, int type, int colour) { return (u32)n | ((u32)type << 4) | ((u32)colour << 6); }

// tray record
FJSP_HD int rec_order(u32 r) { return (int)(r & 127u); }   // ring slot of the tray's order
FJSP_HD int rec_first(u32 r) { return (int)((r >> 7) & 15u); }
FJSP_HD int rec_count(u32 r) { return (int)((r >> 11) & 7u); }
FJSP_HD int rec_processed(u32 r) { return (int)((r >> 14) & 1u); }
FJSP_HD int rec_next(u32 r) { return (int)((r >> 15) & 63u); }
FJSP_HD int rec_stamp(u32 r) { return (int)((r >> 21) & 255u); }
FJSP_HD int rec_lost(u32 r) { return (int)((r >> 29) & 1u); }
constexpr u32 REC_PROCESSED = 1u << 14, REC_LOST = 1u << 29;
FJSP_HD u32 make_rec(int slot, int first, int count, int processed) {
    return (u32)slot | ((u32)first << 7) | ((u32)count << 11) | ((u32)processed << 14);
}
FJSP_HD u32 rec_with_next(u32 r, int next) { return (r & ~(63u << 15)) | ((u32)next << 15); }
FJSP_HD u32 rec_with_stamp(u32 r, int stamp) { return (r & ~(255u << 21)) | ((u32)(stamp & 255) << 21); }
FJSP_HD u32 rec_split_rest(u32 r, int first, int count) { return (r & ~((15u << 7) | (7u << 11))) | ((u32)first << 7) | ((u32)count << 11); }

FJSP_HD int ctz32(u32 x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
FJSP_HD int popc32(u32 x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// ---------------------------------------------------------------------------------------------
// Unpacked hot scalars.  Shared (pickup station, counters) + one struct per cell.  Kept in registers.
// ---------------------------------------------------------------------------------------------
// Bit-fields: the compiler keeps each group in ONE register and extracts / inserts a field where it is used (the state
// of a cell is 26 registers this way instead of 92 — the cell-parallel kernel lives inside 128 registers per thread
// without spilling to local memory, which would cost HBM bandwidth).  Widths are those of the packed words or wider.
struct Fifo {
    u32 head : 8, tail : 8, len : 8;
};
struct Mach {
    u32 busy : 1, has_cur : 1, prog : 1, cur : 8, start : 16;
    Fifo q, r;
};
struct Pack {
    Fifo q, f;  // queued tray records, in-flight tray records (len = record count)
    u32 users : 6, busy : 1, waiters : 1, hascur : 1, cur_idx : 4, cur_order : 12;
    // current_product (sticky): cur_order compact = ring slot (= order id), long = order id; cur_idx = product index
    u32 qcount : 8, progL : 8, completed : 16;
};
struct Hot {  // shared part
    u32 step : 16, num_orders : 12;
    int fault : 4;                // (-1 is used as "none raised" inside the cell-parallel step)
    u32 completed_orders : 12, total_packaged : 16, dock_mask : 4;
    u32 next_order : 12, prod_idx : 4;
    int cur_order : 14;           // -1 = none
    u32 cur_tray_count : 3, alloc_count : 12, ready_count : 12;
    u32 ready_order : 12, ready_idx : 4;   // long: ready_order = ready_head, ready_idx unused
    u32 cur_word : 8;             // long: n4 | type2<<4 | colour2<<6 of the order being loaded
    int act_order : 14;           // long: the order whose tray was taken last (-1 = none) and its slot
    u32 act_slot : 6;
    u32 episode;
};
struct HotCell {
    u32 agv_loc : 3, agv_moving : 1, agv_target : 3, agv_arrive : 16, carry : 8;
    u32 free_lo, free_hi;
    Fifo storage;
    Mach m[2];
    Pack p[4];
};

template <class S>
FJSP_HD void load_hot(S& s, Hot& h) {
    if (S::LONG) {
        u32 w = s.ld_hot(W_CTRL);
        h.step = (int)(w & 0xffffu), h.fault = (int)((w >> 16) & 7u), h.cur_tray_count = (int)((w >> 19) & 7u);
        h.prod_idx = (int)((w >> 22) & 15u), h.ready_idx = 0;
        w = s.ld_hot(W_PS);
        h.num_orders = (int)(w & 0xfffu), h.completed_orders = (int)((w >> 12) & 0xfffu), h.dock_mask = (int)((w >> 24) & 15u);
        w = s.ld_hot(W_PSQ);
        h.next_order = (int)(w & 0xfffu);
        h.cur_order = (int)((w >> 12) & 0x1fffu);
        if (h.cur_order == 0x1fff) h.cur_order = -1;
        w = s.ld_hot(W_LONG_PSQ);
        h.ready_order = (int)(w & 0xfffu), h.ready_count = (int)((w >> 12) & 0xfffu);
        w = s.ld_hot(W_LONG_CNT);
        h.total_packaged = (int)(w & 0xffffu), h.alloc_count = (int)((w >> 16) & 0xfffu);
        w = s.ld_hot(W_LONG_ACT);
        h.act_order = (int)(w & 0x1fffu), h.act_slot = (int)((w >> 13) & 63u), h.cur_word = (int)((w >> 19) & 255u);
        if (h.act_order == 0x1fff) h.act_order = -1;
    } else {
        u32 w = s.ld_hot(W_CTRL);
        h.step = (int)(w & 0xffffu), h.num_orders = (int)((w >> 16) & 63u), h.fault = (int)((w >> 22) & 7u);
        h.completed_orders = (int)((w >> 25) & 63u);
        w = s.ld_hot(W_PS);
        h.total_packaged = (int)(w & 511u), h.next_order = (int)((w >> 9) & 63u), h.cur_order = (int)((w >> 15) & 63u);
        if (h.cur_order == 63) h.cur_order = -1;
        h.prod_idx = (int)((w >> 21) & 15u), h.cur_tray_count = (int)((w >> 25) & 7u);
        w = s.ld_hot(W_PSQ);
        h.alloc_count = (int)(w & 255u), h.ready_count = (int)((w >> 8) & 255u), h.ready_order = (int)((w >> 16) & 63u);
        h.ready_idx = (int)((w >> 22) & 15u), h.dock_mask = (int)((w >> 26) & 15u);
    }
    h.episode = s.ld_hot(W_EPISODE);
}
template <class S>
FJSP_HD void store_hot(S& s, const Hot& h) {
    if (S::LONG) {
        s.st_hot(W_CTRL, (u32)h.step | ((u32)h.fault << 16) | ((u32)h.cur_tray_count << 19) | ((u32)h.prod_idx << 22));
        s.st_hot(W_PS, (u32)h.num_orders | ((u32)h.completed_orders << 12) | ((u32)h.dock_mask << 24));
        s.st_hot(W_PSQ, (u32)h.next_order | ((u32)(h.cur_order < 0 ? 0x1fff : h.cur_order) << 12));
        s.st_hot(W_LONG_PSQ, (u32)h.ready_order | ((u32)h.ready_count << 12));
        s.st_hot(W_LONG_CNT, (u32)h.total_packaged | ((u32)h.alloc_count << 16));
        s.st_hot(W_LONG_ACT, (u32)(h.act_order < 0 ? 0x1fff : h.act_order) | ((u32)h.act_slot << 13) | ((u32)h.cur_word << 19));
    } else {
        s.st_hot(W_CTRL, (u32)h.step | ((u32)h.num_orders << 16) | ((u32)h.fault << 22) | ((u32)h.completed_orders << 25));
        s.st_hot(W_PS, (u32)h.total_packaged | ((u32)h.next_order << 9) | ((u32)(h.cur_order < 0 ? 63 : h.cur_order) << 15) |
                           ((u32)h.prod_idx << 21) | ((u32)h.cur_tray_count << 25));
        s.st_hot(W_PSQ, (u32)h.alloc_count | ((u32)h.ready_count << 8) | ((u32)h.ready_order << 16) | ((u32)h.ready_idx << 22) |
                            ((u32)h.dock_mask << 26));
    }
    s.st_hot(W_EPISODE, h.episode);
}
template <int K, class S>
FJSP_HD void load_cell(S& s, int c, HotCell& h) {
    constexpr bool LONG = S::LONG;
    u32 w = s.ld_hot(cell_word<K, LONG>(c, 0));
    h.agv_loc = (int)(w & 7u), h.agv_moving = (int)((w >> 3) & 1u), h.agv_target = (int)((w >> 4) & 7u);
    h.agv_arrive = (int)((w >> 7) & 0xffffu), h.carry = (int)((w >> 23) & 127u);
    h.free_lo = s.ld_hot(cell_word<K, LONG>(c, 1)), h.free_hi = s.ld_hot(cell_word<K, LONG>(c, 2));
    w = s.ld_hot(cell_word<K, LONG>(c, 3));
    h.storage.head = (int)(w & 63u), h.storage.tail = (int)((w >> 6) & 63u), h.storage.len = (int)((w >> 12) & 255u);
#pragma unroll
    for (int i = 0; i < 2; i++) {
        Mach& m = h.m[i];
        w = s.ld_hot(cell_word<K, LONG>(c, 4 + 2 * i));
        m.busy = (int)(w & 1u), m.has_cur = (int)((w >> 1) & 1u), m.cur = (int)((w >> 2) & 63u);
        m.start = (int)((w >> 8) & 0xffffu), m.prog = (int)((w >> 24) & 1u), m.q.len = (int)((w >> 25) & 63u);
        w = s.ld_hot(cell_word<K, LONG>(c, 5 + 2 * i));
        m.q.head = (int)(w & 63u), m.q.tail = (int)((w >> 6) & 63u), m.r.head = (int)((w >> 12) & 63u);
        m.r.tail = (int)((w >> 18) & 63u), m.r.len = (int)((w >> 24) & 63u);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        Pack& p = h.p[i];
        w = s.ld_hot(cell_word<K, LONG>(c, 8 + 3 * i));
        p.q.head = (int)(w & 63u), p.q.tail = (int)((w >> 6) & 63u), p.q.len = (int)((w >> 12) & 63u);
        p.f.head = (int)((w >> 18) & 63u), p.f.tail = (int)((w >> 24) & 63u);
        w = s.ld_hot(cell_word<K, LONG>(c, 9 + 3 * i));
        p.f.len = (int)(w & 63u), p.users = (int)((w >> 6) & 31u), p.busy = (int)((w >> 11) & 1u);
        p.waiters = (int)((w >> 12) & 1u), p.hascur = (int)((w >> 13) & 1u);
        const u32 wc = s.ld_hot(cell_word<K, LONG>(c, 10 + 3 * i));
        if (LONG) {
            p.qcount = (int)((w >> 14) & 255u);
            p.completed = (int)(wc & 0xffffu), p.progL = (int)((wc >> 16) & 255u);
            const u32 wd = s.ld_hot(cell_word<K, LONG>(c, 20 + i));
            p.cur_order = (int)(wd & 0xfffu), p.cur_idx = (int)((wd >> 12) & 15u);
        } else {
            p.cur_order = (int)((w >> 14) & 31u), p.cur_idx = (int)((w >> 19) & 15u), p.qcount = (int)((w >> 23) & 255u);
            p.completed = (int)(wc & 255u), p.progL = (int)((wc >> 8) & 255u);
        }
    }
}
template <int K, class S>
FJSP_HD void store_cell(S& s, int c, const HotCell& h) {
    constexpr bool LONG = S::LONG;
    s.st_hot(cell_word<K, LONG>(c, 0), (u32)h.agv_loc | ((u32)h.agv_moving << 3) | ((u32)h.agv_target << 4) | ((u32)h.agv_arrive << 7) |
                                           ((u32)h.carry << 23));
    s.st_hot(cell_word<K, LONG>(c, 1), h.free_lo), s.st_hot(cell_word<K, LONG>(c, 2), h.free_hi);
    s.st_hot(cell_word<K, LONG>(c, 3), (u32)h.storage.head | ((u32)h.storage.tail << 6) | ((u32)h.storage.len << 12));
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const Mach& m = h.m[i];
        s.st_hot(cell_word<K, LONG>(c, 4 + 2 * i), (u32)m.busy | ((u32)m.has_cur << 1) | ((u32)m.cur << 2) | ((u32)m.start << 8) |
                                                       ((u32)m.prog << 24) | ((u32)m.q.len << 25));
        s.st_hot(cell_word<K, LONG>(c, 5 + 2 * i), (u32)m.q.head | ((u32)m.q.tail << 6) | ((u32)m.r.head << 12) | ((u32)m.r.tail << 18) |
                                                       ((u32)m.r.len << 24));
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const Pack& p = h.p[i];
        s.st_hot(cell_word<K, LONG>(c, 8 + 3 * i), (u32)p.q.head | ((u32)p.q.tail << 6) | ((u32)p.q.len << 12) | ((u32)p.f.head << 18) |
                                                       ((u32)p.f.tail << 24));
        const u32 b = (u32)p.f.len | ((u32)p.users << 6) | ((u32)p.busy << 11) | ((u32)p.waiters << 12) | ((u32)p.hascur << 13);
        if (LONG) {
            s.st_hot(cell_word<K, LONG>(c, 9 + 3 * i), b | ((u32)p.qcount << 14));
            s.st_hot(cell_word<K, LONG>(c, 10 + 3 * i), (u32)(p.completed & 0xffff) | ((u32)p.progL << 16));
            s.st_hot(cell_word<K, LONG>(c, 20 + i), (u32)p.cur_order | ((u32)p.cur_idx << 12));
        } else {
            s.st_hot(cell_word<K, LONG>(c, 9 + 3 * i), b | ((u32)p.cur_order << 14) | ((u32)p.cur_idx << 19) | ((u32)p.qcount << 23));
            s.st_hot(cell_word<K, LONG>(c, 10 + 3 * i), (u32)(p.completed & 255) | ((u32)p.progL << 8));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Pool + FIFO primitives (pb = first pool word of the cell)
// ---------------------------------------------------------------------------------------------
FJSP_HD int pool_alloc(HotCell& h) {
    if (h.free_lo) {
        int b = ctz32(h.free_lo);
        h.free_lo &= h.free_lo - 1u;
        return b;
    }
    if (h.free_hi) {
        int b = ctz32(h.free_hi);
        h.free_hi &= h.free_hi - 1u;
        return 32 + b;
    }
    return -1;
}
FJSP_HD void pool_free(HotCell& h, int slot) {
    if (slot < 32) h.free_lo |= 1u << slot;
    else h.free_hi |= 1u << (slot - 32);
}
template <class S>
FJSP_HD void fifo_push(S& s, int pb, Fifo& f, int slot) {
    if (f.len == 0) {
        f.head = slot;
    } else {
        u32 r = s.ld(pb + f.tail);
        s.st(pb + f.tail, rec_with_next(r, slot));
    }
    f.tail = slot;
    f.len++;
}
template <class S>
FJSP_HD int fifo_pop(S& s, int pb, Fifo& f) {  // caller checks len > 0
    int slot = f.head;
    f.head = rec_next(s.ld(pb + slot));
    f.len--;
    return slot;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based; replayable on the host).  Streams:
//   orders : counter = (global env, episode, order index, 0) -> n = 1+mulhi(r0,9), type = 1+mulhi(r1,3), colour = 1+mulhi(r2,3)
//   actions: counter = (global env, t_lo, t_hi, 1)           -> a_j = (u16_j * n_j) >> 16
// ---------------------------------------------------------------------------------------------
FJSP_HD u32 mulhi_u32(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (u32)(((uint64_t)a * b) >> 32);
#endif
}
FJSP_HD void philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32 out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        u32 hi0 = mulhi_u32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        u32 hi1 = mulhi_u32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        u32 n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}
FJSP_HD u32 philox_order(uint64_t seed, uint64_t genv, u32 episode, int o) {
    u32 r[4];
    philox4x32_10((u32)genv, episode, (u32)o, 0u, (u32)seed, (u32)(seed >> 32), r);
    return make_order(1 + (int)mulhi_u32(r[0], 9u), 1 + (int)mulhi_u32(r[1], 3u), 1 + (int)mulhi_u32(r[2], 3u));
}
FJSP_HD void philox_actions(uint64_t seed, uint64_t genv, uint64_t t, int a[8]) {
    u32 r[4];
    philox4x32_10((u32)genv, (u32)t, (u32)(t >> 32), 1u, (u32)seed, (u32)(seed >> 32), r);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        u32 hw = (j & 1) ? (r[j >> 1] >> 16) : (r[j >> 1] & 0xffffu);
        a[j] = (int)((hw * (j == 1 ? 8u : 3u)) >> 16);
    }
}

// scaled shop: cell c draws its 8 values with counter word 3 = 1 + 16c; cell 0 also supplies the pickup station's
template <int K>
FJSP_HD void philox_actions_k(uint64_t seed, uint64_t genv, uint64_t t, int* a) {
#pragma unroll
    for (int c = 0; c < K; c++) {
        u32 r[4];
        philox4x32_10((u32)genv, (u32)t, (u32)(t >> 32), 1u + 16u * (u32)c, (u32)seed, (u32)(seed >> 32), r);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (c > 0 && j == 0) continue;
            u32 hw = (j & 1) ? (r[j >> 1] >> 16) : (r[j >> 1] & 0xffffu);
            a[7 * c + j] = (int)((hw * (j == 1 ? 8u : 3u)) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Reset: FJSPSimulation.reset (FJSPSimulation.py:286-323).  `orders` = explicit FjspOrderRec table (n | type<<8 |
// colour<<16; 32 records in the compact layout, `num_orders` in the long one) or nullptr -> Philox.
// ---------------------------------------------------------------------------------------------
// explicit record -> order word; n in 1..9 products, type and colour in 1..3 (FJSPSimulation.py:107-112).  Anything else
// would overflow its bit-field in the order word, so it is clamped and reported (`bad`).
FJSP_HD u32 order_from_rec(u32 r, bool& bad) {
    int n = (int)(r & 0xffu), ty = (int)((r >> 8) & 0xffu), co = (int)((r >> 16) & 0xffu);
    if (n < 1 || n > FJSP_MAX_ORDER_PRODUCTS || ty < 1 || ty > 3 || co < 1 || co > 3 || (r >> 24)) {
        bad = true;
        n = n < 1 ? 1 : n > FJSP_MAX_ORDER_PRODUCTS ? FJSP_MAX_ORDER_PRODUCTS : n;
        ty = ty < 1 ? 1 : ty > 3 ? 3 : ty, co = co < 1 ? 1 : co > 3 ? 3 : co;
    }
    return make_order(n, ty, co);
}

// Everything of a fresh env except the compact layout's 32 order words (the kernels fill those warp-cooperatively, one
// order per lane; the long layout's ring fills as the pickup station pops orders).
template <int K, class S>
FJSP_HD void reset_env_base(S& s, int num_orders, u32 episode, int fault = 0) {
    constexpr bool LONG = S::LONG;
    using L = Lay<K, LONG>;
#pragma unroll
    for (int w = 0; w < L::DYN0; w++) s.st_hot(w, 0u);
#pragma unroll 8
    for (int w = L::DYN0; w < L::DYN_END; w++)
        if (LONG || w < WM<LONG>::W_ORDER || w >= WM<LONG>::W_POOL) s.st(w, 0u);
#pragma unroll
    for (int w = L::DYN_END; w < L::TOTAL; w++) s.st_hot(w, 0u);
    Hot h;
    h.step = 0, h.num_orders = num_orders, h.fault = fault, h.completed_orders = 0;
    h.total_packaged = 0, h.next_order = 0, h.cur_order = -1, h.prod_idx = 0, h.cur_tray_count = 0;
    h.alloc_count = 0, h.ready_count = 0, h.ready_order = 0, h.ready_idx = 0;
    h.cur_word = 0, h.act_order = -1, h.act_slot = 0;
    h.dock_mask = 1;   // the dock of the pickup station is held by cell 0's AGV
    h.episode = episode;
    store_hot(s, h);
    if (LONG) s.st(WM<LONG>::W_SLOT_FREE, 0xffffffffu), s.st(WM<LONG>::W_SLOT_FREE + 1, 0xffffffffu);  // every order slot is free
#pragma unroll
    for (int c = 0; c < K; c++) {
        // AGVAgent.py:41: the AGV starts at PICKUP; further cells' AGVs start at STORAGE (one dock)
        s.st_hot(cell_word<K, LONG>(c, 0), (u32)(c == 0 ? LOC_PICKUP : LOC_STORAGE));
        s.st_hot(cell_word<K, LONG>(c, 1), 0xffffffffu), s.st_hot(cell_word<K, LONG>(c, 2), 0xffffffffu);
    }
}

template <int K, class S>
FJSP_HD void reset_env(S& s, const Params& P, int num_orders, const FjspOrderRec* orders, uint64_t seed, uint64_t genv,
                       u32 episode) {
    (void)P;
    constexpr bool LONG = S::LONG;
    bool bad = false;
    if (LONG) {
        if (orders)
            for (int o = 0; o < num_orders; o++) (void)order_from_rec(orders[o], bad);
        reset_env_base<K>(s, num_orders, episode, bad ? FJSP_FAULT_BAD_ORDER : 0);
        return;
    }
    reset_env_base<K>(s, num_orders, episode);
    for (int o = 0; o < FJSP_MAX_ORDERS; o++) {
        u32 ow = 0u;
        if (o < num_orders) ow = orders ? order_from_rec(orders[o], bad) : philox_order(seed, genv, episode, o);
        s.st(WM<LONG>::W_ORDER + o, ow);
    }
    if (bad) reset_env_base<K>(s, num_orders, episode, FJSP_FAULT_BAD_ORDER);  // (rare path: rewrite the scalars with the fault set)
}

// ---------------------------------------------------------------------------------------------
// Observation (layout O: 7 + 31K floats) + masks (3 + 26K bytes).  SURVEY.md §8a-R9; K = 1: 38 / 29.
// ---------------------------------------------------------------------------------------------
// Where an observation goes.  FloatSink: the reference's float32 row (a2c._flatten_obs order).  FieldSink: the plain
// integers — a station index for the two AGV position fields, the table index L for a packaging progress — from which
// the compact host wire format (include/fjsp_b200.h "wire rows") is packed.
struct FloatSink {
    float* o;
    const Params& P;
    FJSP_HD void set(int i, int v) { o[i] = (float)v; }
    FJSP_HD void set_i8(int i, int v) { o[i] = (float)(int)(int8_t)v; }   // dtype=np.int8 (PackagingAgent.py:59)
    FJSP_HD void set_loc(int i, int loc) { o[i] = (float)P.pos_row[loc], o[i + 1] = (float)P.pos_col[loc]; }
    FJSP_HD void set_prog(int i, int L) { o[i] = P.progress_tab[L]; }
    FJSP_HD FloatSink at(int off) const { return FloatSink{o + off, P}; }
};
struct FieldSink {   // the raw integer value of every field (the wire row packs them into bit-fields)
    int* f;
    FJSP_HD void set(int i, int v) { f[i] = v; }
    FJSP_HD void set_i8(int i, int v) { f[i] = v & 255; }
    FJSP_HD void set_loc(int i, int loc) { f[i] = loc, f[i + 1] = loc; }
    FJSP_HD void set_prog(int i, int L) { f[i] = L; }
    FJSP_HD FieldSink at(int off) const { return FieldSink{f + off}; }
};
enum { OBS_NONE = 0, OBS_FLOAT = 1, OBS_WIRE = 2 };

template <int K>
struct StepOut {
    float* obs;     // OBS_FLOAT: 7 + 31K floats, written in place (a shared-memory staging row on the device)
    u32 wire[FJSP_WIRE_WORDS_K(K)];   // OBS_WIRE: the observation part of the wire row (bit-fields; wire_row adds the rest)
    u32 mask[Lay<K>::MASK / 4];
    float reward[Lay<K>::ACT];
    int reward_g;                     // 10A*r_i = reward_g + A*reward_local10[i]  (exact integers)
    int d_orders, d_products;         // orders completed / products packaged by this step (reward_g = 10*(100*do + 10*dp) - step_size)
    int reward_local10[Lay<K>::ACT];
    u32 flags;      // terminated | truncated<<8 | fault<<16 | was_reset<<24
    u32 results[Lay<K>::ACT / 4];  // u8 action_result bit-fields per agent
    int32_t info[4];
    long long reward_units;  // sum over the agents of 10*A*reward (exact integers): statistic for rollouts
};

FJSP_HD void mask_set(u32* mw, int idx, int v) { mw[idx >> 2] |= (u32)(v & 1) << ((idx & 3) * 8); }

template <class S, class O>
FJSP_HD void observe_shared(S& s, const Params& P, const Hot& h, O obs, u32* mw) {
    // ---- pickup station: PickupStationAgent.get_observation (:58-98) / get_action_mask (:100-142)
    int has_cur_order = h.cur_order >= 0;
    int order_size = 0, remaining = 0, o_type = 0, o_colour = 0;
    if (has_cur_order) {
        u32 ow = S::LONG ? (u32)h.cur_word : s.ld(WM<S::LONG>::W_ORDER + h.cur_order);
        order_size = ord_n(ow), remaining = order_size - h.prod_idx;
        o_type = ord_type(ow), o_colour = ord_colour(ow);
    }
    int tcount = h.cur_tray_count;
    obs.set(0, tcount > 0 ? o_colour : 0);
    obs.set(1, tcount);
    obs.set(2, tcount > 0 ? o_type : 0);
    obs.set(3, remaining > 0 ? o_colour : 0);
    obs.set(4, remaining > 0 ? o_type : 0);
    obs.set(5, order_size);
    obs.set(6, remaining);
    int queue_len = h.num_orders - h.next_order;
    int has_order = has_cur_order || queue_len > 0;
    int has_tray = tcount > 0 || (P.trays_total - h.alloc_count) > 0;
    int tray_not_full = tcount < FJSP_TRAY_CAPACITY;
    int prem = has_cur_order ? (remaining > 0) : (queue_len > 0);
    mask_set(mw, 0, 1);
    mask_set(mw, 1, has_order && has_tray && tray_not_full && prem);
    mask_set(mw, 2, tcount > 0);
}

// the AGV whose fields are in `hc`, on the stations of `hc`: 13 floats, 8 mask bytes from `mo`.  SHARED (shared floor):
// `occ` = station positions taken by other AGVs; otherwise the scaled shop's one-dock rule.
template <bool SHARED, class S, class O>
FJSP_HD void observe_agv(S& s, const Params& P, const Hot& h, const HotCell& hc, int c, O obs, u32* mw, int mo, u32 occ) {
    (void)P;
    const int pb = pool_base<S::LONG>(c);
    // ---- AGV: AGVAgent.get_observation (:53-76) / get_action_mask (:79-178)
    int carrying = hc.carry != 0;
    int c_count = 0, c_type = 0, c_proc = 0;
    if (carrying) {
        u32 r = s.ld(pb + hc.carry - 1);
        c_count = rec_count(r), c_proc = rec_processed(r);
        c_type = ord_type(s.ld(WM<S::LONG>::W_ORDER + rec_order(r)));
    }
    obs.set(0, hc.m[1].busy);
    obs.set(1, hc.m[1].r.len);
    obs.set(2, carrying);
    obs.set(3, h.ready_count);
    obs.set_loc(4, hc.agv_loc);               // [4] = row, [5] = column of the AGV's station
    obs.set(6, hc.m[0].busy);
    obs.set(7, hc.m[0].r.len);
    obs.set(8, hc.storage.len);
    obs.set(9, carrying);                     // a carried tray is never packaged (delivered trays vanish)
    obs.set(10, carrying && !c_proc);
    obs.set(11, c_count);
    obs.set(12, c_type);
    mask_set(mw, mo + 0, 1);
    if (!hc.agv_moving) {
        int loc = hc.agv_loc;
        const u32 blocked = SHARED ? occ : (((h.dock_mask & ~(1 << c)) != 0) ? 1u << LOC_PICKUP : 0u);  // one dock (scaled shop)
        mask_set(mw, mo + 1, loc != LOC_PICKUP && !((blocked >> LOC_PICKUP) & 1u));
        mask_set(mw, mo + 2, loc != LOC_SMALL && !((blocked >> LOC_SMALL) & 1u));
        mask_set(mw, mo + 3, loc != LOC_BIG && !((blocked >> LOC_BIG) & 1u));
        mask_set(mw, mo + 4, loc != LOC_STORAGE && !((blocked >> LOC_STORAGE) & 1u));
        mask_set(mw, mo + 5, loc != LOC_PACKAGING && !((blocked >> LOC_PACKAGING) & 1u));
        if (!carrying) {
            int avail = loc == LOC_PICKUP ? h.ready_count
                        : loc == LOC_SMALL ? hc.m[0].r.len
                        : loc == LOC_BIG ? hc.m[1].r.len
                        : loc == LOC_STORAGE ? hc.storage.len : 0;
            mask_set(mw, mo + 6, avail > 0);
        } else {
            int ok = 0;
            if (loc == LOC_SMALL) ok = !c_proc && (c_type == TYPE_SMALL || c_type == TYPE_MEDIUM);
            else if (loc == LOC_BIG) ok = !c_proc && (c_type == TYPE_BIG || c_type == TYPE_MEDIUM);
            else if (loc == LOC_PACKAGING) ok = c_proc;
            else if (loc == LOC_STORAGE) ok = 1;
            mask_set(mw, mo + 7, ok);  // PICKUP: only an empty tray, never the case
        }
    }
}
// machines (3 + 3 floats) and packaging stations (12) of a cell; 18 mask bytes from `mo`
template <class O>
FJSP_HD void observe_stations(const Params& P, const HotCell& hc, O obs, u32* mw, int mo) {
    // ---- machines: MachineAgent.get_observation (:62-70) / get_action_mask (:72-97)
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const Mach& m = hc.m[i];
        obs.set(3 * i, m.busy);
        obs.set(1 + 3 * i, m.prog ? 1 : 0);
        obs.set(2 + 3 * i, m.q.len);
        mask_set(mw, mo + 3 * i, 1);
        mask_set(mw, mo + 1 + 3 * i, m.q.len > 0 && !m.busy);
        mask_set(mw, mo + 2 + 3 * i, !m.busy && m.has_cur);
    }
    // ---- packaging: PackagingAgent.get_observation (:54-62) / get_action_mask (:64-89)
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const Pack& p = hc.p[i];
        obs.set(6 + 3 * i, p.busy);
        obs.set_prog(7 + 3 * i, p.progL);
        obs.set_i8(8 + 3 * i, p.qcount);
        mask_set(mw, mo + 6 + 3 * i, 1);
        mask_set(mw, mo + 7 + 3 * i, p.qcount > 0 && !p.busy && p.users < P.pack_capacity);
        mask_set(mw, mo + 8 + 3 * i, !p.busy && p.hascur);
    }
}
// one cell: AGV (13) + small/big machine (3 + 3) + four packaging stations (12) = 31 floats; 26 mask bytes from `mo`
template <class S, class O>
FJSP_HD void observe_cell(S& s, const Params& P, const Hot& h, const HotCell& hc, int c, O obs, u32* mw, int mo) {
    observe_agv<false>(s, P, h, hc, c, obs, mw, mo, 0u);
    observe_stations(P, hc, obs.at(13), mw, mo + 8);
}

// whole observation; cell 0's hot words are the caller's registers, further cells are re-read
template <int K, class S, class O>
FJSP_HD void observe(S& s, const Params& P, const Hot& h, const HotCell& c0, O obs, u32* mw) {
#pragma unroll
    for (int i = 0; i < Lay<K>::MASK / 4; i++) mw[i] = 0u;
    observe_shared(s, P, h, obs, mw);
    observe_cell(s, P, h, c0, 0, obs.at(7), mw, 3);
#pragma unroll
    for (int c = 1; c < K; c++) {
        HotCell hc;
        load_cell<K>(s, c, hc);
        observe_cell(s, P, h, hc, c, obs.at(7 + 31 * c), mw, 3 + 26 * c);
    }
}
// ---------------------------------------------------------------------------------------------
// Wire row (include/fjsp_b200.h FJSP_WIRE_WORDS_K): everything a step returns for one env as BIT-FIELDS, 2 + 5K words
// (32 bytes for K = 1).  No field straddles a word, so the host decoder extracts each with one shift and one mask.
//   S0  ps obs: tray_color2 | tray_count3<<2 | tray_type2<<5 | next_color2<<7 | next_type2<<9 | order_size4<<11 |
//       remaining4<<15 ; flags6<<19 (terminated, truncated, fault x3, was_reset) ; ps mask bits 1,2 <<25 ; ps reward code2<<27
//   S1  products packaged by the step 10 | orders completed 9 <<10 | pickup_ready_trays 13 <<19
//   per cell c, words 2 + 5c ..:
//   C0  agv: station3 | carrying1<<3 | needs_processing1<<4 | tray_count3<<5 | tray_type2<<8 | big_ready6<<10 |
//       small_ready6<<16 | agv mask bits 1..7 <<22 | agv reward code3<<29
//   C1  storage8 | small machine (busy1, progress1, queue6)<<8 | big machine<<16 | small mask bits 1,2<<24 | big<<26 |
//       small reward code2<<28 | big<<30
//   C2  station0 | station3: progL8<<21, busy1<<29, mask bits2<<30
//   C3  station1 | station3: queue8<<21, reward code2<<29
//   C4  station2             with stationN = busy1 | progL8<<1 | queue8<<9 | mask bits 1,2 <<17 | reward code2<<19
// Mask bit 0 of every agent is constant 1 and not sent.  Reward codes index the agent kind's few possible local rewards
// (RewardModel.py:46-97); the global part follows from the two counters in S1.
// ---------------------------------------------------------------------------------------------
template <int K>
struct Wire {
    static constexpr int WORDS = FJSP_WIRE_WORDS_K(K), USED = 2 + 5 * K;
    static_assert(USED <= WORDS && WORDS % 2 == 0, "wire row layout");
};
// local rewards in tenths <-> codes
FJSP_HD u32 code_ps(int l) { return l == 0 ? 0u : l == -10 ? 1u : l == 10 ? 2u : 3u; }                 // 0, -1, +1, +6
FJSP_HD u32 code_agv(int l) { return l == 0 ? 0u : l == -1 ? 1u : l == 20 ? 2u : l == 120 ? 3u : 4u; } // 0, -0.1, +2, +12, -5
FJSP_HD u32 code_mach(int l) { return l == 0 ? 0u : l == -20 ? 1u : l == 10 ? 2u : 3u; }               // 0, -2, +1, +5
FJSP_HD u32 code_pack(int l) { return l == 0 ? 0u : l == -10 ? 1u : l == 20 ? 2u : 3u; }               // 0, -1, +2, +20
// reward LUT of the decoder: [kind 0 ps, 1 agv, 2 machine, 3 packaging][code], tenths
#define FJSP_WIRE_REWARD_LUT {0, -10, 10, 60, 0, 0, 0, 0,   0, -1, 20, 120, -50, 0, 0, 0,   0, -20, 10, 50, 0, 0, 0, 0,   0, -10, 20, 200, 0, 0, 0, 0}
// flags word of fjsp_step (terminated | truncated << 8 | fault << 16 | was_reset << 24) <-> the 6 flag bits of S0
FJSP_HD u32 wire_flag_bits(u32 flags) { return (flags & 1u) | ((flags >> 7) & 2u) | ((flags >> 14) & 0x1cu) | ((flags >> 19) & 0x20u); }
FJSP_HD u32 wire_flags(u32 s0) {
    const u32 fb = (s0 >> 19) & 63u;
    return (fb & 1u) | ((fb & 2u) << 7) | ((fb & 0x1cu) << 14) | ((fb & 0x20u) << 19);
}
// four mask bytes (0/1 each) -> four bits
FJSP_HD u32 mask_nibble(u32 bytes4) { return ((bytes4 & 0x01010101u) * 0x01020408u) >> 24; }

// observation parts: fs = the pickup station's 7 fields, psbits = its 3 mask bits; f = one cell's 31 fields, bits = its 26 mask bits
FJSP_HD u32 wire_s0_obs(const int* fs, u32 psbits) {
    return (u32)fs[0] | ((u32)fs[1] << 2) | ((u32)fs[2] << 5) | ((u32)fs[3] << 7) | ((u32)fs[4] << 9) | ((u32)fs[5] << 11) |
           ((u32)fs[6] << 15) | (((psbits >> 1) & 3u) << 25);
}
FJSP_HD void wire_cell_obs(const int* f, u32 bits, u32* w) {
    w[0] = (u32)f[4] | ((u32)f[2] << 3) | ((u32)f[10] << 4) | ((u32)f[11] << 5) | ((u32)f[12] << 8) | ((u32)f[1] << 10) |
           ((u32)f[7] << 16) | (((bits >> 1) & 0x7fu) << 22);
    w[1] = (u32)f[8] | ((u32)f[13] << 8) | ((u32)f[14] << 9) | ((u32)f[15] << 10) | ((u32)f[16] << 16) | ((u32)f[17] << 17) |
           ((u32)f[18] << 18) | (((bits >> 9) & 3u) << 24) | (((bits >> 12) & 3u) << 26);
#pragma unroll
    for (int i = 0; i < 3; i++)
        w[2 + i] = (u32)f[19 + 3 * i] | ((u32)f[20 + 3 * i] << 1) | ((u32)f[21 + 3 * i] << 9) | (((bits >> (15 + 3 * i)) & 3u) << 17);
    w[2] |= ((u32)f[29] << 21) | ((u32)f[28] << 29) | (((bits >> 24) & 3u) << 30);
    w[3] |= (u32)f[30] << 21;
}
// result parts, OR-ed onto the observation parts (a step that ends an episode reports ITS rewards and flags with the
// next episode's first observation)
FJSP_HD u32 wire_s0_res(u32 flags, int ps_local10) { return (wire_flag_bits(flags) << 19) | (code_ps(ps_local10) << 27); }
FJSP_HD u32 wire_s1_res(int d_orders, int d_products) { return (u32)d_products | ((u32)d_orders << 10); }
FJSP_HD void wire_cell_res(const int* l10, u32* w) {  // l10: agv, small machine, big machine, four packaging stations
    w[0] |= code_agv(l10[0]) << 29;
    w[1] |= (code_mach(l10[1]) << 28) | (code_mach(l10[2]) << 30);
    w[2] |= code_pack(l10[3]) << 19;
    w[3] |= (code_pack(l10[4]) << 19) | (code_pack(l10[6]) << 29);
    w[4] |= code_pack(l10[5]) << 19;
}

template <int K>
FJSP_HD void wire_row(const StepOut<K>& out, u32* row) {
#pragma unroll
    for (int i = 0; i < Wire<K>::WORDS; i++) row[i] = i < Wire<K>::USED ? out.wire[i] : 0u;
    row[0] |= wire_s0_res(out.flags, out.reward_local10[0]);
    row[1] |= wire_s1_res(out.d_orders, out.d_products);
#pragma unroll
    for (int c = 0; c < K; c++) wire_cell_res(out.reward_local10 + 1 + 7 * c, row + 2 + 5 * c);
}

// the observation part of the wire row of the current state
template <int K, class S>
FJSP_HD void observe_wire(S& s, const Params& P, const Hot& h, const HotCell& c0, u32* wire) {
    {
        int fs[7];
        u32 mw = 0u;
        observe_shared(s, P, h, FieldSink{fs}, &mw);
        wire[0] = wire_s0_obs(fs, mask_nibble(mw));
        wire[1] = (u32)h.ready_count << 19;
    }
#pragma unroll
    for (int c = 0; c < K; c++) {
        HotCell hc;
        if (c > 0) load_cell<K>(s, c, hc);
        int f[31];
        u32 mw[7] = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
        observe_cell(s, P, h, c == 0 ? c0 : hc, c, FieldSink{f}, mw, 0);
        u32 bits = 0u;
#pragma unroll
        for (int i = 0; i < 7; i++) bits |= mask_nibble(mw[i]) << (4 * i);
        wire_cell_obs(f, bits, wire + 2 + 5 * c);
    }
}

// observation into the sink selected by MODE (OBS_FLOAT: out.obs + out.mask, OBS_WIRE: out.wire)
template <int K, int MODE, class S>
FJSP_HD void observe_out(S& s, const Params& P, const Hot& h, const HotCell& c0, StepOut<K>& out) {
    if (MODE == OBS_FLOAT) {
        observe<K>(s, P, h, c0, FloatSink{out.obs, P}, out.mask);
    } else if (MODE == OBS_WIRE) {
        observe_wire<K>(s, P, h, c0, out.wire);
    }
}

// ---------------------------------------------------------------------------------------------
// Packaging helpers
// ---------------------------------------------------------------------------------------------
// Move the first `g` queued products of station p into flight (grant events, PackagingAgent.py:138-141).
template <class S>
FJSP_HD void pack_grant(S& s, int pb, Hot& h, HotCell& hc, Pack& p, int g, int stamp) {
    while (g > 0 && p.q.len > 0) {
        int slot = p.q.head;
        u32 r = s.ld(pb + slot);
        int cnt = rec_count(r);
        int fslot;
        int last_idx;
        if (cnt <= g) {
            fifo_pop(s, pb, p.q);
            fslot = slot;
            last_idx = rec_first(r) + cnt - 1;
            g -= cnt, p.qcount -= cnt;
        } else {  // capacity split: the first g products start, the rest keeps waiting (R-PKG-cap-a)
            fslot = pool_alloc(hc);
            if (fslot < 0) {
                h.fault = FJSP_FAULT_POOL_EXHAUSTED;
                return;
            }
            u32 started = make_rec(rec_order(r), rec_first(r), g, 1);
            s.st(pb + slot, rec_split_rest(r, rec_first(r) + g, cnt - g));
            r = started;
            last_idx = rec_first(r) + g - 1;
            p.qcount -= g;
            g = 0;
        }
        s.st(pb + fslot, rec_with_stamp(r, stamp));
        fifo_push(s, pb, p.f, fslot);
        p.busy = 1, p.hascur = 1;
        p.cur_order = S::LONG ? (int)((s.ld(WM<S::LONG>::W_ORDER_B + rec_order(r)) >> 9) & 0xfffu) : rec_order(r), p.cur_idx = last_idx;
    }
}

// ---------------------------------------------------------------------------------------------
// Action phase
// ---------------------------------------------------------------------------------------------
// R1 pickup station (PickupStationAgent.py:146-232)
// STORE = false: a lane that only mirrors the pickup station's registers (cell-parallel step): same arithmetic, no writes
template <bool STORE = true, class S>
FJSP_HD void act_pickup(S& s, const Params& P, Hot& h, int a0, int& local10, u32& res) {
    constexpr bool LONG = S::LONG;
    using W = WM<LONG>;
    int loaded = 0, tray_done = 0, idle_orders = 0, success = 0;
    // a tray leaves for ready_trays: compact = a cut bit in the order word (the FIFO is implicit), long = one entry of the
    // ready FIFO at the tray's allocation index
    auto release = [&](bool exhausted) {
        if (STORE) {
            if (LONG) {
                s.rq_st(h.alloc_count - 1, make_rq(h.cur_order, h.prod_idx - h.cur_tray_count, h.cur_tray_count));
            } else if (!exhausted) {
                const u32 ow = s.ld(W::W_ORDER + h.cur_order);
                s.st(W::W_ORDER + h.cur_order, ow | (1u << (8 + h.prod_idx - 1)));  // cut after the last loaded product
            }
        }
        h.cur_tray_count = 0, h.ready_count++;
    };
    if (a0 == 0) {
        idle_orders = (h.num_orders - h.next_order) > 0 || h.cur_order >= 0;
        success = 1;
    } else if (a0 == 1) {
        int ok = 1;
        if (h.cur_order < 0) {
            if (h.next_order < h.num_orders) {
                h.cur_order = h.next_order++, h.prod_idx = 0;
                // long: the order's attributes come from the stream / table now (every lane that mirrors the pickup station
                // computes the same word)
                if (LONG) h.cur_word = (int)(s.fetch_order(h.cur_order, h.episode) & 0xffu);
            } else ok = 0;
        }
        if (ok && h.cur_tray_count == 0) {
            if (h.alloc_count < P.trays_total) h.alloc_count++;  // trays_at_station.pop(0)
            else ok = 0;
        }
        if (ok) {
            const int n = LONG ? (h.cur_word & 15) : ord_n(s.ld(W::W_ORDER + h.cur_order));
            h.cur_tray_count++, h.prod_idx++;
            loaded = 1, success = 1;
            if (h.prod_idx >= n) {                  // order exhausted: tray released (:201-208)
                release(true);
                h.cur_order = -1, h.prod_idx = 0;
                tray_done = 1;
            } else if (h.cur_tray_count >= FJSP_TRAY_CAPACITY) {  // tray full (:211-216)
                release(false);
                tray_done = 1;
            }
        }
    } else if (a0 == 2) {
        if (h.cur_tray_count > 0) {                 // SIGNAL (:226-230): the order is not exhausted here
            release(false);
            success = 1;
        }
    }
    res = (success ? FJSP_RES_SUCCESS : 0) | (loaded ? FJSP_RES_PS_LOADED : 0) | (tray_done ? FJSP_RES_PS_TRAY_DONE : 0) |
          (idle_orders ? FJSP_RES_PS_IDLE_ORDERS : 0);
    // RewardModel.py:53-60: +1 load, +5 tray completed, -1 idle with orders
    local10 = (loaded ? 10 : 0) + (tray_done ? 50 : 0) - ((a0 == 0 && idle_orders) ? 10 : 0);
}

// Long layout, start of a step (the lane that owns the pickup station): order slots freed by the previous step's
// completions become available now — one step late on purpose, so that allocation (an AGV's pickup, action phase) and
// release (a packaging station's finish, run phase) never meet inside a step whatever the execution order of the cells.
template <class S>
FJSP_HD void begin_step_slots(S& s) {
    if (S::LONG) {
        using W = WM<S::LONG>;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const u32 f = s.ld(W::W_SLOT_FREED + i);
            if (f) s.st(W::W_SLOT_FREE + i, s.ld(W::W_SLOT_FREE + i) | f), s.st(W::W_SLOT_FREED + i, 0u);
        }
    }
}

// FJSPSimulation.py:216-220; with order arrivals switched on an episode can only terminate once every order has arrived
template <class S>
FJSP_HD int all_orders_done(const Params& P, const Hot& h) {
    const int all = h.completed_orders == h.num_orders && h.num_orders > 0 && h.next_order == h.num_orders;
    return all && !(S::LONG && P.arrival_q16 > 0 && h.num_orders < P.arrival_max);
}

// Long order streams: one Bernoulli(arrival_q16 / 65536) ARRIVAL per step from the Philox stream (counter = global env,
// episode, step, 5), until arrival_max orders exist; order i's attributes are those of index i of the order stream.
// Runs before the pickup station acts; every lane that mirrors the shared scalars computes the same.
template <class S>
FJSP_HD void arrivals(S& s, const Params& P, Hot& h) {
    if (S::LONG && P.arrival_q16 > 0 && h.num_orders < P.arrival_max) {
        if ((s.arrival_draw(h.episode, (u32)h.step) & 0xffffu) < (u32)P.arrival_q16) h.num_orders++;
    }
}

// a tray lost to one of the reference's quirks: the compact layout keeps its record (marked) so that is_processed stays
// exportable; the long layout keeps processed bits in the ring and gives the pool slot back at once
template <class S>
FJSP_HD void lose_tray(S& s, HotCell& hc, int pb, int slot, u32 r) {
    if (S::LONG) pool_free(hc, slot);
    else s.st(pb + slot, r | REC_LOST);
}

// R2-R4: the AGV whose fields are in `hc` acts on the stations of `hc` (AGVAgent.py:180-368).  SHARED (shared floor,
// include/fjsp_b200.h): `occ` = station positions taken by OTHER AGVs (bit per location); a move there is invalid.
// Otherwise the scaled shop's one-dock rule applies (h.dock_mask).
template <bool SHARED, class S>
FJSP_HD void act_agv(S& s, const Params& P, Hot& h, HotCell& hc, int c, int k, int act, int& local10_out, u32& res_out, u32 occ) {
    constexpr bool LONG = S::LONG;
    using W = WM<LONG>;
    const int pb = pool_base<LONG>(c);
    auto lose = [&](int slot, u32 r) { lose_tray(s, hc, pb, slot, r); };
    // ---- R2-R4 AGV (AGVAgent.py:180-368)
    {
        int invalid = 0, moved = 0, pick = 0, drop = 0, to_pack = 0, success = 0;
        if (hc.agv_moving) {
            invalid = 1;
        } else if (act == 0) {
            success = 1;
        } else if (act <= 5) {
            const int tl = act == 1 ? LOC_PICKUP : act == 2 ? LOC_SMALL : act == 3 ? LOC_BIG : act == 4 ? LOC_STORAGE : LOC_PACKAGING;
            const int d = P.dist[hc.agv_loc][tl];
            if (SHARED ? (d != 0 && ((occ >> tl) & 1u)) : (tl == LOC_PICKUP && d != 0 && (h.dock_mask & ~(1 << c)) != 0)) {
                invalid = 1;  // shared floor: the position is another AGV's; scaled shop: the single dock is taken
            } else {
                success = 1;
                if (d != 0) {
                    if (!SHARED && tl == LOC_PICKUP) h.dock_mask |= 1 << c;  // granted now: later AGVs of this step already see it
                    moved = 1;
                    hc.agv_moving = 1, hc.agv_target = tl;        // resolved in the run phase
                    hc.agv_arrive = k + P.delay[hc.agv_loc][tl];
                }
            }
        } else if (act == 6) {
            const int loc = hc.agv_loc;
            if (hc.carry != 0 || loc == LOC_PACKAGING) {
                invalid = 1;
            } else if (loc == LOC_PICKUP) {
                if (h.ready_count > 0) {
                    int slot = pool_alloc(hc);
                    if (slot < 0) {
                        h.fault = FJSP_FAULT_POOL_EXHAUSTED, invalid = 1;
                    } else {
                        if (LONG) {
                            // head of the ready FIFO; the tray's order enters an ORDER SLOT with its first tray
                            const u32 e = s.rq_ld(h.ready_order);
                            const int o = rq_order(e), first = rq_first(e), cnt = rq_count(e);
                            int os = h.act_slot;
                            if (o != h.act_order) {
                                const u32 f0 = s.ld(W::W_SLOT_FREE), f1 = s.ld(W::W_SLOT_FREE + 1);
                                os = f0 ? ctz32(f0) : f1 ? 32 + ctz32(f1) : -1;
                                if (os >= 0) {
                                    s.st(W::W_SLOT_FREE + (os >> 5), (os < 32 ? f0 : f1) & ~(1u << (os & 31)));
                                    s.st(W::W_ORDER + os, s.fetch_order(o, h.episode) & 0xffu);
                                    s.st(W::W_ORDER_B + os, (u32)o << 9);
                                    s.st(W::W_ORDER_C + os, (u32)h.ready_order << 16);
                                    h.act_order = o, h.act_slot = os;
                                }
                            }
                            if (os < 0) {
                                h.fault = FJSP_FAULT_ORDER_SLOTS, invalid = 1;
                                pool_free(hc, slot);
                            } else {
                                const u32 ow = s.ld(W::W_ORDER + os);
                                if (first + cnt < ord_n(ow)) s.st(W::W_ORDER + os, ow | (1u << (8 + first + cnt - 1)));  // (tray ids of the export)
                                s.st(pb + slot, make_rec(os, first, cnt, 0));
                                h.ready_order++, h.ready_count--;
                                hc.carry = slot + 1, pick = 1;
                            }
                        } else {
                        // head ready tray = products [ready_idx, next cut] of order ready_order
                        u32 ow = s.ld(W::W_ORDER + h.ready_order);
                        u32 cuts = ord_cut(ow) >> h.ready_idx;
                        int cnt = cuts ? ctz32(cuts) + 1 : ord_n(ow) - h.ready_idx;
                        s.st(pb + slot, make_rec(h.ready_order, h.ready_idx, cnt, 0));
                        h.ready_idx += cnt;
                        if (h.ready_idx >= ord_n(ow)) h.ready_order++, h.ready_idx = 0;
                        h.ready_count--;
                        hc.carry = slot + 1, pick = 1;
                        }
                    }
                } else invalid = 1;
            } else {
                // (three static branches, not a reference picked at run time: that would put the FIFOs in local memory)
                int got = -1;
                if (loc == LOC_SMALL) {
                    if (hc.m[0].r.len > 0) got = fifo_pop(s, pb, hc.m[0].r);
                } else if (loc == LOC_BIG) {
                    if (hc.m[1].r.len > 0) got = fifo_pop(s, pb, hc.m[1].r);
                } else {
                    if (hc.storage.len > 0) got = fifo_pop(s, pb, hc.storage);
                }
                if (got >= 0) hc.carry = got + 1, pick = 1;
                else invalid = 1;
            }
            success = pick;
        } else if (act == 7) {
            const int loc = hc.agv_loc;
            if (hc.carry == 0 || loc == LOC_PICKUP) {
                invalid = 1;  // empty-handed, or a non-empty tray at PICKUP (:310-317)
            } else {
                const int slot = hc.carry - 1;
                u32 r = s.ld(pb + slot);
                u32 ow = s.ld(W::W_ORDER + rec_order(r));
                const int ty = ord_type(ow), proc = rec_processed(r);
                if (loc == LOC_SMALL || loc == LOC_BIG) {
                    int compat = loc == LOC_SMALL ? (ty == TYPE_SMALL || ty == TYPE_MEDIUM) : (ty == TYPE_BIG || ty == TYPE_MEDIUM);
                    if (!proc && compat) {
                        if (loc == LOC_SMALL) fifo_push(s, pb, hc.m[0].q, slot);
                        else fifo_push(s, pb, hc.m[1].q, slot);
                        drop = 1;
                    } else invalid = 1;
                } else if (loc == LOC_STORAGE) {
                    if (hc.storage.len < P.storage_capacity) fifo_push(s, pb, hc.storage, slot);
                    else lose(slot, r);  // Storage.add_tray False ignored: tray vanishes (Storage.py:18-22)
                    drop = 1;
                } else {  // PACKAGING (:352-360) -> add_tray_to_packaging (FJSPSimulation.py:402-430)
                    if (proc) {
                        const int col = ord_colour(ow), cnt = rec_count(r);
                        int st = -1;
                        if (col == COL_BLUE) st = hc.p[0].users < P.pack_capacity ? 0 : (hc.p[1].users < P.pack_capacity ? 1 : -1);
                        else if (col == COL_RED) st = hc.p[2].users < P.pack_capacity ? 2 : -1;
                        else st = hc.p[3].users < P.pack_capacity ? 3 : -1;
                        if (st == 0) fifo_push(s, pb, hc.p[0].q, slot), hc.p[0].qcount += cnt;
                        else if (st == 1) fifo_push(s, pb, hc.p[1].q, slot), hc.p[1].qcount += cnt;
                        else if (st == 2) fifo_push(s, pb, hc.p[2].q, slot), hc.p[2].qcount += cnt;
                        else if (st == 3) fifo_push(s, pb, hc.p[3].q, slot), hc.p[3].qcount += cnt;
                        else lose(slot, r);  // no station with capacity: products dropped (:426-427)
                        drop = 1, to_pack = 1;
                    } else invalid = 1;
                }
                if (drop) hc.carry = 0, success = 1;
            }
        } else {
            invalid = 1;
        }
        res_out = (success ? FJSP_RES_SUCCESS : 0) | (invalid ? FJSP_RES_AGV_INVALID : 0) | (moved ? FJSP_RES_AGV_MOVED : 0) |
                 (pick ? FJSP_RES_AGV_PICKUP : 0) | (drop ? FJSP_RES_AGV_DROP : 0) | (to_pack ? FJSP_RES_AGV_TO_PACK : 0);
        // RewardModel.py:62-77: +2 pickup, +2 drop, +10 delivered to packaging, -0.1 move, -5 invalid
        local10_out = (pick ? 20 : 0) + (drop ? 20 : 0) + (to_pack ? 100 : 0) - (moved ? 1 : 0) - (invalid ? 50 : 0);
    }
}

// one cell's seven agents: agv, small machine, big machine, four packaging stations.  a/local10/res point at the cell's
// first column.  pk_start[i] = products whose packaging processes START creates in this step (resolved in run_cell).
template <class S>
FJSP_HD void act_stations(S& s, const Params& P, Hot& h, HotCell& hc, int c, int k, const int* a, int* local10, u32* res, int* pk_start);
template <class S>
FJSP_HD void act_cell(S& s, const Params& P, Hot& h, HotCell& hc, int c, int k, const int* a, int* local10, u32* res, int* pk_start) {
    act_agv<false>(s, P, h, hc, c, k, a[0], local10[0], res[0], 0u);
    act_stations(s, P, h, hc, c, k, a + 1, local10 + 1, res + 1, pk_start);
}
// the six station agents of a cell: a/local10/res point at the small machine's column
template <class S>
FJSP_HD void act_stations(S& s, const Params& P, Hot& h, HotCell& hc, int c, int k, const int* a, int* local10, u32* res, int* pk_start) {
    constexpr bool LONG = S::LONG;
    const int pb = pool_base<LONG>(c);
    auto lose = [&](int slot, u32 r) { lose_tray(s, hc, pb, slot, r); };
    // ---- R5 machines (MachineAgent.py:99-169).  START takes effect in this step's run phase, but no later
    //      agent reads machine state inside the action phase, so it is applied here.
#pragma unroll
    for (int i = 0; i < 2; i++) {
        Mach& m = hc.m[i];
        const int act = a[i];
        int started = 0, completed = 0, idle_q = 0, success = 0;
        if (act == 0) {
            idle_q = m.q.len > 0 && !m.busy;
            success = 1;
        } else if (act == 1) {
            if (m.q.len > 0 && !m.busy) {
                int slot = fifo_pop(s, pb, m.q);
                if (m.has_cur) lose(m.cur, s.ld(pb + m.cur));  // unsignalled finished tray is overwritten and lost (:160)
                m.busy = 1, m.has_cur = 1, m.cur = slot, m.start = k;
                started = 1, success = 1;
            }
        } else if (act == 2) {
            if (!m.busy && m.has_cur) {
                fifo_push(s, pb, m.r, m.cur);
                m.has_cur = 0, m.cur = 0;
                completed = 1, success = 1;
            }
        }
        res[i] = (success ? FJSP_RES_SUCCESS : 0) | (started ? FJSP_RES_M_STARTED : 0) | (completed ? FJSP_RES_M_COMPLETED : 0) |
                     (idle_q ? FJSP_RES_M_IDLE_QUEUE : 0);
        // RewardModel.py:79-86: +1 start, +5 signal complete, -2 idle with queue
        local10[i] = (started ? 10 : 0) + (completed ? 50 : 0) - ((act == 0 && idle_q) ? 20 : 0);
    }
    // ---- R6 packaging (PackagingAgent.py:91-125)
#pragma unroll
    for (int i = 0; i < 4; i++) {
        Pack& p = hc.p[i];
        const int act = a[2 + i];
        int started = 0, completed = 0, idle_q = 0, success = 0;
        pk_start[i] = 0;
        if (act == 0) {
            idle_q = p.qcount > 0 && !p.busy;
            success = 1;
        } else if (act == 1) {
            if (p.qcount > 0) {  // no busy/capacity check (:111-118)
                started = 1, success = 1;
                p.progL = p.qcount;
                if (p.waiters) h.fault = FJSP_FAULT_PKG_RESTART_WITH_WAITERS;  // reference raises (R-PKG-cap-b)
                else pk_start[i] = p.qcount;
            }
        } else if (act == 2) {
            if (!p.busy && p.hascur) completed = 1;  // success stays False (:120-123)
        }
        res[2 + i] = (success ? FJSP_RES_SUCCESS : 0) | (started ? FJSP_RES_M_STARTED : 0) | (completed ? FJSP_RES_M_COMPLETED : 0) |
                     (idle_q ? FJSP_RES_M_IDLE_QUEUE : 0);
        // RewardModel.py:88-95: +2 start, +20 signal complete, -1 idle with queue
        local10[2 + i] = (started ? 20 : 0) + (completed ? 200 : 0) - ((act == 0 && idle_q) ? 10 : 0);
    }
}

// ---------------------------------------------------------------------------------------------
// Run phase of one cell: env.run(until = now + step_size) (FJSPSimulation.py:183-184), rule R0
// ---------------------------------------------------------------------------------------------
// `dock_after` collects who holds the dock once the run phase is over; h.dock_mask itself keeps its action-phase meaning
// until every cell has acted (cells are processed one after the other here, but all actions precede all runs).
// AGV arrival (AGVAgent.py:387-396)
FJSP_HD void run_agv(HotCell& hc, int k) {
    if (hc.agv_moving && hc.agv_arrive == (u32)(k & 0xffff)) hc.agv_loc = hc.agv_target, hc.agv_moving = 0;
}
template <class S>
FJSP_HD void run_cell(S& s, const Params& P, Hot& h, HotCell& hc, int c, int k, const int* pk_start, int& dock_after) {
    constexpr bool LONG = S::LONG;
    using W = WM<LONG>;
    const int pb = pool_base<LONG>(c);
    run_agv(hc, k);
    // the dock is held while standing at PICKUP or under way to it
    dock_after |= (hc.agv_moving ? (hc.agv_target == LOC_PICKUP) : (hc.agv_loc == LOC_PICKUP)) << c;
    // machines: product i (1-based) is flagged in step start + P*i; the tray finishes at start + P*n
#pragma unroll
    for (int i = 0; i < 2; i++) {
        Mach& m = hc.m[i];
        if (m.busy) {
            u32 r = s.ld(pb + m.cur);
            const int per = i == 0 ? P.small_steps : P.big_steps;
            if (k == m.start + per * rec_count(r)) {
                s.st(pb + m.cur, r | REC_PROCESSED);
                if (LONG) s.or_word(W::W_ORDER_B + rec_order(r), ((1u << rec_count(r)) - 1u) << rec_first(r));  // is_processed, kept in the ring
                m.busy = 0, m.prog = 1;
            }
        }
    }
    // packaging stations, event order inside the run:
    //   URGENT Initialize of this step's requests (grant while users < capacity, users counted BEFORE finishes)
    //   -> NORMAL timeouts due at the boundary (finish, release) -> grants -> releases grant waiters FIFO.
#pragma unroll
    for (int i = 0; i < 4; i++) {
        Pack& p = hc.p[i];
        int g = 0;
        if (pk_start[i] > 0) {
            int room = P.pack_capacity - p.users;
            g = pk_start[i] < room ? pk_start[i] : room;
            if (g < pk_start[i]) p.waiters = 1;
            p.users += g;
        }
        int released = 0;
        while (p.f.len > 0) {
            int slot = p.f.head;
            u32 r = s.ld(pb + slot);
            if (rec_stamp(r) != (k & 255)) break;
            fifo_pop(s, pb, p.f);
            pool_free(hc, slot);
            const int o = rec_order(r), cnt = rec_count(r);
            // or_word: a plain read-modify-write when one thread owns the env, an atomic OR when the cells of an env run
            // on different threads (two cells may package products of the same order in the same step)
            const u32 bits = (((1u << cnt) - 1u) << rec_first(r)) << 16;   // is_packaged (:143)
            const u32 before = s.or_word(W::W_ORDER + o, bits);
            const u32 ow = before | bits;
            p.completed += cnt, h.total_packaged += cnt, p.users -= cnt, released += cnt;
            p.busy = 0;
            const u32 full = (1u << ord_n(ow)) - 1u;
            if (ord_packaged(ow) == full && ord_packaged(before) != full) {  // _check_order_completions (FJSPSimulation.py:245-258)
                h.completed_orders++;
                if (LONG) {
                    s.or_word(W::W_ORDER_C + o, (u32)((k + 1) & 0xffff));
                    s.or_word(W::W_SLOT_FREED + (o >> 5), 1u << (o & 31));  // the slot is free again from the next step on
                } else {
                    s.or_word(W::W_CSTEP + (o >> 2), (u32)((k + 1) & 255) << ((o & 3) * 8));
                }
            }
        }
        const int stamp = (k + P.pack_steps) & 255;
        if (g > 0) pack_grant(s, pb, h, hc, p, g, stamp);
        if (p.waiters && released > 0) {
            int wgrant = released < p.qcount ? released : p.qcount;
            int room = P.pack_capacity - p.users;
            if (wgrant > room) wgrant = room;
            p.users += wgrant;
            pack_grant(s, pb, h, hc, p, wgrant, stamp);
            if (p.qcount == 0) p.waiters = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One environment step of a K-cell shop.  `h` = shared hot words and `c0` = cell 0's hot words live with the caller
// (registers: loaded/stored once per step in the step kernel, once per LAUNCH in the K-steps-per-launch kernel);
// cells 1..K-1 are loaded and stored here, one at a time.  Agent order: pickup station, then cell by cell agv, small
// machine, big machine, four packaging stations (FJSPSimulation.py:76-82,172-174 for K = 1).
// ---------------------------------------------------------------------------------------------
template <int K, int MODE, class S>
FJSP_HD void step_env_hot(S& s, const Params& P, Hot& h, HotCell& c0, const int* a, StepOut<K>& out) {
    constexpr int A = Lay<K>::AGENTS;
    const int k = h.step;
    const int orders_before = h.completed_orders, products_before = h.total_packaged;
    int local10[Lay<K>::ACT];  // local rewards in tenths: every RewardModel constant is a multiple of 0.1 (RewardModel.py:12-32)
    u32 res[Lay<K>::ACT];
#pragma unroll
    for (int i = A; i < Lay<K>::ACT; i++) local10[i] = 0, res[i] = 0u;

    if (k > P.max_episode_steps) {
        // The episode ended with the previous step (truncation fires when step == max_episode_steps) and the caller did not
        // reset: the reference would simulate on, the packed counters are not sized for that -> the env is inert and says so.
        out.d_orders = 0, out.d_products = 0, out.reward_g = 0, out.reward_units = 0;
#pragma unroll
        for (int i = 0; i < Lay<K>::ACT; i++) out.reward_local10[i] = 0, out.reward[i] = 0.0f;
#pragma unroll
        for (int i = 0; i < Lay<K>::ACT / 4; i++) out.results[i] = 0u;
        out.flags = (1u << 8) | ((u32)FJSP_FAULT_PAST_END << 16);
        out.info[0] = h.step, out.info[1] = h.completed_orders, out.info[2] = h.total_packaged, out.info[3] = 0;
        observe_out<K, MODE>(s, P, h, c0, out);
        return;
    }
    arrivals(s, P, h);
    begin_step_slots(s);
    act_pickup(s, P, h, a[0], local10[0], res[0]);
    int dock_after = 0;
    {
        int pk_start[4];
        act_cell(s, P, h, c0, 0, k, a + 1, local10 + 1, res + 1, pk_start);
        run_cell(s, P, h, c0, 0, k, pk_start, dock_after);
    }
#pragma unroll
    for (int c = 1; c < K; c++) {
        HotCell hc;
        int pk_start[4];
        load_cell<K>(s, c, hc);
        act_cell(s, P, h, hc, c, k, a + 1 + 7 * c, local10 + 1 + 7 * c, res + 1 + 7 * c, pk_start);
        run_cell(s, P, h, hc, c, k, pk_start, dock_after);
        store_cell<K>(s, c, hc);
    }
    h.dock_mask = dock_after;

    // ===== rewards (FJSPSimulation.py:190-209, RewardModel.py:34-44,99-110) =====
    // r_i = g/A + local_i with g = 100*orders + 10*products - 0.1*step_size and A = 1 + 7K agents is an exact multiple of
    // 1/(10A):  10A*r_i = 10*(100*orders + 10*products) - step_size + A*(10*local_i)   (A = 8: multiples of 1/80).
    // One correctly rounded fp32 division of that integer reproduces the fp32 rounding of the reference's float64 value
    // (no FP64 in the kernel; a non-dyadic k/(10A) is never within double-rounding distance of an fp32 midpoint).
    {
        out.d_orders = h.completed_orders - orders_before, out.d_products = h.total_packaged - products_before;
        const int g = 10 * (100 * out.d_orders + 10 * out.d_products) - P.step_size;
        long long units = 0;
        out.reward_g = g;
#pragma unroll
        for (int i = 0; i < Lay<K>::ACT; i++) {
            out.reward_local10[i] = local10[i];
            if (i < A) {
                const int n = g + A * local10[i];
                out.reward[i] = (float)n / (float)(10 * A);
                units += n;
            } else {
                out.reward[i] = 0.0f;
            }
        }
        out.reward_units = units;
    }
    // ===== termination / truncation (FJSPSimulation.py:216-224), pre-increment step =====
    const int all_done = all_orders_done<S>(P, h);
    const int truncated = k >= P.max_episode_steps;
    out.flags = (u32)all_done | ((u32)truncated << 8) | ((u32)h.fault << 16);
#pragma unroll
    for (int i = 0; i < Lay<K>::ACT / 4; i++)
        out.results[i] = res[4 * i] | (res[4 * i + 1] << 8) | (res[4 * i + 2] << 16) | (res[4 * i + 3] << 24);
    h.step = k + 1;
    out.info[0] = h.step, out.info[1] = h.completed_orders, out.info[2] = h.total_packaged, out.info[3] = 0;
    observe_out<K, MODE>(s, P, h, c0, out);
}

template <int K, int MODE, class S>
FJSP_HD void step_env(S& s, const Params& P, const int* a, StepOut<K>& out) {
    Hot h;
    HotCell c0;
    load_hot(s, h);
    load_cell<K>(s, 0, c0);
    step_env_hot<K, MODE>(s, P, h, c0, a, out);
    store_hot(s, h);
    store_cell<K>(s, 0, c0);
}


// ---------------------------------------------------------------------------------------------
// Cell-parallel step (K >= 2): ONE THREAD PER (env, cell).  Same results as step_env_hot; the cells of an env meet at
// three points only, through a small per-env exchange area X (a shared-memory column on the device):
//   1. the pickup station acts first (every lane mirrors its registers, the lane of cell 0 writes the order words) and
//      every AGV posts whether it asks for the single dock                                                  -> barrier
//      (long layout: it also makes last step's freed order slots available and writes released trays to the ready FIFO)
//   2. each lane resolves the dock in agent order from the posted requests (an earlier AGV's grant blocks a later one),
//      runs its cell's seven agents and its run phase; order words are updated with atomic ORs; it posts the ready-tray
//      cursor if its AGV took a tray at the pickup station (only the dock holder can), its packaged / completed
//      counts, its dock occupancy and a fault code                                                          -> barrier
//   3. every lane rebuilds the same shared scalars from X, computes its agents' rewards and observation, posts the
//      reward numerators and mask bits                                                                      -> barrier
//   4. output rows are assembled in 32-byte pieces, one per lane.
// X slots: X_RDF = dock requests (bits 0..3) | dock occupancy after the step (4..7) | fault code of cell c (8+3c..10+3c) |
// fault code of the pickup station (20..22);
// X_READY = ready cursor posted by the picking lane (bit 31 = valid); X_DELTA = completed orders << 16 | packaged
// products; X_MASK.. = 3 mask bits of the pickup station, then 26 per cell; then one u16 per action column:
// local reward (9-bit signed, tenths) | action_result << 9.
// ---------------------------------------------------------------------------------------------
enum { X_RDF = 0, X_READY = 1, X_DELTA = 2, X_MASK = 3 };
template <int K, bool LONG = false>
struct Xl {
    static constexpr int LOCAL = X_MASK + 1 + K;  // first word of the u16 table
    static constexpr int ACT = LOCAL + Lay<K>::ACT / 2;  // long layout only: the order slot taken by the picking lane
    static constexpr int WORDS = ACT + (LONG ? 1 : 0);
};

struct CellLane {
    Hot h;
    HotCell hc;
    int c, k, fault_in, orders_in, packaged_in;
    int local10[8];  // [0] pickup station (meaningful on the lane of cell 0), [1..7] the cell's agents
    u32 res[8];
    int g;           // after cells_finish: 10 * (100 * orders + 10 * products) - step_size
    int d_orders, d_products;
    u32 flags;
    bool inert;      // stepped past the end of the episode (see step_env_hot)
};

template <int K, class S, class X>
FJSP_HD void cells_begin(S& s, X& x, const Params& P, CellLane& L, int a0, const int* a7) {
    L.k = L.h.step, L.fault_in = L.h.fault;
    L.inert = L.k > P.max_episode_steps;
    if (L.inert) return;
    arrivals(s, P, L.h);
    if (L.c == 0) {
        begin_step_slots(s);
        act_pickup<true>(s, P, L.h, a0, L.local10[0], L.res[0]);
        if (L.h.fault != L.fault_in) x.atom_or(X_RDF, (u32)(L.h.fault & 7) << 20);  // (order ring overflow: long layout)
    } else {
        act_pickup<false>(s, P, L.h, a0, L.local10[0], L.res[0]);
    }
    const int wants_dock = !L.hc.agv_moving && a7[0] == 1 && P.dist[L.hc.agv_loc][LOC_PICKUP] != 0;
    if (wants_dock) x.atom_or(X_RDF, 1u << L.c);
}

template <int K, class S, class X>
FJSP_HD void cells_act_run(S& s, X& x, const Params& P, CellLane& L, const int* a7) {
    if (L.inert) return;
    // the dock as this cell's AGV sees it: holders before the step + grants to earlier cells of this step
    const u32 reqs = x.ld(X_RDF) & 15u;
    int m = L.h.dock_mask;
#pragma unroll
    for (int j = 0; j < K - 1; j++)
        if (j < L.c && ((reqs >> j) & 1u) && (m & ~(1 << j)) == 0) m |= 1 << j;
    L.h.dock_mask = m;
    const int rc = L.h.ready_count, ro = L.h.ready_order, ri = L.h.ready_idx;
    L.orders_in = L.h.completed_orders, L.packaged_in = L.h.total_packaged;
    L.h.completed_orders = 0, L.h.total_packaged = 0, L.h.fault = -1;
    int pk_start[4];
    act_cell(s, P, L.h, L.hc, L.c, L.k, a7, L.local10 + 1, L.res + 1, pk_start);
    if (L.h.ready_count != rc || L.h.ready_order != ro || L.h.ready_idx != ri) {
        x.st(X_READY, (u32)L.h.ready_count | ((u32)L.h.ready_order << 12) | ((u32)L.h.ready_idx << 24) | (1u << 31));
        if (S::LONG) x.st(Xl<K, S::LONG>::ACT, (u32)(L.h.act_order & 0x1fff) | ((u32)L.h.act_slot << 13) | (1u << 31));
    }
    int dock_after = 0;
    run_cell(s, P, L.h, L.hc, L.c, L.k, pk_start, dock_after);
    u32 post = (u32)dock_after << 4;
    if (L.h.fault != -1) post |= (u32)(L.h.fault & 7) << (8 + 3 * L.c);
    if (post) x.atom_or(X_RDF, post);
    const u32 delta = ((u32)L.h.completed_orders << 16) | (u32)L.h.total_packaged;
    if (delta) x.atom_add(X_DELTA, delta);
}

template <int K, class S, class X>
FJSP_HD void cells_finish(X& x, const Params& P, CellLane& L, int32_t* info) {
    constexpr int A = Lay<K>::AGENTS;
    Hot& h = L.h;
    if (L.inert) {
        L.g = 0, L.d_orders = 0, L.d_products = 0;
        L.flags = (1u << 8) | ((u32)FJSP_FAULT_PAST_END << 16);
#pragma unroll
        for (int i = 0; i < 8; i++) L.local10[i] = 0, L.res[i] = 0u;
        L.orders_in = h.completed_orders, L.packaged_in = h.total_packaged;
        info[0] = h.step, info[1] = h.completed_orders, info[2] = h.total_packaged, info[3] = 0;
    }
    const u32 rdf = x.ld(X_RDF), ready = x.ld(X_READY), delta = x.ld(X_DELTA);
    if (!L.inert) {
    if (ready >> 31) h.ready_count = (int)(ready & 0xfffu), h.ready_order = (int)((ready >> 12) & 0xfffu), h.ready_idx = (int)((ready >> 24) & 15u);
    if (S::LONG) {
        const u32 act = x.ld(Xl<K, S::LONG>::ACT);
        if (act >> 31) h.act_order = (int)(act & 0x1fffu), h.act_slot = (int)((act >> 13) & 63u);
    }
    const int d_orders = (int)(delta >> 16), d_products = (int)(delta & 0xffffu);
    h.completed_orders = L.orders_in + d_orders, h.total_packaged = L.packaged_in + d_products;
    h.dock_mask = (int)((rdf >> 4) & 15u);
    h.fault = L.fault_in;
    if ((rdf >> 20) & 7u) h.fault = (int)((rdf >> 20) & 7u);
#pragma unroll
    for (int j = 0; j < K; j++) {  // the last agent that raised a fault wins, as in agent order
        const int f = (int)((rdf >> (8 + 3 * j)) & 7u);
        if (f) h.fault = f;
    }
    L.g = 10 * (100 * d_orders + 10 * d_products) - P.step_size;
    L.d_orders = d_orders, L.d_products = d_products;
    const int all_done = all_orders_done<S>(P, h);
    const int truncated = L.k >= P.max_episode_steps;
    L.flags = (u32)all_done | ((u32)truncated << 8) | ((u32)h.fault << 16);
    h.step = L.k + 1;
    info[0] = h.step, info[1] = h.completed_orders, info[2] = h.total_packaged, info[3] = 0;
    }
    // post this lane's columns: local reward (tenths, 9-bit signed) | action_result << 9
    if (L.c == 0) x.st16(2 * Xl<K>::LOCAL, ((u32)L.local10[0] & 0x1ffu) | (L.res[0] << 9));
#pragma unroll
    for (int i = 1; i < 8; i++) x.st16(2 * Xl<K>::LOCAL + 7 * L.c + i, ((u32)L.local10[i] & 0x1ffu) | (L.res[i] << 9));
    if (L.c == K - 1) {
#pragma unroll
        for (int i = A; i < Lay<K>::ACT; i++) x.st16(2 * Xl<K>::LOCAL + i, 0u);
    }
}

// this lane's part of the observation: its cell (31 fields into `obs_cell`), plus the pickup station on the lane of
// cell 0 (`obs_shared`); the mask bits are posted to X.
template <int K, class S, class X, class O>
FJSP_HD void cells_observe(S& s, X& x, const Params& P, const CellLane& L, O obs_shared, O obs_cell) {
    u32 mw[8];
    if (L.c == 0) {
        mw[0] = 0u;
        observe_shared(s, P, L.h, obs_shared, mw);
        x.st(X_MASK, mask_nibble(mw[0]));
    }
#pragma unroll
    for (int i = 0; i < 8; i++) mw[i] = 0u;
    observe_cell(s, P, L.h, L.hc, L.c, obs_cell, mw, 0);
    u32 bits = 0u;
#pragma unroll
    for (int i = 0; i < 7; i++) bits |= mask_nibble(mw[i]) << (4 * i);
    x.st(X_MASK + 1 + L.c, bits & 0x3ffffffu);
}

// the lane's part of the env's WIRE row: its cell's five words (observation + its seven agents' reward codes), plus the
// two shared words on the lane of cell 0.  Nothing is posted: every bit of a word comes from the lane that writes it.
template <int K, class S>
FJSP_HD void cells_observe_wire(S& s, const Params& P, const CellLane& L, u32* shared2, u32* cell5) {
    if (L.c == 0) {
        int fs[7];
        u32 mw = 0u;
        observe_shared(s, P, L.h, FieldSink{fs}, &mw);
        shared2[0] = wire_s0_obs(fs, mask_nibble(mw)) | wire_s0_res(L.flags, L.local10[0]);
        shared2[1] = ((u32)L.h.ready_count << 19) | wire_s1_res(L.d_orders, L.d_products);
    }
    int f[31];
    u32 mw[7] = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
    observe_cell(s, P, L.h, L.hc, L.c, FieldSink{f}, mw, 0);
    u32 bits = 0u;
#pragma unroll
    for (int i = 0; i < 7; i++) bits |= mask_nibble(mw[i]) << (4 * i);
    wire_cell_obs(f, bits, cell5);
    wire_cell_res(L.local10 + 1, cell5);
}

// mask bits [32 * word, 32 * word + 32) of the env's mask row, from the posted pieces
template <int K, class X>
FJSP_HD u32 cells_mask_word(X& x, int word) {
    uint64_t acc = (word == 0) ? (uint64_t)(x.ld(X_MASK) & 7u) : 0ull;
#pragma unroll
    for (int j = 0; j < K; j++) {
        const int rel = 3 + 26 * j - 32 * word;  // cell j's 26 bits land at [rel, rel + 26) of this word
        if (rel > -26 && rel < 32) {
            const uint64_t b = (uint64_t)x.ld(X_MASK + 1 + j);
            acc |= rel >= 0 ? (b << rel) : (b >> (-rel));
        }
    }
    return (u32)acc;
}
FJSP_HD int x_local10(u32 v16) { return (int)((v16 & 0x1ffu) ^ 0x100u) - 0x100; }  // 9-bit sign extension
FJSP_HD u32 x_result(u32 v16) { return (v16 >> 9) & 0x7fu; }
// four bits -> four mask bytes (inverse of mask_nibble)
FJSP_HD u32 nibble_bytes(u32 n) { return ((n & 15u) * 0x00204081u) & 0x01010101u; }

// observation of the current state without stepping (reset path)
template <int K, class S>
FJSP_HD void observe_env(S& s, const Params& P, float* obs, u32* mw) {
    Hot h;
    HotCell c0;
    load_hot(s, h);
    load_cell<K>(s, 0, c0);
    observe<K>(s, P, h, c0, FloatSink{obs, P}, mw);
}

}  // namespace fjsp
#endif  // FJSP_CORE_H
