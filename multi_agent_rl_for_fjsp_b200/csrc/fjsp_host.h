// fjsp_host.h — host-side helpers shared by the C-ABI library and the test harness:
//   * FjspConfig validation + derivation of the kernel Params,
//   * decode of one env's packed words into the canonical record S (include/fjsp_b200.h).
// No simulation logic lives here.
#ifndef FJSP_HOST_H
#define FJSP_HOST_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "fjsp_core.h"

namespace fjsp {

inline void default_config(FjspConfig* c) {
    memset(c, 0, sizeof(*c));
    c->struct_size = (int32_t)sizeof(FjspConfig);
    // constants.py:5-11
    c->pos[LOC_PICKUP][0] = 0, c->pos[LOC_PICKUP][1] = 0;
    c->pos[LOC_BIG][0] = 0, c->pos[LOC_BIG][1] = 3;
    c->pos[LOC_SMALL][0] = 2, c->pos[LOC_SMALL][1] = 3;
    c->pos[LOC_STORAGE][0] = 3, c->pos[LOC_STORAGE][1] = 0;
    c->pos[LOC_PACKAGING][0] = 3, c->pos[LOC_PACKAGING][1] = 5;
    c->grid_rows = 4, c->grid_cols = 6;                        // constants.py:26-27
    c->proc_small = 60, c->proc_big = 120, c->proc_pack = 30;  // constants.py:14-18
    c->step_size = 10, c->agv_speed = 1, c->max_episode_steps = 200;
    c->storage_capacity = 100, c->pack_capacity = 20, c->tray_capacity = 5, c->num_trays = 1000;
    c->num_cells = 1;
}

// Returns nullptr if ok, else a static message.
inline const char* make_params(const FjspConfig& c, Params* P) {
    static thread_local char msg[256];
    if (c.struct_size != (int32_t)sizeof(FjspConfig)) return "FjspConfig.struct_size mismatch (ABI)";
    if (c.step_size <= 0 || c.agv_speed <= 0) return "step_size and agv_speed must be positive";
    const int32_t procs[3] = {c.proc_small, c.proc_big, c.proc_pack};
    for (int i = 0; i < 3; i++)
        if (procs[i] < c.step_size || procs[i] % c.step_size != 0)
            return "processing times must be positive multiples of step_size (timers are resolved on step boundaries)";
    if (c.proc_pack / c.step_size > 200) return "proc_pack too long";
    if (c.tray_capacity != FJSP_TRAY_CAPACITY) return "tray_capacity must be 5 (the reference's mask hard-codes CONFIG['tray_capacity'])";
    if (c.long_streams != 0 && c.long_streams != 1) return "long_streams must be 0 or 1";
    if (!c.long_streams && (c.max_episode_steps < 1 || c.max_episode_steps > 240))
        return "max_episode_steps must be in 1..240 in the compact layout (bounds the 64-slot tray pool and the 6/8-bit counters); set long_streams = 1 for longer episodes";
    if (c.long_streams && (c.max_episode_steps < 1 || c.max_episode_steps > FJSP_LONG_MAX_STEPS)) return "max_episode_steps must be in 1..65000";
    if (!c.long_streams && (c.arrival_prob_q16 || c.arrival_max_orders)) return "order arrivals need long_streams = 1";
    if (c.arrival_prob_q16 < 0 || c.arrival_prob_q16 > 65536) return "arrival_prob_q16 must be in 0..65536";
    if (c.arrival_max_orders < 0 || c.arrival_max_orders > FJSP_LONG_MAX_ORDERS) return "arrival_max_orders must be in 0..4095";
    if (c.long_streams && c.num_trays > 4095) return "num_trays must be <= 4095 with long_streams (12-bit tray ids in the canonical record)";
    if (c.pack_capacity < 1 || c.pack_capacity > 31) return "pack_capacity must be in 1..31";
    if (c.storage_capacity < 0) return "storage_capacity must be >= 0";
    if (c.num_trays < 0 || c.num_trays > 65535) return "num_trays must be in 0..65535";
    if (c.num_cells < 1 || c.num_cells > FJSP_MAX_CELLS) return "num_cells must be in 1..4";
    if (c.shared_agvs < 0 || c.shared_agvs > FJSP_MAX_SHARED_AGVS) return "shared_agvs must be in 0..4";
    if (c.shared_agvs >= 2 && (c.num_cells != 1 || c.long_streams)) return "the shared floor (shared_agvs >= 2) needs num_cells = 1 and long_streams = 0";
    for (int a = 0; a < FJSP_NUM_LOCATIONS; a++)
        for (int b = a + 1; b < FJSP_NUM_LOCATIONS; b++)
            if (c.pos[a][0] == c.pos[b][0] && c.pos[a][1] == c.pos[b][1]) return "station positions must be distinct";
    memset(P, 0, sizeof(*P));
    for (int a = 0; a < FJSP_NUM_LOCATIONS; a++) {
        P->pos_row[a] = c.pos[a][0], P->pos_col[a] = c.pos[a][1];
        for (int b = 0; b < FJSP_NUM_LOCATIONS; b++) {
            int d = abs(c.pos[a][0] - c.pos[b][0]) + abs(c.pos[a][1] - c.pos[b][1]);
            P->dist[a][b] = d;
            P->delay[a][b] = d / (c.agv_speed * c.step_size);
            if (P->delay[a][b] > 60000) {
                snprintf(msg, sizeof msg, "distance %d too large", d);
                return msg;
            }
        }
    }
    P->small_steps = c.proc_small / c.step_size, P->big_steps = c.proc_big / c.step_size;
    P->pack_steps = c.proc_pack / c.step_size;
    P->max_episode_steps = c.max_episode_steps, P->storage_capacity = c.storage_capacity;
    P->pack_capacity = c.pack_capacity;
    P->num_trays = c.num_trays;
    P->trays_total = c.num_trays < 1000 ? c.num_trays : 1000;  // FJSPSimulation.py:96
    if (!c.long_streams && P->trays_total > 255) P->trays_total = 255;  // compact: an episode of <= 241 steps allocates <= 241 trays
    P->arrival_q16 = c.arrival_prob_q16, P->arrival_max = c.arrival_max_orders;
    P->step_size = c.step_size;
    P->progress_tab[0] = 0.0f;
    for (int L = 1; L < 256; L++) P->progress_tab[L] = (float)((1.0 / (double)L) * 100.0);
    return nullptr;
}

// FJSP_BOUNDS_CHECK (set by the test-only host harness build): every word index the step function computes is checked
// against the env's word count — `ld/st` must stay inside the dynamically indexed words, `ld_hot/st_hot` outside them.
// The CUDA kernels run the same index arithmetic, so a clean run of the CPU suite under the check is the stand-in for
// compute-sanitizer (closed on this pool) as far as the step function's addressing goes.
#ifdef FJSP_BOUNDS_CHECK
#define FJSP_CHECK_IDX(cond) do { if (!(cond)) { fprintf(stderr, "fjsp bounds check failed: %s (index %d, line %d)\n", #cond, i, __LINE__); abort(); } } while (0)
#else
#define FJSP_CHECK_IDX(cond) do { } while (0)
#endif
template <bool LONG_ = false>
struct ArrayStateT {
    static constexpr bool LONG = LONG_;
    u32* w;
    int dyn_end = 1 << 30, total = 1 << 30;  // word counts of the env (set by the harness when the check is compiled in)
    uint64_t seed = 0, genv = 0;             // long layout: the order / arrival streams (see fjsp_kernels.cuh)
    const FjspOrderRec* otab = nullptr;
    u32* rq = nullptr;                       // long layout: this env's ready FIFO (READY_FIFO_WORDS entries)
    FJSP_HD u32 rq_ld(int i) const { FJSP_CHECK_IDX(i >= 0 && i < READY_FIFO_WORDS); return rq[i & (READY_FIFO_WORDS - 1)]; }
    FJSP_HD void rq_st(int i, u32 v) { FJSP_CHECK_IDX(i >= 0 && i < READY_FIFO_WORDS); rq[i & (READY_FIFO_WORDS - 1)] = v; }
    static constexpr int DYN0 = WM<LONG_>::W_DYN0;
    FJSP_HD u32 ld(int i) const { FJSP_CHECK_IDX(i >= DYN0 && i < dyn_end); return w[i]; }
    FJSP_HD void st(int i, u32 v) { FJSP_CHECK_IDX(i >= DYN0 && i < dyn_end); w[i] = v; }
    FJSP_HD u32 ld_hot(int i) const { FJSP_CHECK_IDX((i >= 0 && i < DYN0) || (i >= dyn_end && i < total)); return w[i]; }
    FJSP_HD void st_hot(int i, u32 v) { FJSP_CHECK_IDX((i >= 0 && i < DYN0) || (i >= dyn_end && i < total)); w[i] = v; }
    FJSP_HD u32 or_word(int i, u32 v) {
        FJSP_CHECK_IDX(i >= DYN0 && i < WM<LONG_>::W_POOL);
        const u32 old = w[i];
        w[i] = old | v;
        return old;
    }
    FJSP_HD u32 fetch_order(int o, u32 episode) const {
        if (otab) {
            bool bad = false;
            return order_from_rec(otab[o], bad);
        }
        return philox_order(seed, genv, episode, o);
    }
    FJSP_HD u32 arrival_draw(u32 episode, u32 step) const {
        u32 r[4];
        philox4x32_10((u32)genv, episode, step, 5u, (u32)seed, (u32)(seed >> 32), r);
        return r[0];
    }
};
using ArrayState = ArrayStateT<false>;
// exchange area of the cell-parallel step, host emulation (lanes run one after the other)
struct ArrayXchg {
    u32* w;
    FJSP_HD u32 ld(int i) const { return w[i]; }
    FJSP_HD void st(int i, u32 v) { w[i] = v; }
    FJSP_HD void atom_or(int i, u32 v) { w[i] |= v; }
    FJSP_HD void atom_add(int i, u32 v) { w[i] += v; }
    FJSP_HD void st16(int i16, u32 v) { reinterpret_cast<uint16_t*>(w)[i16] = (uint16_t)v; }
    FJSP_HD u32 ld16(int i16) const { return reinterpret_cast<const uint16_t*>(w)[i16]; }
};

// is_processed of the products of every order slot: finished trays (compact: packaged bits and live pool records; long:
// slot word B), plus the products a busy machine has flagged so far.  is_processed is a property of the products,
// whatever cell their tray is in: every cell's pool and machines are scanned.
template <int K, bool LONG>
inline void processed_by_slot(const u32* words, const Params& P, int32_t* proc) {
    using W = WM<LONG>;
    ArrayStateT<LONG> s{const_cast<u32*>(words)};
    s.dyn_end = Lay<K, LONG>::DYN_END, s.total = Lay<K, LONG>::TOTAL;
    Hot h;
    load_hot(s, h);
    for (int sl = 0; sl < W::SLOTS; sl++)
        proc[sl] = LONG ? (int32_t)(words[W::W_ORDER_B + sl] & 0x1ffu) : (int32_t)ord_packaged(words[W::W_ORDER + sl]);
    for (int cc = 0; cc < K; cc++) {
        HotCell x;
        load_cell<K>(s, cc, x);
        const int xb = pool_base<LONG>(cc);
        uint64_t free_bits = (uint64_t)x.free_lo | ((uint64_t)x.free_hi << 32);
        for (int slot = 0; slot < FJSP_POOL_SLOTS; slot++) {
            if ((free_bits >> slot) & 1u) continue;
            u32 r = words[xb + slot];
            if (rec_processed(r)) proc[rec_order(r)] |= (int32_t)(((1u << rec_count(r)) - 1u) << rec_first(r));
        }
        for (int i = 0; i < 2; i++) {
            const Mach& m = x.m[i];
            if (!m.busy) continue;  // products flagged so far: i <= (last executed step - start) / per
            u32 r = words[xb + m.cur];
            int per = i == 0 ? P.small_steps : P.big_steps;
            int done = (h.step - 1 - m.start) / per;
            if (done > rec_count(r)) done = rec_count(r);
            if (done < 0) done = 0;
            proc[rec_order(r)] |= (int32_t)(((1u << done) - 1u) << rec_first(r));
        }
    }
}

// long layout: the slot an order occupies, or -1 (queued, loaded but untouched, or complete and retired)
template <bool LONG>
inline int slot_of_order(const u32* words, int order) {
    using W = WM<LONG>;
    if (!LONG) return order;
    const uint64_t free_bits = (uint64_t)words[W::W_SLOT_FREE] | ((uint64_t)words[W::W_SLOT_FREE + 1] << 32);
    for (int sl = 0; sl < W::SLOTS; sl++)
        if (!((free_bits >> sl) & 1u) && (int)((words[W::W_ORDER_B + sl] >> 9) & 0xfffu) == order) return sl;
    return -1;
}

// per-order records for orders [first, first + count): {packaged_mask, processed_mask, complete, completion_step}.
// Long layout: an order that is in process reads from its slot; one that is popped, has no slot and precedes the order
// whose tray was taken last is complete and retired (all bits, completion step unknown: -1); anything else is untouched.
template <int K, bool LONG>
inline void export_orders_k(const u32* words, const Params& P, int first, int count, int32_t* out4) {
    using W = WM<LONG>;
    ArrayStateT<LONG> s{const_cast<u32*>(words)};
    s.dyn_end = Lay<K, LONG>::DYN_END, s.total = Lay<K, LONG>::TOTAL;
    Hot h;
    load_hot(s, h);
    int32_t proc[W::SLOTS];
    processed_by_slot<K, LONG>(words, P, proc);
    for (int i = 0; i < count; i++) {
        const int o = first + i;
        int32_t* r = out4 + 4 * i;
        r[0] = r[1] = r[2] = 0, r[3] = -1;
        if (o < 0 || o >= h.num_orders) continue;
        if (LONG && o >= h.next_order) continue;                      // still in the queue
        const int sl = slot_of_order<LONG>(words, o);
        if (sl < 0) {
            if (LONG && h.act_order >= 0 && o <= h.act_order) r[0] = r[1] = 0x1ff, r[2] = 1;   // retired: complete by construction
            continue;
        }
        const u32 ow = words[W::W_ORDER + sl];
        r[0] = (int32_t)ord_packaged(ow);
        r[1] = proc[sl];
        const int cs = LONG ? (int)(words[W::W_ORDER_C + sl] & 0xffffu) : (int)((words[W::W_CSTEP + (sl >> 2)] >> ((sl & 3) * 8)) & 255u);
        r[2] = cs != 0, r[3] = cs - 1;
    }
}

// ---- canonical record S from packed words: shared pickup station / orders + the given cell ----
// Long layout (`rq` = the env's ready FIFO): the per-order arrays describe orders [order_base, order_base + 32) with
// order_base = max(0, next_order - 32) (the 32 most recently popped), tray entries are FJSP_TRAY_ENTRY_LONG, product
// ids order * 100 + idx as in the reference.  Orders that are loaded but untouched read as zeros, retired ones as complete.
template <int K, bool LONG>
inline void export_canon_k(const u32* words, const u32* rq, const Params& P, int cell, FjspCanonState* out, int32_t* order_base_out = nullptr) {
    using W = WM<LONG>;
    ArrayStateT<LONG> s{const_cast<u32*>(words)};
    s.dyn_end = Lay<K, LONG>::DYN_END, s.total = Lay<K, LONG>::TOTAL;
    Hot h;
    HotCell hc;
    load_hot(s, h);
    load_cell<K>(s, cell, hc);
    const int pb = pool_base<LONG>(cell);
    FjspCanonState& c = *out;
    memset(&c, 0, sizeof(c));
    const int order_base = LONG ? (h.next_order > FJSP_MAX_ORDERS ? h.next_order - FJSP_MAX_ORDERS : 0) : 0;
    if (order_base_out) *order_base_out = order_base;
    auto slot_word = [&](int slot) { return words[W::W_ORDER + slot]; };
    auto order_id = [&](int slot) { return LONG ? (int)((words[W::W_ORDER_B + slot] >> 9) & 0xfffu) : slot; };
    auto alloc_idx = [&](int slot, int first) {
        if (LONG) return (int)((words[W::W_ORDER_C + slot] >> 16) & 0xfffu) + popc32(ord_cut(slot_word(slot)) & ((1u << first) - 1u));
        int a = 0;
        for (int q = 0; q < slot; q++) a += popc32(ord_cut(slot_word(q))) + 1;
        return a + popc32(ord_cut(slot_word(slot)) & ((1u << first) - 1u));
    };
    auto make_entry = [&](int alloc, int order, int first, int count) {
        const int id = P.num_trays - 1 - alloc;
        return LONG ? FJSP_TRAY_ENTRY_LONG(id, order, first, count) : FJSP_TRAY_ENTRY(id, order, first, count);
    };
    auto entry = [&](int slot, int first, int count) { return make_entry(alloc_idx(slot, first), order_id(slot), first, count); };
    auto rec_entry = [&](u32 r) { return entry(rec_order(r), rec_first(r), rec_count(r)); };
    auto walk = [&](const Fifo& f, int32_t* dst, int cap) {
        for (int i = 0; i < cap; i++) dst[i] = -1;
        int slot = f.head;
        for (int i = 0; i < f.len && i < cap; i++) {
            u32 r = words[pb + slot];
            dst[i] = rec_entry(r);
            slot = rec_next(r);
        }
        return f.len;
    };
    c.current_step = h.step, c.num_orders = h.num_orders, c.fault = h.fault;
    c.agv_row = P.pos_row[hc.agv_loc], c.agv_col = P.pos_col[hc.agv_loc];
    c.agv_carry = hc.carry ? rec_entry(words[pb + hc.carry - 1]) : -1;
    c.agv_is_moving = hc.agv_moving;
    c.ps_order_queue_len = h.num_orders - h.next_order;
    c.ps_current_order = h.cur_order;
    c.ps_product_idx = h.prod_idx;
    if (h.cur_tray_count > 0)
        c.ps_current_tray = LONG ? make_entry(h.alloc_count - 1, h.cur_order, h.prod_idx - h.cur_tray_count, h.cur_tray_count)
                                 : entry(h.cur_order, h.prod_idx - h.cur_tray_count, h.cur_tray_count);
    else
        c.ps_current_tray = -1;
    c.ps_trays_at_station = (P.num_trays < 1000 ? P.num_trays : 1000) - h.alloc_count;
    for (int i = 0; i < FJSP_CANON_PS_READY; i++) c.ps_ready[i] = -1;
    if (LONG) {  // the ready FIFO: entry index = allocation index of the tray
        for (int i = 0; i < h.ready_count && i < FJSP_CANON_PS_READY; i++) {
            const u32 e = rq ? rq[(h.ready_order + i) & (READY_FIFO_WORDS - 1)] : 0u;
            c.ps_ready[i] = make_entry(h.ready_order + i, rq_order(e), rq_first(e), rq_count(e));
        }
        c.ps_ready_n = h.ready_count;
    } else {
        int o = h.ready_order, f = h.ready_idx;
        for (int i = 0; i < h.ready_count && i < FJSP_CANON_PS_READY; i++) {
            u32 ow = slot_word(o);
            u32 cuts = ord_cut(ow) >> f;
            int cnt = cuts ? ctz32(cuts) + 1 : ord_n(ow) - f;
            c.ps_ready[i] = entry(o, f, cnt);
            f += cnt;
            if (f >= ord_n(ow)) o++, f = 0;
        }
        c.ps_ready_n = h.ready_count;
    }
    for (int i = 0; i < 2; i++) {
        const Mach& m = hc.m[i];
        c.machine[i].is_busy = m.busy;
        c.machine[i].current_tray = m.has_cur ? rec_entry(words[pb + m.cur]) : -1;
        c.machine[i].progress_done = m.prog;
        c.machine[i].queue_n = walk(m.q, c.machine[i].queue, FJSP_CANON_MAXQ);
        c.machine[i].ready_n = walk(m.r, c.machine[i].ready, FJSP_CANON_MAXQ);
    }
    c.storage_n = walk(hc.storage, c.storage, FJSP_CANON_MAXQ);
    for (int i = 0; i < 4; i++) {
        const Pack& p = hc.p[i];
        c.pack[i].is_busy = p.busy;
        c.pack[i].current_product = p.hascur ? p.cur_order * 100 + p.cur_idx : -1;
        c.pack[i].progress_L = p.progL;
        c.pack[i].products_completed = p.completed;
        c.pack[i].users = p.users;
        for (int k = 0; k < FJSP_CANON_MAXPQ; k++) c.pack[i].queue[k] = -1;
        int slot = p.q.head, n = 0;
        for (int k = 0; k < p.q.len; k++) {
            u32 r = words[pb + slot];
            for (int j = 0; j < rec_count(r) && n < FJSP_CANON_MAXPQ; j++) c.pack[i].queue[n++] = order_id(rec_order(r)) * 100 + rec_first(r) + j;
            slot = rec_next(r);
        }
        c.pack[i].queue_n = p.qcount;
    }
    // per-order arrays: orders order_base .. order_base + 31 (compact: 0..31)
    int32_t out4[4 * FJSP_MAX_ORDERS];
    export_orders_k<K, LONG>(words, P, order_base, FJSP_MAX_ORDERS, out4);
    for (int i = 0; i < FJSP_MAX_ORDERS; i++) {
        c.packaged_mask[i] = out4[4 * i], c.processed_mask[i] = out4[4 * i + 1];
        c.order_complete[i] = out4[4 * i + 2], c.order_completion_step[i] = out4[4 * i + 3];
    }
    c.total_products_packaged = h.total_packaged;
    c.completed_orders = h.completed_orders;
}

template <class F>
inline void dispatch_layout(int cells, bool long_streams, F&& f) {
#define FJSP_CASE(KK)                                                       \
    case KK:                                                                \
        if (long_streams) f(std::integral_constant<int, KK>{}, std::true_type{});   \
        else f(std::integral_constant<int, KK>{}, std::false_type{});       \
        break;
    switch (cells) {
        FJSP_CASE(1) FJSP_CASE(2) FJSP_CASE(3)
        default:
            if (long_streams) f(std::integral_constant<int, 4>{}, std::true_type{});
            else f(std::integral_constant<int, 4>{}, std::false_type{});
            break;
    }
#undef FJSP_CASE
}

inline void export_canon(const u32* words, const u32* rq, const Params& P, int cells, bool long_streams, int cell, FjspCanonState* out,
                         int32_t* order_base = nullptr) {
    dispatch_layout(cells, long_streams, [&](auto k, auto l) { export_canon_k<decltype(k)::value, decltype(l)::value>(words, rq, P, cell, out, order_base); });
}
// shared floor (fjsp_shared.h): the canonical record of AGV `agv` = the reference shop's record with that AGV's word in
// place of agv_0's (the stations, the pickup station and the orders are common to all AGVs)
inline void export_canon_shared(const u32* words, const Params& P, int agv, FjspCanonState* out) {
    u32 w[FJSP_STATE_WORDS];
    memcpy(w, words, sizeof(w));
    if (agv >= 1) w[W_AGV] = words[FJSP_STATE_WORDS + agv - 1];
    export_canon_k<1, false>(w, nullptr, P, 0, out);
}
inline void export_orders(const u32* words, const Params& P, int cells, bool long_streams, int first, int count, int32_t* out4) {
    dispatch_layout(cells, long_streams, [&](auto k, auto l) { export_orders_k<decltype(k)::value, decltype(l)::value>(words, P, first, count, out4); });
}

}  // namespace fjsp
#endif
