// fjsp_host.h — host-side helpers shared by the C-ABI library and the test harness:
//   * FjspConfig validation + derivation of the kernel Params,
//   * decode of one env's packed words into the canonical record S (include/fjsp_b200.h).
// No simulation logic lives here.
#ifndef FJSP_HOST_H
#define FJSP_HOST_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
    if (c.max_episode_steps < 1 || c.max_episode_steps > 240) return "max_episode_steps must be in 1..240 (bounds the 64-slot tray pool and the 6/8-bit queue counters)";
    if (c.pack_capacity < 1 || c.pack_capacity > 31) return "pack_capacity must be in 1..31";
    if (c.storage_capacity < 0) return "storage_capacity must be >= 0";
    if (c.num_trays < 0 || c.num_trays > 65535) return "num_trays must be in 0..65535";
    if (c.num_cells < 1 || c.num_cells > FJSP_MAX_CELLS) return "num_cells must be in 1..4";
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
    if (P->trays_total > 255) P->trays_total = 255;            // an episode of <= 253 steps allocates <= 253 trays
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
struct ArrayState {
    u32* w;
    int dyn_end = 1 << 30, total = 1 << 30;  // word counts of the env (set by the harness when the check is compiled in)
    FJSP_HD u32 ld(int i) const { FJSP_CHECK_IDX(i >= W_CSTEP && i < dyn_end); return w[i]; }
    FJSP_HD void st(int i, u32 v) { FJSP_CHECK_IDX(i >= W_CSTEP && i < dyn_end); w[i] = v; }
    FJSP_HD u32 ld_hot(int i) const { FJSP_CHECK_IDX((i >= 0 && i < W_CSTEP) || (i >= dyn_end && i < total)); return w[i]; }
    FJSP_HD void st_hot(int i, u32 v) { FJSP_CHECK_IDX((i >= 0 && i < W_CSTEP) || (i >= dyn_end && i < total)); w[i] = v; }
    FJSP_HD u32 or_word(int i, u32 v) {
        FJSP_CHECK_IDX(i >= W_CSTEP && i < W_POOL);
        const u32 old = w[i];
        w[i] = old | v;
        return old;
    }
};
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

// ---- canonical record S from packed words: shared pickup station / orders + the given cell ----
template <int K>
inline void export_canon_k(const u32* words, const Params& P, int cell, FjspCanonState* out) {
    ArrayState s{const_cast<u32*>(words)};
    s.dyn_end = Lay<K>::DYN_END, s.total = Lay<K>::TOTAL;
    Hot h;
    HotCell hc;
    load_hot(s, h);
    load_cell<K>(s, cell, hc);
    const int pb = pool_base(cell);
    FjspCanonState& c = *out;
    memset(&c, 0, sizeof(c));
    auto order_word = [&](int o) { return words[W_ORDER + o]; };
    auto alloc_idx = [&](int o, int first) {
        int a = 0;
        for (int q = 0; q < o; q++) a += popc32(ord_cut(order_word(q))) + 1;
        return a + popc32(ord_cut(order_word(o)) & ((1u << first) - 1u));
    };
    auto entry = [&](int o, int first, int count) {
        int id = P.num_trays - 1 - alloc_idx(o, first);
        return FJSP_TRAY_ENTRY(id, o, first, count);
    };
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
    c.ps_current_order = h.cur_order == 63 ? -1 : h.cur_order;
    c.ps_product_idx = h.prod_idx;
    c.ps_current_tray = h.cur_tray_count > 0 ? entry(h.cur_order, h.prod_idx - h.cur_tray_count, h.cur_tray_count) : -1;
    c.ps_trays_at_station = (P.num_trays < 1000 ? P.num_trays : 1000) - h.alloc_count;
    for (int i = 0; i < FJSP_CANON_PS_READY; i++) c.ps_ready[i] = -1;
    {
        int o = h.ready_order, f = h.ready_idx;
        for (int i = 0; i < h.ready_count && i < FJSP_CANON_PS_READY; i++) {
            u32 ow = order_word(o);
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
        c.pack[i].current_product = p.hascur ? (p.curprod & 31) * 100 + (p.curprod >> 5) : -1;
        c.pack[i].progress_L = p.progL;
        c.pack[i].products_completed = p.completed;
        c.pack[i].users = p.users;
        for (int k = 0; k < FJSP_CANON_MAXPQ; k++) c.pack[i].queue[k] = -1;
        int slot = p.q.head, n = 0;
        for (int k = 0; k < p.q.len; k++) {
            u32 r = words[pb + slot];
            for (int j = 0; j < rec_count(r) && n < FJSP_CANON_MAXPQ; j++) c.pack[i].queue[n++] = rec_order(r) * 100 + rec_first(r) + j;
            slot = rec_next(r);
        }
        c.pack[i].queue_n = p.qcount;
    }
    for (int o = 0; o < FJSP_MAX_ORDERS; o++) {
        u32 ow = order_word(o);
        c.packaged_mask[o] = (int32_t)ord_packaged(ow);
        c.processed_mask[o] = (int32_t)ord_packaged(ow);
        int cs = (int)((words[W_CSTEP + (o >> 2)] >> ((o & 3) * 8)) & 255u);
        c.order_complete[o] = cs != 0;
        c.order_completion_step[o] = cs - 1;
    }
    // is_processed is a property of the products, whatever cell their tray is in: scan every cell's pool and machines
    for (int cc = 0; cc < K; cc++) {
        HotCell x;
        load_cell<K>(s, cc, x);
        const int xb = pool_base(cc);
        uint64_t free_bits = (uint64_t)x.free_lo | ((uint64_t)x.free_hi << 32);
        for (int slot = 0; slot < FJSP_POOL_SLOTS; slot++) {
            if ((free_bits >> slot) & 1u) continue;
            u32 r = words[xb + slot];
            if (rec_processed(r)) c.processed_mask[rec_order(r)] |= (int32_t)(((1u << rec_count(r)) - 1u) << rec_first(r));
        }
        for (int i = 0; i < 2; i++) {
            const Mach& m = x.m[i];
            if (!m.busy) continue;  // products flagged so far: i <= (last executed step - start) / per
            u32 r = words[xb + m.cur];
            int per = i == 0 ? P.small_steps : P.big_steps;
            int done = (h.step - 1 - m.start) / per;
            if (done > rec_count(r)) done = rec_count(r);
            if (done < 0) done = 0;
            c.processed_mask[rec_order(r)] |= (int32_t)(((1u << done) - 1u) << rec_first(r));
        }
    }
    c.total_products_packaged = h.total_packaged;
    c.completed_orders = h.completed_orders;
}

inline void export_canon(const u32* words, const Params& P, int cells, int cell, FjspCanonState* out) {
    switch (cells) {
        case 1: export_canon_k<1>(words, P, cell, out); break;
        case 2: export_canon_k<2>(words, P, cell, out); break;
        case 3: export_canon_k<3>(words, P, cell, out); break;
        default: export_canon_k<4>(words, P, cell, out); break;
    }
}

}  // namespace fjsp
#endif
