// fjsp_shared.h — SHARED FLOOR (include/fjsp_b200.h "shared floor"): A = 2..4 AGVs on ONE set of stations, each station
// position holding at most one AGV.  A builder-defined extension of the reference shop (one AGV,
// /root/reference/FJSPSimulation.py:62-82, agents/AGVAgent.py:180-396); A = 1 is the reference and is served by the
// ordinary step (fjsp_core.h).  Host + device code: the same functions run in the kernels (fjsp_shared.cuh) and in the
// test-only host harness.
//
// Packed state: the reference shop's 128 words (fjsp_core.h; word W_AGV is agv_0) + words 128..130 = agv_1..agv_3 in the
// format of W_AGV (loc3 | moving1<<3 | target3<<4 | arrive16<<7 | carry7<<23), word 131 unused: 132 words = 528 B.
// One pool of 64 tray records serves every AGV and station.
//
// Step: pickup station -> agv_0 .. agv_{A-1} (each sees the positions taken by the others AFTER the earlier AGVs of this
// step acted: an earlier AGV that leaves a position frees it at once, one that is granted a position holds it at once)
// -> machines -> packaging stations -> run phase.  The AGV logic is the reference's (act_agv / observe_agv of fjsp_core.h);
// the only new rule is the occupancy test on moves.
#pragma once

#include "fjsp_core.h"

namespace fjsp {

template <int A>
struct ShLay {
    static_assert(A >= 2 && A <= FJSP_MAX_SHARED_AGVS, "shared floor: 2..4 AGVs");
    static constexpr int AGVS = A;
    static constexpr int AGENTS = FJSP_SHARED_AGENTS(A), ACT = FJSP_SHARED_ACT_DIM(A), OBS = FJSP_SHARED_OBS_DIM(A), MASK = FJSP_SHARED_MASK_DIM(A);
    static constexpr int TOTAL = FJSP_SHARED_STATE_WORDS;   // words per env
    static constexpr int W_AGVX = Lay<1>::TOTAL;            // word of agv_j (j >= 1) = W_AGVX + j - 1
    static constexpr int DYN0 = Lay<1>::DYN0, DYN_END = Lay<1>::DYN_END;   // the dynamically indexed words are the reference shop's
};
constexpr int SH_START_LOC[FJSP_MAX_SHARED_AGVS] = {LOC_PICKUP, LOC_STORAGE, LOC_SMALL, LOC_BIG};
FJSP_HD int sh_start_loc(int j) { return j == 0 ? LOC_PICKUP : j == 1 ? LOC_STORAGE : j == 2 ? LOC_SMALL : LOC_BIG; }

// AGV word <-> the AGV fields of a HotCell (the stations of the HotCell stay as they are)
FJSP_HD u32 agv_get(const HotCell& hc) {
    return (u32)hc.agv_loc | ((u32)hc.agv_moving << 3) | ((u32)hc.agv_target << 4) | ((u32)hc.agv_arrive << 7) | ((u32)hc.carry << 23);
}
FJSP_HD void agv_set(HotCell& hc, u32 w) {
    hc.agv_loc = w & 7u, hc.agv_moving = (w >> 3) & 1u, hc.agv_target = (w >> 4) & 7u;
    hc.agv_arrive = (w >> 7) & 0xffffu, hc.carry = (w >> 23) & 127u;
}
// the station position an AGV holds: where it stands, or where it is going
FJSP_HD u32 agv_holds(u32 w) { return 1u << (((w >> 3) & 1u) ? ((w >> 4) & 7u) : (w & 7u)); }
// AGV arrival on the word (AGVAgent.py:387-396)
FJSP_HD u32 agv_word_run(u32 w, int k) {
    if (((w >> 3) & 1u) && ((w >> 7) & 0xffffu) == (u32)(k & 0xffff)) w = (w & ~0xfu) | ((w >> 4) & 7u);   // loc = target, moving = 0
    return w;
}
template <int A>
FJSP_HD u32 occupied_by_others(const u32* w, int j) {
    u32 occ = 0u;
#pragma unroll
    for (int i = 0; i < A; i++)
        if (i != j) occ |= agv_holds(w[i]);
    return occ;
}

template <int A>
struct ShOut {
    float* obs;                      // ShLay<A>::OBS floats, written in place
    u32 mask[ShLay<A>::MASK / 4];
    float reward[ShLay<A>::ACT];
    u32 flags;                       // terminated | truncated<<8 | fault<<16 | was_reset<<24
    u32 results[ShLay<A>::ACT / 4];
    int32_t info[4];
    long long reward_units;          // sum over the agents of 10 * AGENTS * reward (exact integers)
};

// observation: pickup station (7) | 13 per AGV | machines (6) | packaging stations (12); masks 3 | 8 per AGV | 6 | 12.
// `c0` holds the stations and agv_0; `ax` = words of agv_1...
template <int A, class S, class O>
FJSP_HD void shared_observe(S& s, const Params& P, const Hot& h, HotCell& c0, const u32* ax, O obs, u32* mw) {
#pragma unroll
    for (int i = 0; i < ShLay<A>::MASK / 4; i++) mw[i] = 0u;
    observe_shared(s, P, h, obs, mw);
    u32 w[A];
    w[0] = agv_get(c0);
#pragma unroll
    for (int j = 1; j < A; j++) w[j] = ax[j - 1];
#pragma unroll
    for (int j = 0; j < A; j++) {
        agv_set(c0, w[j]);
        observe_agv<true>(s, P, h, c0, 0, obs.at(7 + 13 * j), mw, 3 + 8 * j, occupied_by_others<A>(w, j));
    }
    agv_set(c0, w[0]);
    observe_stations(P, c0, obs.at(7 + 13 * A), mw, 3 + 8 * A);
}

// The few dynamically indexed words an observation reads — the order being loaded at the pickup station, and per AGV the
// carried tray's record and its order word — taken BEFORE the tile's shared memory is handed to the bulk store, so that
// the observation rows can be staged in the same shared memory afterwards (fjsp_shared.cuh).  A word accessor like the
// tile columns: ld(w) answers from the snapshot (any other index would be a bug: it returns 0).
template <int A>
struct ObsSnapshot {
    static constexpr bool LONG = false;
    int idx[1 + 2 * A];
    u32 val[1 + 2 * A];
    FJSP_HD u32 ld(int w) const {
        u32 v = 0u;
#pragma unroll
        for (int i = 0; i < 1 + 2 * A; i++)
            if (idx[i] == w) v = val[i];
        return v;
    }
    template <class S>
    FJSP_HD void take(S& s, const Hot& h, const HotCell& c0, const u32* ax) {
#pragma unroll
        for (int i = 0; i < 1 + 2 * A; i++) idx[i] = -1, val[i] = 0u;
        if (h.cur_order >= 0) idx[0] = W_ORDER + h.cur_order, val[0] = s.ld(idx[0]);
        const int pb = pool_base<false>(0);
#pragma unroll
        for (int j = 0; j < A; j++) {
            const int carry = j == 0 ? (int)c0.carry : (int)((ax[j - 1] >> 23) & 127u);
            if (carry != 0) {
                idx[1 + 2 * j] = pb + carry - 1, val[1 + 2 * j] = s.ld(idx[1 + 2 * j]);
                idx[2 + 2 * j] = W_ORDER + rec_order(val[1 + 2 * j]), val[2 + 2 * j] = s.ld(idx[2 + 2 * j]);
            }
        }
    }
};

// One step.  h / c0 (stations + agv_0) / ax (agv_1.. words) live with the caller (registers).  OBSERVE = false: the
// caller computes the observation and the masks itself afterwards (shared_observe), e.g. from an ObsSnapshot.
template <int A, bool OBSERVE = true, class S>
FJSP_HD void shared_step(S& s, const Params& P, Hot& h, HotCell& c0, u32* ax, const int* a, ShOut<A>& out) {
    using L = ShLay<A>;
    constexpr int AG = L::AGENTS;
    const int k = h.step;
    const int orders_before = h.completed_orders, products_before = h.total_packaged;
    int local10[L::ACT];
    u32 res[L::ACT];
#pragma unroll
    for (int i = 0; i < L::ACT; i++) local10[i] = 0, res[i] = 0u;
    if (k > P.max_episode_steps) {  // stepped past the end without a reset: inert (fjsp_core.h step_env_hot)
#pragma unroll
        for (int i = 0; i < L::ACT; i++) out.reward[i] = 0.0f;
#pragma unroll
        for (int i = 0; i < L::ACT / 4; i++) out.results[i] = 0u;
        out.reward_units = 0;
        out.flags = (1u << 8) | ((u32)FJSP_FAULT_PAST_END << 16);
        out.info[0] = h.step, out.info[1] = h.completed_orders, out.info[2] = h.total_packaged, out.info[3] = 0;
        if (OBSERVE) shared_observe<A>(s, P, h, c0, ax, FloatSink{out.obs, P}, out.mask);
        return;
    }
    act_pickup(s, P, h, a[0], local10[0], res[0]);
    u32 w[A];
    w[0] = agv_get(c0);
#pragma unroll
    for (int j = 1; j < A; j++) w[j] = ax[j - 1];
#pragma unroll
    for (int j = 0; j < A; j++) {   // agent order: an earlier AGV's move is visible to the later ones
        agv_set(c0, w[j]);
        act_agv<true>(s, P, h, c0, 0, k, a[1 + j], local10[1 + j], res[1 + j], occupied_by_others<A>(w, j));
        w[j] = agv_get(c0);
    }
    agv_set(c0, w[0]);
    int pk_start[4];
    act_stations(s, P, h, c0, 0, k, a + 1 + A, local10 + 1 + A, res + 1 + A, pk_start);
    int dock_after = 0;
    run_cell(s, P, h, c0, 0, k, pk_start, dock_after);   // agv_0's arrival + machines + packaging stations
#pragma unroll
    for (int j = 1; j < A; j++) ax[j - 1] = agv_word_run(w[j], k);
    h.dock_mask = dock_after;
    // rewards: 10*AG*r_i = 10*(100*orders + 10*products) - step_size + AG*(10*local_i), one correctly rounded fp32 division
    // (fjsp_core.h step_env_hot; AG = 7 + A agents)
    {
        const int d_orders = h.completed_orders - orders_before, d_products = h.total_packaged - products_before;
        const int g = 10 * (100 * d_orders + 10 * d_products) - P.step_size;
        long long units = 0;
#pragma unroll
        for (int i = 0; i < L::ACT; i++) {
            if (i < AG) {
                const int n = g + AG * local10[i];
                out.reward[i] = (float)n / (float)(10 * AG);
                units += n;
            } else {
                out.reward[i] = 0.0f;
            }
        }
        out.reward_units = units;
    }
    const int all_done = all_orders_done<S>(P, h);
    const int truncated = k >= P.max_episode_steps;
    out.flags = (u32)all_done | ((u32)truncated << 8) | ((u32)h.fault << 16);
#pragma unroll
    for (int i = 0; i < L::ACT / 4; i++) out.results[i] = res[4 * i] | (res[4 * i + 1] << 8) | (res[4 * i + 2] << 16) | (res[4 * i + 3] << 24);
    h.step = k + 1;
    out.info[0] = h.step, out.info[1] = h.completed_orders, out.info[2] = h.total_packaged, out.info[3] = 0;
    if (OBSERVE) shared_observe<A>(s, P, h, c0, ax, FloatSink{out.obs, P}, out.mask);
}

// words of agv_1.. of a fresh env
template <int A, class S>
FJSP_HD void shared_reset_agvs(S& s) {
#pragma unroll
    for (int j = 1; j < 4; j++) s.st_hot(ShLay<A>::W_AGVX + j - 1, j < A ? (u32)sh_start_loc(j) : 0u);
    s.st_hot(ShLay<A>::W_AGVX + 3, 0u);
}
template <int A, class S>
FJSP_HD void shared_reset(S& s, const Params& P, int num_orders, const FjspOrderRec* orders, uint64_t seed, uint64_t genv, u32 episode) {
    reset_env<1>(s, P, num_orders, orders, seed, genv, episode);
    shared_reset_agvs<A>(s);
}
template <int A, class S>
FJSP_HD void shared_load_agvs(S& s, u32* ax) {
#pragma unroll
    for (int j = 1; j < A; j++) ax[j - 1] = s.ld_hot(ShLay<A>::W_AGVX + j - 1);
}
template <int A, class S>
FJSP_HD void shared_store_agvs(S& s, const u32* ax) {
#pragma unroll
    for (int j = 1; j < A; j++) s.st_hot(ShLay<A>::W_AGVX + j - 1, ax[j - 1]);
}

// Philox actions: the reference stream's eight values serve (pickup station, agv_0, machines, stations); agv_1.. draw from
// counter word 3 = 17 (u16 lane j - 1)
template <int A>
FJSP_HD void philox_actions_shared(uint64_t seed, uint64_t genv, uint64_t t, int* a) {
    int base[8];
    philox_actions(seed, genv, t, base);
    u32 r[4];
    philox4x32_10((u32)genv, (u32)t, (u32)(t >> 32), 17u, (u32)seed, (u32)(seed >> 32), r);
#pragma unroll
    for (int i = 0; i < ShLay<A>::ACT; i++) a[i] = 0;
    a[0] = base[0], a[1] = base[1];
#pragma unroll
    for (int j = 1; j < A; j++) {
        const u32 hw = ((j - 1) & 1) ? (r[(j - 1) >> 1] >> 16) : (r[(j - 1) >> 1] & 0xffffu);
        a[1 + j] = (int)((hw * 8u) >> 16);
    }
#pragma unroll
    for (int i = 0; i < 6; i++) a[1 + A + i] = base[2 + i];
}

}  // namespace fjsp
