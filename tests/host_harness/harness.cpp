// tests/host_harness/harness.cpp — TEST-ONLY host build of the device step function.
//
// Compiles multi_agent_rl_for_fjsp_b200/csrc/fjsp_core.h (the code the CUDA kernels run per env) with g++
// so the CPU-only test suite (`pytest -m "not gpu"`) can check the packed-state logic against the oracle
// on millions of steps without a GPU.  It is NOT part of the product: nothing in the package or in
// libfjsp_b200.so loads it, and the product has no CPU execution path.
#include <stdlib.h>

#include "../../multi_agent_rl_for_fjsp_b200/csrc/fjsp_host.h"
#include "../../multi_agent_rl_for_fjsp_b200/csrc/fjsp_shared.h"

using namespace fjsp;

struct HostEnv {
    Params P;
    int cells;
    int shared;   // shared floor: number of AGVs (0 = off)
    bool long_streams;
    uint64_t seed, genv;                       // long layout: order / arrival streams
    FjspOrderRec otab[FJSP_LONG_MAX_ORDERS];   // long layout: explicit order table of the current episode
    int otab_n;
    u32 rq[FJSP_LONG_READY_FIFO];              // long layout: the ready FIFO (a side buffer in HBM on the device)
    u32 words[FJSP_STATE_WORDS_LONG_K(FJSP_MAX_CELLS)];
};

#define DISPATCH_K1(e, ...)                                  \
    switch ((e)->cells) {                                   \
        case 1: { constexpr int K = 1; __VA_ARGS__; } break;  \
        case 2: { constexpr int K = 2; __VA_ARGS__; } break;  \
        case 3: { constexpr int K = 3; __VA_ARGS__; } break;  \
        default: { constexpr int K = 4; __VA_ARGS__; } break; \
    }
// K and the layout; `s` is the env's word accessor of the matching type
#define DISPATCH_K(e, ...)                                                                       \
    if ((e)->long_streams) { constexpr bool LONG = true; auto s = env_state<true>(e); (void)LONG; (void)s; DISPATCH_K1(e, __VA_ARGS__) } \
    else { constexpr bool LONG = false; auto s = env_state<false>(e); (void)LONG; (void)s; DISPATCH_K1(e, __VA_ARGS__) }

// shared floor: A AGVs; `s` is the env's word accessor
#define DISPATCH_A(e, ...)                                    \
    {                                                         \
        auto s = env_state<false>(e);                         \
        (void)s;                                              \
        switch ((e)->shared) {                                \
            case 2: { constexpr int A = 2; __VA_ARGS__; } break;  \
            case 3: { constexpr int A = 3; __VA_ARGS__; } break;  \
            default: { constexpr int A = 4; __VA_ARGS__; } break; \
        }                                                     \
    }

static int state_words(const HostEnv* e) { return e->shared ? FJSP_SHARED_STATE_WORDS : e->long_streams ? FJSP_STATE_WORDS_LONG_K(e->cells) : FJSP_STATE_WORDS_K(e->cells); }
template <bool LONG>
static ArrayStateT<LONG> env_state(HostEnv* e) {
    ArrayStateT<LONG> s{e->words};
    s.dyn_end = WM<LONG>::W_POOL + 64 * e->cells, s.total = state_words(e);
    s.seed = e->seed, s.genv = e->genv, s.otab = e->otab_n > 0 ? e->otab : nullptr, s.rq = e->rq;
    return s;
}

extern "C" {

void* hh_create(const FjspConfig* cfg) {
    FjspConfig c;
    if (cfg) c = *cfg; else default_config(&c);
    HostEnv* e = (HostEnv*)calloc(1, sizeof(HostEnv));
    if (make_params(c, &e->P) != nullptr) {
        free(e);
        return nullptr;
    }
    e->cells = c.num_cells;
    e->long_streams = c.long_streams != 0;
    e->shared = c.shared_agvs >= 2 ? c.shared_agvs : 0;
    return e;
}
const char* hh_check_config(const FjspConfig* cfg) {
    Params P;
    const char* m = make_params(*cfg, &P);
    return m ? m : "";
}
void hh_destroy(void* p) { free(p); }
int hh_bounds_checked(void) {  // 1 when every word index of the step function is range-checked (fjsp_host.h)
#ifdef FJSP_BOUNDS_CHECK
    return 1;
#else
    return 0;
#endif
}

static void unpack_bytes(const u32* w, int n, uint8_t* dst) {
    for (int i = 0; i < n; i++) dst[i] = (uint8_t)((w[i >> 2] >> ((i & 3) * 8)) & 0xff);
}

void hh_observe(void* p, float* obs, int8_t* masks) {
    HostEnv* e = (HostEnv*)p;
    if (e->shared) {
        DISPATCH_A(e, {
            Hot h;
            HotCell c0;
            u32 ax[3] = {0u, 0u, 0u}, mw[ShLay<A>::MASK / 4];
            load_hot(s, h), load_cell<1>(s, 0, c0), shared_load_agvs<A>(s, ax);
            shared_observe<A>(s, e->P, h, c0, ax, FloatSink{obs, e->P}, mw);
            unpack_bytes(mw, ShLay<A>::MASK, (uint8_t*)masks);
        })
        return;
    }
    DISPATCH_K(e, {
        u32 mw[Lay<K>::MASK / 4];
        observe_env<K>(s, e->P, obs, mw);
        unpack_bytes(mw, Lay<K>::MASK, (uint8_t*)masks);
    })
}

void hh_reset(void* p, const FjspOrderRec* orders, int num_orders, uint64_t seed, uint64_t genv, uint32_t episode) {
    HostEnv* e = (HostEnv*)p;
    e->seed = seed, e->genv = genv, e->otab_n = 0;
    if (e->long_streams && orders) {  // the ring fills as orders are popped: keep the table for the episode
        // (with arrivals the table describes every order that can ever exist: arrival_max of them)
        e->otab_n = e->P.arrival_q16 > 0 && e->P.arrival_max > num_orders ? e->P.arrival_max : num_orders;
        memcpy(e->otab, orders, sizeof(FjspOrderRec) * (size_t)e->otab_n);
        orders = e->otab;
    }
    if (e->shared) {
        DISPATCH_A(e, shared_reset<A>(s, e->P, num_orders, orders, seed, genv, episode))
        return;
    }
    DISPATCH_K(e, reset_env<K>(s, e->P, num_orders, orders, seed, genv, episode))
}

void hh_step(void* p, const uint8_t* actions, float* obs, int8_t* masks, float* rewards, uint8_t* flags, uint8_t* results,
             int32_t* infos) {
    HostEnv* e = (HostEnv*)p;
    if (e->shared) {
        DISPATCH_A(e, {
            int a[ShLay<A>::ACT];
            for (int i = 0; i < ShLay<A>::ACT; i++) a[i] = actions[i];
            ShOut<A> out;
            out.obs = obs;
            Hot h;
            HotCell c0;
            u32 ax[3] = {0u, 0u, 0u};
            load_hot(s, h), load_cell<1>(s, 0, c0), shared_load_agvs<A>(s, ax);
            shared_step<A>(s, e->P, h, c0, ax, a, out);
            store_hot(s, h), store_cell<1>(s, 0, c0), shared_store_agvs<A>(s, ax);
            unpack_bytes(out.mask, ShLay<A>::MASK, (uint8_t*)masks);
            for (int i = 0; i < ShLay<A>::ACT; i++) rewards[i] = out.reward[i];
            flags[0] = out.flags & 0xff, flags[1] = (out.flags >> 8) & 0xff, flags[2] = (out.flags >> 16) & 0xff, flags[3] = 0;
            if (results) unpack_bytes(out.results, ShLay<A>::ACT, results);
            if (infos)
                for (int i = 0; i < 4; i++) infos[i] = out.info[i];
        })
        return;
    }
    DISPATCH_K(e, {
        int a[Lay<K>::ACT];
        for (int i = 0; i < Lay<K>::ACT; i++) a[i] = actions[i];
        StepOut<K> out;
        out.obs = obs;
        step_env<K, OBS_FLOAT>(s, e->P, a, out);
        unpack_bytes(out.mask, Lay<K>::MASK, (uint8_t*)masks);
        for (int i = 0; i < Lay<K>::ACT; i++) rewards[i] = out.reward[i];
        flags[0] = out.flags & 0xff, flags[1] = (out.flags >> 8) & 0xff, flags[2] = (out.flags >> 16) & 0xff, flags[3] = 0;
        if (results) unpack_bytes(out.results, Lay<K>::ACT, results);
        if (infos)
            for (int i = 0; i < 4; i++) infos[i] = out.info[i];
    })
}

// the same step twice from the same state: once into float tensors, once into a wire row (state advanced once)
void hh_step_wire(void* p, const uint8_t* actions, float* obs, int8_t* masks, float* rewards, uint8_t* flags, uint32_t* wire) {
    HostEnv* e = (HostEnv*)p;
    static thread_local HostEnv copy;
    copy = *e;
    hh_step(p, actions, obs, masks, rewards, flags, nullptr, nullptr);
    DISPATCH_K(&copy, {
        int a[Lay<K>::ACT];
        for (int i = 0; i < Lay<K>::ACT; i++) a[i] = actions[i];
        StepOut<K> out;
        out.obs = nullptr;
        step_env<K, OBS_WIRE>(s, copy.P, a, out);
        wire_row<K>(out, wire);
    })
}

// Cell-parallel step (fjsp_core.h), host emulation: the K lanes of the env run one after the other inside each phase
// (ascending or descending cell order: the result must not depend on it), "barriers" are the phase boundaries.
int hh_step_cells(void* p, const uint8_t* actions, float* obs, int8_t* masks, float* rewards, uint8_t* flags, uint8_t* results,
                  int32_t* infos, int reverse) {
    HostEnv* e = (HostEnv*)p;
    int done = 0;
    DISPATCH_K(e, {
        if constexpr (K >= 2) {
            constexpr int AG = Lay<K>::AGENTS;
            u32 xw[Xl<K, LONG>::WORDS];
            for (int i = 0; i < Xl<K, LONG>::WORDS; i++) xw[i] = 0u;
            ArrayXchg x{xw};
            CellLane L[K];
            int a7[K][7];
            for (int c = 0; c < K; c++) {
                L[c].c = c;
                load_hot(s, L[c].h);
                load_cell<K>(s, c, L[c].hc);
                for (int i = 0; i < 7; i++) a7[c][i] = actions[1 + 7 * c + i];
            }
            int32_t info[K][4];
            for (int i = 0; i < K; i++) { const int c = reverse ? K - 1 - i : i; cells_begin<K>(s, x, e->P, L[c], actions[0], a7[c]); }
            for (int i = 0; i < K; i++) { const int c = reverse ? K - 1 - i : i; cells_act_run<K>(s, x, e->P, L[c], a7[c]); }
            for (int i = 0; i < K; i++) { const int c = reverse ? K - 1 - i : i; cells_finish<K, ArrayStateT<LONG>>(x, e->P, L[c], info[c]); }
            for (int i = 0; i < K; i++) {
                const int c = reverse ? K - 1 - i : i;
                cells_observe<K>(s, x, e->P, L[c], FloatSink{obs, e->P}, FloatSink{obs + 7 + 31 * c, e->P});
            }
            store_hot(s, L[0].h);
            for (int c = 0; c < K; c++) store_cell<K>(s, c, L[c].hc);
            for (int c = 0; c < K; c++) {
                const u32 mbits = cells_mask_word<K>(x, c);
                for (int i = 0; i < 32; i++) masks[32 * c + i] = (int8_t)((mbits >> i) & 1u);
                u32 nb[8];
                for (int i = 0; i < 8; i++) nb[i] = nibble_bytes(mbits >> (4 * i));
                if (memcmp(nb, masks + 32 * c, 32) != 0) return -1;  // nibble_bytes must be the same expansion
                for (int i = 0; i < 8; i++) {
                    const u32 v = x.ld16(2 * Xl<K>::LOCAL + 8 * c + i);
                    rewards[8 * c + i] = (8 * c + i < AG) ? (float)(L[c].g + AG * x_local10(v)) / (float)(10 * AG) : 0.0f;
                    if (results) results[8 * c + i] = (uint8_t)x_result(v);
                }
                if (L[c].flags != L[0].flags || L[c].g != L[0].g) return -2;  // every lane rebuilds the same shared scalars
            }
            flags[0] = L[0].flags & 0xff, flags[1] = (L[0].flags >> 8) & 0xff, flags[2] = (L[0].flags >> 16) & 0xff, flags[3] = 0;
            if (infos)
                for (int i = 0; i < 4; i++) infos[i] = info[0][i];
            done = 1;
        }
    })
    return done;
}

void hh_export(void* p, int cell, FjspCanonState* out) {
    HostEnv* e = (HostEnv*)p;
    if (e->shared) {
        export_canon_shared(e->words, e->P, cell, out);
        return;
    }
    export_canon(e->words, e->rq, e->P, e->cells, e->long_streams, cell, out);
}
void hh_export_orders(void* p, int first, int count, int32_t* out4, int32_t* order_base) {
    HostEnv* e = (HostEnv*)p;
    export_orders(e->words, e->P, e->cells, e->long_streams, first, count, out4);
    if (order_base) {
        FjspCanonState c;
        export_canon(e->words, e->rq, e->P, e->cells, e->long_streams, 0, &c, order_base);
    }
}
void hh_words(void* p, uint32_t* out) { memcpy(out, ((HostEnv*)p)->words, sizeof(u32) * state_words((HostEnv*)p)); }

void hh_philox_actions(uint64_t seed, uint64_t genv, uint64_t t, int cells, uint8_t* out) {
    int a[FJSP_ACT_DIM_K(FJSP_MAX_CELLS)] = {0};
    switch (cells) {
        case 1: philox_actions_k<1>(seed, genv, t, a); break;
        case 2: philox_actions_k<2>(seed, genv, t, a); break;
        case 3: philox_actions_k<3>(seed, genv, t, a); break;
        default: philox_actions_k<4>(seed, genv, t, a); break;
    }
    for (int i = 0; i < FJSP_ACT_DIM_K(cells); i++) out[i] = (uint8_t)a[i];
}
void hh_philox_actions_shared(uint64_t seed, uint64_t genv, uint64_t t, int agvs, uint8_t* out) {
    int a[FJSP_SHARED_ACT_DIM(FJSP_MAX_SHARED_AGVS)] = {0};
    switch (agvs) {
        case 2: philox_actions_shared<2>(seed, genv, t, a); break;
        case 3: philox_actions_shared<3>(seed, genv, t, a); break;
        default: philox_actions_shared<4>(seed, genv, t, a); break;
    }
    for (int i = 0; i < FJSP_SHARED_ACT_DIM(agvs); i++) out[i] = (uint8_t)a[i];
}
void hh_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out); }
}
