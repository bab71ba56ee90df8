// tests/host_harness/harness.cpp — TEST-ONLY host build of the device step function.
//
// Compiles multi_agent_rl_for_fjsp_b200/csrc/fjsp_core.h (the code the CUDA kernels run per env) with g++
// so the CPU-only test suite (`pytest -m "not gpu"`) can check the packed-state logic against the oracle
// on millions of steps without a GPU.  It is NOT part of the product: nothing in the package or in
// libfjsp_b200.so loads it, and the product has no CPU execution path.
#include <stdlib.h>

#include "../../multi_agent_rl_for_fjsp_b200/csrc/fjsp_host.h"

using namespace fjsp;

struct HostEnv {
    Params P;
    u32 words[W_TOTAL];
};

extern "C" {

void* hh_create(const FjspConfig* cfg) {
    FjspConfig c;
    if (cfg) c = *cfg; else default_config(&c);
    HostEnv* e = (HostEnv*)calloc(1, sizeof(HostEnv));
    if (make_params(c, &e->P) != nullptr) {
        free(e);
        return nullptr;
    }
    return e;
}
const char* hh_check_config(const FjspConfig* cfg) {
    Params P;
    const char* m = make_params(*cfg, &P);
    return m ? m : "";
}
void hh_destroy(void* p) { free(p); }

static void unpack_masks(const u32* mw, int8_t* masks) {
    for (int i = 0; i < FJSP_MASK_DIM; i++) masks[i] = (int8_t)((mw[i >> 2] >> ((i & 3) * 8)) & 0xff);
}

void hh_observe(void* p, float* obs, int8_t* masks) {
    HostEnv* e = (HostEnv*)p;
    ArrayState s{e->words};
    u32 mw[FJSP_MASK_DIM / 4];
    observe_env(s, e->P, obs, mw);
    unpack_masks(mw, masks);
}

void hh_reset(void* p, const FjspOrderRec* orders, int num_orders, uint64_t seed, uint64_t genv, uint32_t episode) {
    HostEnv* e = (HostEnv*)p;
    ArrayState s{e->words};
    reset_env(s, e->P, num_orders, orders, seed, genv, episode);
}

void hh_step(void* p, const uint8_t* actions, float* obs, int8_t* masks, float* rewards, uint8_t* flags, uint8_t* results,
             int32_t* infos) {
    HostEnv* e = (HostEnv*)p;
    ArrayState s{e->words};
    int a[8];
    for (int i = 0; i < 8; i++) a[i] = actions[i];
    StepOut out;
    out.obs = obs;
    step_env<true>(s, e->P, a, out);
    unpack_masks(out.mask, masks);
    for (int i = 0; i < 8; i++) rewards[i] = out.reward[i];
    flags[0] = out.flags & 0xff, flags[1] = (out.flags >> 8) & 0xff, flags[2] = (out.flags >> 16) & 0xff, flags[3] = 0;
    if (results)
        for (int i = 0; i < 8; i++) results[i] = (uint8_t)((out.results[i >> 2] >> ((i & 3) * 8)) & 0xff);
    if (infos)
        for (int i = 0; i < 4; i++) infos[i] = out.info[i];
}

void hh_export(void* p, FjspCanonState* out) {
    HostEnv* e = (HostEnv*)p;
    export_canon(e->words, e->P, out);
}
void hh_words(void* p, uint32_t* out) { memcpy(out, ((HostEnv*)p)->words, sizeof(u32) * W_TOTAL); }

void hh_philox_actions(uint64_t seed, uint64_t genv, uint64_t t, uint8_t* out) {
    int a[8];
    philox_actions(seed, genv, t, a);
    for (int i = 0; i < 8; i++) out[i] = (uint8_t)a[i];
}
void hh_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out); }
}
