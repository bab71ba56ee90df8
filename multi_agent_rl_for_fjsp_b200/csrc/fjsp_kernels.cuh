// fjsp_kernels.cuh — sm_100a kernels of the batched FJSP environment.
//
// HBM layout: the packed state is an ARRAY OF TILES.  One tile = 64 envs x 128 words, stored
// word-major  tile[w][lane]  (u32), i.e. 32 KB that are contiguous in HBM.  A CTA owns one tile per step:
//   1. one elected thread issues ONE bulk async copy (TMA engine, cp.async.bulk, SASS UBLKCP) of the tile's 26 KB of
//      dynamically indexed words (24..127: completion steps, orders, tray pool) into shared memory and arms an
//      mbarrier with the byte count; all threads meanwhile fetch their 24 hot words (coalesced 32-bit loads: a warp
//      reads 128 consecutive bytes per word) and the 8 action bytes of their env (coalesced 64-bit load);
//   2. every thread steps its own env: hot words in registers, the rest out of shared memory (thread t owns column t
//      of the [104][64] sub-tile, so all accesses of a warp hit 32 distinct banks regardless of the word index each
//      lane follows — the FIFO pointer chasing is conflict-free by construction);
//   3. observations are staged row-major in shared memory; masks / rewards / flags leave as 128-bit stores;
//   4. hot words go back with coalesced 32-bit stores; after a proxy fence + barrier the elected thread issues two bulk
//      async stores (the 26 KB sub-tile, the observation rows).
// Nothing but the step's own inputs and outputs crosses HBM: 512 B state in, 512 B state out, 8 + 152 + 32 + 32 + 4
// bytes of I/O per env-step.
#pragma once

#include <cuda_runtime.h>

#include "fjsp_core.h"

namespace fjsp {

constexpr int TILE = FJSP_TILE_ENVS;                       // 64 envs per tile
constexpr int TILE_WORDS = FJSP_STATE_WORDS * TILE;        // 8192 u32
constexpr int TILE_BYTES = TILE_WORDS * 4;                 // 32768
constexpr int HOT_BYTES = W_CSTEP * TILE * 4;              // 6144: words 0..23, through registers (coalesced LDG/STG)
constexpr int DYN_WORDS = (W_TOTAL - W_CSTEP) * TILE;      // words 24..127, through shared memory (one bulk copy)
constexpr int DYN_BYTES = DYN_WORDS * 4;                   // 26624
constexpr int OBS_ROW_BYTES = FJSP_OBS_DIM * 4;            // 152
constexpr int OBS_TILE_BYTES = OBS_ROW_BYTES * TILE;       // 9728
constexpr int STEP_SMEM_BYTES = DYN_BYTES + OBS_TILE_BYTES + 16;
constexpr int ROLLOUT_SMEM_BYTES = DYN_BYTES + 16;

// One env of a tile inside a kernel: the dynamically indexed words (24..127) live in shared memory, column `lane` of
// the [104][64] sub-tile; the 24 hot words are read/written straight from/to the HBM tile with compile-time indices
// (a warp touches 128 consecutive bytes per word: coalesced) and otherwise live in registers (struct Hot).
struct TileColumn {
    u32* dyn;  // &s_dyn[0][lane] - W_CSTEP * TILE, so dyn[w * TILE] is word w
    u32* hot;  // &g_tile[0][lane]
    __device__ __forceinline__ u32 ld(int w) const { return dyn[w * TILE]; }
    __device__ __forceinline__ void st(int w, u32 v) { dyn[w * TILE] = v; }
    __device__ __forceinline__ u32 ld_hot(int w) const { return hot[w * TILE]; }
    __device__ __forceinline__ void st_hot(int w, u32 v) { hot[w * TILE] = v; }
};
// Whole column addressed directly in HBM (reset / export paths, not hot).
struct GmemColumn {
    u32* base;
    __device__ __forceinline__ u32 ld(int w) const { return base[w * TILE]; }
    __device__ __forceinline__ void st(int w, u32 v) { base[w * TILE] = v; }
    __device__ __forceinline__ u32 ld_hot(int w) const { return base[w * TILE]; }
    __device__ __forceinline__ void st_hot(int w, u32 v) { base[w * TILE] = v; }
};

// ---- PTX wrappers (mbarrier + bulk async copy; see /opt/skills/guides/blackwell_cuda_programming.md) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Warp-cooperative auto-reset.  FJSP_MAX_ORDERS == 32 == warp size: for every lane whose episode just ended (ballot),
// the 32 lanes draw that env's 32 Philox orders in parallel (one order per lane, counter = (global env, episode, lane))
// and store them into the ending lane's shared-memory column; the ending lane itself re-initialises its scalars.
// A per-lane reset would make the whole warp wait for 32 sequential Philox calls whenever ANY of its envs ends.
// Must be called by all 32 lanes of the warp.  Returns true for lanes that were reset.
__device__ __forceinline__ bool warp_autoreset(TileColumn s, u32* s_dyn, int tid, bool do_reset, u32 cur_episode,
                                               int num_orders, uint64_t seed, uint64_t genv_lane0) {
    const unsigned need = __ballot_sync(0xffffffffu, do_reset);
    if (need == 0u) return false;
    const int lane = tid & 31, wbase = tid & ~31;
    u32 episode = 0u;
    if (do_reset) {
        episode = cur_episode + 1u;
        reset_env_base(s, num_orders, episode);  // hot words go to the HBM tile; the caller reloads its registers
    }
    unsigned rem = need;
    while (rem) {
        const int src = __ffs((int)rem) - 1;
        rem &= rem - 1u;
        const u32 ep = __shfl_sync(0xffffffffu, episode, src);
        const u32 ow = lane < num_orders ? philox_order(seed, genv_lane0 + (uint64_t)src, ep, lane) : 0u;
        s_dyn[(W_ORDER - W_CSTEP + lane) * TILE + wbase + src] = ow;
    }
    __syncwarp();
    return do_reset;
}

struct StepArgs {
    u32* state;              // tiles
    const uint8_t* actions;  // [N][8]
    float* obs;              // [N][38]
    int8_t* masks;           // [N][32]
    float* rewards;          // [N][8]
    uint8_t* flags;          // [N][4]
    uint8_t* results;        // [N][8] or null
    int32_t* infos;          // [N][4] or null
    int64_t num_envs, first_env;
    int64_t tile_begin;      // first tile of this launch (host-buffer path steps the batch in pipelined chunks)
    uint64_t seed;
    int32_t num_orders, autoreset;
};

// ---------------------------------------------------------------------------------------------
// Reset: FJSPSimulation.reset for the masked envs.  Thread per env, state addressed in place.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE) fjsp_reset_kernel(const __grid_constant__ Params P, u32* state, const uint8_t* env_mask,
                                                           const FjspOrderRec* orders, int num_orders, uint64_t seed,
                                                           int64_t num_envs, int64_t first_env, float* obs, int8_t* masks) {
    const int64_t env = (int64_t)blockIdx.x * TILE + threadIdx.x;
    const bool pad = env >= num_envs;
    if (!pad && env_mask && env_mask[env] == 0) return;
    GmemColumn s{state + (int64_t)blockIdx.x * TILE_WORDS + threadIdx.x};
    reset_env(s, P, pad ? 0 : num_orders, (pad || !orders) ? nullptr : orders + env * FJSP_MAX_ORDERS, seed,
              (uint64_t)(first_env + env), 0u);
    if (pad) return;
    if (obs && masks) {
        float o[FJSP_OBS_DIM];
        u32 mw[FJSP_MASK_DIM / 4];
        observe_env(s, P, o, mw);
#pragma unroll
        for (int i = 0; i < FJSP_OBS_DIM; i++) obs[env * FJSP_OBS_DIM + i] = o[i];
        uint4* m4 = reinterpret_cast<uint4*>(masks + env * FJSP_MASK_DIM);
        m4[0] = make_uint4(mw[0], mw[1], mw[2], mw[3]);
        m4[1] = make_uint4(mw[4], mw[5], mw[6], mw[7]);
    }
}

// ---------------------------------------------------------------------------------------------
// Step: one CTA = one tile of 64 envs; one launch = one lockstep step of all envs.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE) fjsp_step_kernel(const __grid_constant__ Params P, const StepArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u32* s_dyn = reinterpret_cast<u32*>(smem_raw);
    float* s_obs = reinterpret_cast<float*>(smem_raw + DYN_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + DYN_BYTES + OBS_TILE_BYTES);

    const int tid = threadIdx.x;
    const int64_t tile = A.tile_begin + blockIdx.x;
    const int64_t env = tile * TILE + tid;
    const bool valid = env < A.num_envs;
    u32* g_tile = A.state + tile * TILE_WORDS;

    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {  // ONE bulk async copy (TMA engine) for the 26 KB of dynamically indexed words of the tile
        mbar_expect_tx(bar, DYN_BYTES);
        bulk_g2s(s_dyn, g_tile + W_CSTEP * TILE, DYN_BYTES, bar);
    }
    // meanwhile: the 24 hot words (coalesced 32-bit loads, straight into registers) and the 8 action bytes (64-bit load)
    TileColumn s{s_dyn + tid - W_CSTEP * TILE, g_tile + tid};
    Hot h;
    load_hot(s, h);
    int a[8];
    {
        uint2 av = make_uint2(0u, 0u);
        if (valid) av = __ldg(reinterpret_cast<const uint2*>(A.actions) + env);
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = (int)((av.x >> (8 * i)) & 0xffu), a[4 + i] = (int)((av.y >> (8 * i)) & 0xffu);
    }
    mbar_wait(bar, 0);

    // padding lanes of a ragged last tile are inert: their words travel through unchanged
    StepOut out;
    out.obs = s_obs + tid * FJSP_OBS_DIM;
    out.flags = 0u;
    if (valid) step_env_hot<true>(s, P, h, a, out);
    const bool ended = valid && A.autoreset && (out.flags & 0x00ffffffu);
    if (warp_autoreset(s, s_dyn, tid, ended, h.episode, A.num_orders, A.seed, (uint64_t)(A.first_env + env - (tid & 31)))) {
        load_hot(s, h);
        observe(s, P, h, out.obs, out.mask);  // the observation returned with an ended episode is the new episode's first
        out.flags |= 1u << 24;
    }
    store_hot(s, h);
    if (valid) {
        uint4* m4 = reinterpret_cast<uint4*>(A.masks + env * FJSP_MASK_DIM);
        m4[0] = make_uint4(out.mask[0], out.mask[1], out.mask[2], out.mask[3]);
        m4[1] = make_uint4(out.mask[4], out.mask[5], out.mask[6], out.mask[7]);
        float4* r4 = reinterpret_cast<float4*>(A.rewards + env * FJSP_NUM_AGENTS);
        r4[0] = make_float4(out.reward[0], out.reward[1], out.reward[2], out.reward[3]);
        r4[1] = make_float4(out.reward[4], out.reward[5], out.reward[6], out.reward[7]);
        reinterpret_cast<u32*>(A.flags)[env] = out.flags;
        if (A.results) reinterpret_cast<uint2*>(A.results)[env] = make_uint2(out.results[0], out.results[1]);
        if (A.infos) reinterpret_cast<int4*>(A.infos)[env] = make_int4(out.info[0], out.info[1], out.info[2], out.info[3]);
    }
    // generic-proxy writes to shared memory must be visible to the async proxy before the bulk stores
    fence_async_smem();
    __syncthreads();
    const int64_t remaining = A.num_envs - tile * TILE;
    const int nvalid = remaining >= TILE ? TILE : (int)remaining;
    const bool obs_bulk = ((nvalid * OBS_ROW_BYTES) & 15) == 0 && ((reinterpret_cast<uintptr_t>(A.obs) & 15) == 0);
    if (tid == 0) {
        bulk_s2g(g_tile + W_CSTEP * TILE, s_dyn, DYN_BYTES);
        if (obs_bulk) bulk_s2g(A.obs + tile * TILE * FJSP_OBS_DIM, s_obs, (uint32_t)(nvalid * OBS_ROW_BYTES));
        bulk_commit();
    }
    if (!obs_bulk) {  // ragged last tile or unaligned caller buffer: cooperative coalesced 32-bit stores
        float* dst = A.obs + tile * TILE * FJSP_OBS_DIM;
        for (int i = tid; i < nvalid * FJSP_OBS_DIM; i += TILE) dst[i] = s_obs[i];
    }
    if (tid == 0) bulk_wait_read0();  // shared memory must stay allocated until the TMA engine has read it
}

// ---------------------------------------------------------------------------------------------
// Uniform-random policy stand-in (BASELINE.json configs 1/2/4): a_i ~ U{0..n_i-1} from Philox.
// ---------------------------------------------------------------------------------------------
__global__ void fjsp_random_actions_kernel(uint8_t* actions, int64_t num_envs, int64_t first_env, uint64_t seed, uint64_t t) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= num_envs) return;
    int a[8];
    philox_actions(seed, (uint64_t)(first_env + env), t, a);
    uint2 v;
    v.x = (u32)a[0] | ((u32)a[1] << 8) | ((u32)a[2] << 16) | ((u32)a[3] << 24);
    v.y = (u32)a[4] | ((u32)a[5] << 8) | ((u32)a[6] << 16) | ((u32)a[7] << 24);
    reinterpret_cast<uint2*>(actions)[env] = v;
}

// ---------------------------------------------------------------------------------------------
// Rollout: `steps` lockstep steps per launch with in-kernel Philox actions and auto-reset; the tile stays in
// shared memory between steps, so HBM sees the state once per `steps` steps.  Accumulates exact integer stats.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE) fjsp_rollout_kernel(const __grid_constant__ Params P, u32* state, int64_t num_envs,
                                                             int64_t first_env, uint64_t seed, uint64_t t0, int steps,
                                                             int num_orders, unsigned long long* stats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u32* s_dyn = reinterpret_cast<u32*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + DYN_BYTES);
    __shared__ unsigned long long s_acc[6];
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t env = tile * TILE + tid;
    const bool valid = env < num_envs;
    u32* g_tile = state + tile * TILE_WORDS;
    if (tid == 0) mbar_init(bar, 1);
    if (tid < 6) s_acc[tid] = 0ull;
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, DYN_BYTES);
        bulk_g2s(s_dyn, g_tile + W_CSTEP * TILE, DYN_BYTES, bar);
    }
    TileColumn s{s_dyn + tid - W_CSTEP * TILE, g_tile + tid};
    Hot h;  // hot words live in registers for the whole launch
    load_hot(s, h);
    mbar_wait(bar, 0);
    unsigned long long n_steps = 0, n_eps = 0, n_orders = 0, n_prod = 0, n_fault = 0;
    long long r40 = 0;
    const uint64_t genv = (uint64_t)(first_env + env);
    for (int k = 0; k < steps; k++) {  // uniform trip count: all lanes stay together for the cooperative reset
        bool ended = false;
        if (valid) {
            int a[8];
            philox_actions(seed, genv, t0 + (uint64_t)k, a);
            StepOut out;
            out.obs = nullptr;
            const int before_o = h.completed_orders, before_p = h.total_packaged;
            step_env_hot<false>(s, P, h, a, out);
            n_steps += 1;
            n_orders += (unsigned long long)(h.completed_orders - before_o);
            n_prod += (unsigned long long)(h.total_packaged - before_p);
            r40 += out.reward40;
            if (out.flags & 0x00ffffffu) {
                ended = true;
                n_eps += 1;
                n_fault += (out.flags >> 16) & 0xffu ? 1 : 0;
            }
        }
        if (warp_autoreset(s, s_dyn, tid, ended, h.episode, num_orders, seed, genv - (uint64_t)(tid & 31))) load_hot(s, h);
    }
    store_hot(s, h);
    // warp shuffle reduce, then one shared atomic per warp, one global atomic per CTA and counter
    unsigned long long v[6] = {n_steps, n_eps, n_orders, n_prod, n_fault, (unsigned long long)r40};
#pragma unroll
    for (int j = 0; j < 6; j++) {
        unsigned long long x = v[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if ((tid & 31) == 0) atomicAdd(&s_acc[j], x);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        bulk_s2g(g_tile + W_CSTEP * TILE, s_dyn, DYN_BYTES);
        bulk_commit();
    }
    if (tid < 6) atomicAdd(&stats[tid], s_acc[tid]);
    if (tid == 0) bulk_wait_read0();
}

}  // namespace fjsp
