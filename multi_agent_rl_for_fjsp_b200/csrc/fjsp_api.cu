// fjsp_api.cu — the C ABI of libfjsp_b200.so (declared in include/fjsp_b200.h).
// Host-side glue only: argument checks, launches, copies.  There is no CPU implementation of the step here;
// every entry point that advances the simulation launches a kernel from fjsp_kernels.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <new>
#include <string>

#include "fjsp_host.h"
#include "fjsp_kernels.cuh"
#include "fjsp_wire.h"
#include "fjsp_a2c.cuh"
#include "fjsp_umma.cuh"
#include "fjsp_shared.cuh"

using namespace fjsp;

#define FJSP_HOST_MAX_CHUNKS 32
#define FJSP_HOST_MAX_STREAMS 4

struct FjspHandle {
    FjspConfig cfg;
    Params P;
    int device;
    int cells;               // K: 1 = the reference shop, 2..4 = scaled shop
    bool long_streams;       // long order streams: the second packed layout (fjsp_core.h)
    FjspOrderRec* d_otab;    // long layout: the handle's copy of the explicit order tables [num_envs][otab_stride], or null
    int otab_stride;
    u32* d_rq;               // long layout: ready FIFOs [num_envs][FJSP_LONG_READY_FIFO]
    int act, obs, mask;      // row widths of the I/O tensors (FJSP_*_DIM_K / FJSP_SHARED_*)
    int shared;              // shared floor: number of AGVs (0 = off)
    size_t tile_bytes;       // 64 envs x FJSP_STATE_WORDS_K words
    int64_t num_envs, first_env, num_tiles;
    u32* state;  // num_tiles * tile_bytes
    uint64_t seed;
    int num_orders;
    int64_t launches;
    // device staging for the host-buffer entry point (allocated on first use)
    uint8_t* d_actions;
    u32* d_wire;             // device wire rows  [num_envs][wire_words]
    u32* h_wire;             // pinned host copy of the same
    int wire_words;
    cudaStream_t hs[FJSP_HOST_MAX_STREAMS];  // the streams the host-buffer path deals its chunks round-robin on
    cudaEvent_t hev[FJSP_HOST_MAX_STREAMS], hin;
    int nstreams;            // 2 by default (FJSP_HOST_STREAMS = 1..4)
    cudaEvent_t cev[FJSP_HOST_MAX_CHUNKS];  // "chunk c has landed in h_wire"
    DecodePool* pool;        // host threads turning wire rows into the caller's float32 / int8 tensors
    int decode_threads;      // 0 = every CPU this process may run on (fjsp_set_decode_threads)
    bool staging_ready;      // host-buffer path: device / pinned staging, streams, events and decode workers exist
    int prefetch_tiles;      // cell-parallel kernel: L2 prefetch distance in tiles
    int prefetch_tiles_env;  // thread-per-env kernel: the same (0 = off)
    bool pdl;                // small batches: step launches carry the programmatic-serialization attribute
};

static thread_local std::string g_err;
static int fail(const std::string& m) {
    g_err = m;
    return 1;
}
static int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return 2;
}
#define CK(call)                                        \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(true) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        else if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

// The kernels are templated on the number of cells K; one switch per launch site.
#define DISPATCH_K(cells, ...)                                 \
    switch (cells) {                                           \
        case 1: { constexpr int K = 1; __VA_ARGS__; } break;   \
        case 2: { constexpr int K = 2; __VA_ARGS__; } break;   \
        case 3: { constexpr int K = 3; __VA_ARGS__; } break;   \
        default: { constexpr int K = 4; __VA_ARGS__; } break;  \
    }
// ... and on the layout (compact / long order streams)
// shared floor: on the number of AGVs
#define DISPATCH_A(agvs, ...)                                  \
    switch (agvs) {                                            \
        case 2: { constexpr int A = 2; __VA_ARGS__; } break;   \
        case 3: { constexpr int A = 3; __VA_ARGS__; } break;   \
        default: { constexpr int A = 4; __VA_ARGS__; } break;  \
    }
#define NOT_SHARED(h, what) \
    if ((h) && (h)->shared) return fail(what " is not available on the shared floor (FjspConfig.shared_agvs >= 2)");
#define DISPATCH_KL(h, ...)                                                  \
    if ((h)->long_streams) { constexpr bool LONG = true; DISPATCH_K((h)->cells, __VA_ARGS__) } \
    else { constexpr bool LONG = false; DISPATCH_K((h)->cells, __VA_ARGS__) }

template <class... Args>
static void launch_ex(void (*kern)(Args...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = pdl ? 1u : 0u;
    cudaLaunchKernelEx(&cfg, kern, args...);
}

template <int K, bool LONG>
static cudaError_t set_smem_attrs() {
    using G = Geo<K, LONG>;
    cudaError_t e = cudaFuncSetAttribute(fjsp_step_kernel<K, false, LONG>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::STEP_SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(fjsp_step_kernel<K, true, LONG>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::STEP_WIRE_SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(fjsp_rollout_kernel<K, LONG>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::ROLLOUT_SMEM_BYTES);
    if constexpr (K >= 2) {
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(fjsp_step_cells_kernel<K, false, LONG>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::CELLS_SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(fjsp_step_cells_kernel<K, true, LONG>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::CELLS_WIRE_SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(fjsp_rollout_cells_kernel<K, LONG>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::ROLLOUT_CELLS_SMEM_BYTES);
    }
    return e;
}

// One lockstep step of `tiles` tiles starting at A.tile_begin.  K = 1: one thread per env.  K >= 2: one thread per
// (env, cell) — unless FJSP_STEP_PER_ENV is set, which keeps the thread-per-env kernel for A/B measurements.
template <int K, bool WIRE, bool LONG>
static void launch_step(const FjspHandle* h, const StepArgs& A, unsigned tiles, cudaStream_t st, bool pdl = false) {
    using G = Geo<K, LONG>;
    static const bool per_env = getenv("FJSP_STEP_PER_ENV") != nullptr;
    if constexpr (K >= 2) {
        if (!per_env) {
            launch_ex(fjsp_step_cells_kernel<K, WIRE, LONG>, tiles, TILE * K, WIRE ? G::CELLS_WIRE_SMEM_BYTES : G::CELLS_SMEM_BYTES, st, pdl,
                      (const Params)h->P, (const StepArgs)A);
            return;
        }
    }
    launch_ex(fjsp_step_kernel<K, WIRE, LONG>, tiles, TILE, WIRE ? G::STEP_WIRE_SMEM_BYTES : G::STEP_SMEM_BYTES, st, pdl, (const Params)h->P,
              (const StepArgs)A);
}

template <int K, bool LONG>
static void launch_rollout(const FjspHandle* h, int steps, uint64_t seed, uint64_t t0, unsigned long long* stats, cudaStream_t st) {
    using G = Geo<K, LONG>;
    static const bool per_env = getenv("FJSP_STEP_PER_ENV") != nullptr;
    if constexpr (K >= 2) {
        if (!per_env) {
            fjsp_rollout_cells_kernel<K, LONG><<<(unsigned)h->num_tiles, TILE * K, G::ROLLOUT_CELLS_SMEM_BYTES, st>>>(
                h->P, h->state, h->num_envs, h->first_env, seed, t0, steps, h->num_orders, stats, h->d_rq);
            return;
        }
    }
    fjsp_rollout_kernel<K, LONG><<<(unsigned)h->num_tiles, TILE, G::ROLLOUT_SMEM_BYTES, st>>>(h->P, h->state, h->num_envs, h->first_env, seed,
                                                                                           t0, steps, h->num_orders, stats, h->d_rq);
}

// ---- tcgen05 grouped GEMM (fjsp_umma.cuh): the actor / critic layers of the batched A2C trainer ----
template <int AOP, int BOP>
static int launch_gemm(const void* probs, int nprob, int max_ctas, int passes, cudaStream_t st) {
    static thread_local int attr_dev = -1;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        CK(cudaFuncSetAttribute(umma::fjsp_gemm_kernel<AOP, BOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, umma::G_SMEM_BYTES));
        attr_dev = dev;
    }
    // programmatic dependent launch: the ~100 GEMM launches of a rollout run back to back, each a single wave
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)max_ctas, (unsigned)nprob), cfg.blockDim = dim3(umma::G_THREADS);
    cfg.dynamicSmemBytes = umma::G_SMEM_BYTES, cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, umma::fjsp_gemm_kernel<AOP, BOP>, static_cast<const umma::GemmProb*>(probs), passes));
    CK(cudaGetLastError());
    return 0;
}

extern "C" {

static void free_staging(FjspHandle* h);

const char* fjsp_last_error(void) { return g_err.c_str(); }
int fjsp_abi_version(void) { return FJSP_ABI_VERSION; }

int fjsp_default_config(FjspConfig* cfg) {
    if (!cfg) return fail("cfg is NULL");
    default_config(cfg);
    return 0;
}

int fjsp_create(const FjspConfig* cfg, int64_t num_envs, int64_t first_env, int device, FjspHandle** out) {
    if (!out) return fail("out is NULL");
    *out = nullptr;
    if (num_envs <= 0) return fail("num_envs must be positive");
    if (first_env < 0 || first_env + num_envs > 0xffffffffLL) return fail("global env index must fit 32 bits (Philox counter word)");
    FjspConfig c;
    if (cfg) c = *cfg; else default_config(&c);
    Params P;
    if (const char* m = make_params(c, &P)) return fail(m);
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail("no such CUDA device (this library has no CPU path)");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 9) return fail("device lacks bulk async copies (needs sm_90+; built for sm_100a)");
    DeviceGuard g(device);
    if (!g.ok) return fail("cudaSetDevice failed");
    FjspHandle* h = new (std::nothrow) FjspHandle();
    if (!h) return fail("out of host memory");
    h->cfg = c, h->P = P, h->device = device;
    h->num_envs = num_envs, h->first_env = first_env;
    h->cells = c.num_cells;
    h->long_streams = c.long_streams != 0;
    h->act = FJSP_ACT_DIM_K(h->cells), h->obs = FJSP_OBS_DIM_K(h->cells), h->mask = FJSP_MASK_DIM_K(h->cells);
    h->tile_bytes = (size_t)(h->long_streams ? FJSP_STATE_WORDS_LONG_K(h->cells) : FJSP_STATE_WORDS_K(h->cells)) * TILE * sizeof(u32);
    h->shared = c.shared_agvs >= 2 ? c.shared_agvs : 0;
    if (h->shared) {
        h->act = FJSP_SHARED_ACT_DIM(h->shared), h->obs = FJSP_SHARED_OBS_DIM(h->shared), h->mask = FJSP_SHARED_MASK_DIM(h->shared);
        h->tile_bytes = (size_t)FJSP_SHARED_STATE_WORDS * TILE * sizeof(u32);
    }
    h->wire_words = FJSP_WIRE_WORDS_K(h->cells);
    // measured on B200 (K = 4, 2^19 envs): distances of 32..148 tiles all give +13..15 %, 296 and more lose; SMs / 2 it is.
    // FJSP_PREFETCH_TILES overrides (0 switches the prefetch off)
    h->prefetch_tiles = prop.multiProcessorCount / 2 > 0 ? prop.multiProcessorCount / 2 : 1;
    if (const char* e = getenv("FJSP_PREFETCH_TILES")) h->prefetch_tiles = atoi(e);
    // thread-per-env kernel (6 CTAs per SM): one SM count ahead is worth +2 % at 2^20 envs (0.204 vs 0.208 ms), 3x that is
    // neutral, 6x loses 25 %
    h->prefetch_tiles_env = prop.multiProcessorCount;
    if (const char* e = getenv("FJSP_PREFETCH_TILES_ENV")) h->prefetch_tiles_env = atoi(e);
    h->num_tiles = (num_envs + TILE - 1) / TILE;
    // a batch that does not fill the GPU is launch-latency bound: its step kernels are launched as programmatic
    // dependents (the next launch's CTAs become resident while the current step runs).  FJSP_PDL=0/1 overrides.
    h->pdl = h->num_tiles <= 2 * (int64_t)prop.multiProcessorCount;
    if (const char* e = getenv("FJSP_PDL")) h->pdl = atoi(e) != 0;
    h->seed = 0, h->num_orders = 30, h->launches = 0;
    cudaError_t e = cudaMalloc(&h->state, (size_t)h->num_tiles * h->tile_bytes);
    if (e == cudaSuccess && h->long_streams) {
        e = cudaMalloc(&h->d_rq, (size_t)h->num_envs * FJSP_LONG_READY_FIFO * sizeof(u32));
        if (e == cudaSuccess) e = cudaMemset(h->d_rq, 0, (size_t)h->num_envs * FJSP_LONG_READY_FIFO * sizeof(u32));
        if (e != cudaSuccess) cudaFree(h->state);
    }
    if (e != cudaSuccess) {
        delete h;
        return cuda_fail(e, "cudaMalloc(state)");
    }
    DISPATCH_KL(h, e = (set_smem_attrs<K, LONG>()))
    if (e != cudaSuccess) {
        cudaFree(h->state);
        cudaFree(h->d_rq);
        delete h;
        return cuda_fail(e, "cudaFuncSetAttribute (is this an sm_100a device?)");
    }
    // all envs start as freshly reset, empty shops (num_orders = 0) until fjsp_reset is called
    if (h->shared) {
        DISPATCH_A(h->shared, (fjsp_shared_reset_kernel<A><<<(unsigned)h->num_tiles, TILE>>>(h->P, h->state, nullptr, nullptr, 0, 0ull, h->num_envs,
                                                                                            h->first_env, nullptr, nullptr)))
    } else {
    DISPATCH_KL(h, (fjsp_reset_kernel<K, LONG><<<(unsigned)h->num_tiles, TILE>>>(h->P, h->state, nullptr, nullptr, 0, 0, 0ull, h->num_envs,
                                                                                 h->first_env, nullptr, nullptr)))
    }
    h->launches++;
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(h->state);
        cudaFree(h->d_rq);
        delete h;
        return cuda_fail(e, "initial reset kernel");
    }
    *out = h;
    return 0;
}

int fjsp_destroy(FjspHandle* h) {
    if (!h) return 0;
    DeviceGuard g(h->device);
    cudaFree(h->state);
    cudaFree(h->d_otab);
    cudaFree(h->d_rq);
    free_staging(h);
    delete h;
    return 0;
}

int64_t fjsp_num_envs(const FjspHandle* h) { return h ? h->num_envs : 0; }
int fjsp_num_cells(const FjspHandle* h) { return h ? h->cells : 0; }
static int state_words(const FjspHandle* h) { return h->shared ? FJSP_SHARED_STATE_WORDS : h->long_streams ? FJSP_STATE_WORDS_LONG_K(h->cells) : FJSP_STATE_WORDS_K(h->cells); }
size_t fjsp_state_bytes(const FjspHandle* h) { return h ? (size_t)state_words(h) * 4 : 0; }
void* fjsp_state_ptr(FjspHandle* h) { return h ? h->state : nullptr; }
int64_t fjsp_launch_count(const FjspHandle* h) { return h ? h->launches : 0; }

int fjsp_reset(FjspHandle* h, const uint8_t* env_mask, uint64_t seed, const FjspOrderRec* orders, int num_orders, float* obs,
               int8_t* masks, void* stream) {
    if (!h) return fail("handle is NULL");
    if (!h->long_streams && (num_orders < 0 || num_orders > FJSP_MAX_ORDERS))
        return fail("num_orders must be in 0..32 in the compact layout (create the handle with FjspConfig.long_streams = 1 for more)");
    if (h->long_streams && (num_orders < 0 || num_orders > FJSP_LONG_MAX_ORDERS)) return fail("num_orders must be in 0..4095");
    if (h->long_streams && h->cfg.arrival_prob_q16 > 0 && num_orders > h->cfg.arrival_max_orders)
        return fail("num_orders (orders present at reset) must not exceed arrival_max_orders");
    if ((obs == nullptr) != (masks == nullptr)) return fail("obs and masks must both be given or both be NULL");
    DeviceGuard g(h->device);
    int stride = FJSP_MAX_ORDERS;
    if (h->long_streams) {
        // The ring fills as orders are popped, so explicit tables must outlive this call: the handle keeps a device copy
        // ([num_envs][num_orders]).  A masked reset may only replace rows of a table of the same width.
        // explicit tables: one record per order that can ever exist (with arrivals: arrival_max_orders of them)
        stride = h->cfg.arrival_prob_q16 > 0 && h->cfg.arrival_max_orders > num_orders ? h->cfg.arrival_max_orders : num_orders;
        if (orders) {
            if (env_mask && h->d_otab && h->otab_stride != stride) return fail("masked reset with explicit orders: the table width must equal the installed table's");
            if (!h->d_otab || h->otab_stride != stride) {
                CK(cudaStreamSynchronize((cudaStream_t)stream));
                cudaFree(h->d_otab);
                h->d_otab = nullptr;
                if (stride > 0) CK(cudaMalloc(&h->d_otab, (size_t)h->num_envs * stride * sizeof(FjspOrderRec)));
                h->otab_stride = stride;
            }
            if (stride > 0) {
                if (env_mask) {
                    // rows of the envs that are not reset keep their tables: copy row by row under the mask on the device
                    fjsp_copy_masked_rows_kernel<<<(unsigned)((h->num_envs * stride + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
                        h->d_otab, orders, env_mask, h->num_envs, stride);
                } else {
                    CK(cudaMemcpyAsync(h->d_otab, orders, (size_t)h->num_envs * stride * sizeof(FjspOrderRec), cudaMemcpyDeviceToDevice,
                                       (cudaStream_t)stream));
                }
                orders = h->d_otab;
            }
        } else if (!env_mask) {
            cudaFree(h->d_otab);   // back to the Philox order stream for every env
            h->d_otab = nullptr, h->otab_stride = 0;
        } else if (h->d_otab) {
            return fail("masked reset to the Philox order stream while explicit order tables are installed is not supported");
        }
    }
    // the handle-wide seed / num_orders are what auto-reset and the rollout kernels use for EVERY env: a masked reset
    // (some envs only) must not change them under the envs it does not touch
    if (!env_mask) h->seed = seed, h->num_orders = num_orders;
    if (h->shared) {
        DISPATCH_A(h->shared, (fjsp_shared_reset_kernel<A><<<(unsigned)h->num_tiles, TILE, 0, (cudaStream_t)stream>>>(
                                   h->P, h->state, env_mask, orders, num_orders, seed, h->num_envs, h->first_env, obs, masks)))
        h->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    DISPATCH_KL(h, (fjsp_reset_kernel<K, LONG><<<(unsigned)h->num_tiles, TILE, 0, (cudaStream_t)stream>>>(
                        h->P, h->state, env_mask, orders, stride, num_orders, seed, h->num_envs, h->first_env, obs, masks)))
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

int fjsp_step(FjspHandle* h, const uint8_t* actions, float* obs, int8_t* masks, float* rewards, uint8_t* flags, uint8_t* results,
              int32_t* infos, int autoreset, void* stream) {
    if (!h) return fail("handle is NULL");
    if (!actions || !obs || !masks || !rewards || !flags) return fail("actions/obs/masks/rewards/flags must be device pointers");
    if ((reinterpret_cast<uintptr_t>(actions) & 7) || (reinterpret_cast<uintptr_t>(masks) & 15) ||
        (reinterpret_cast<uintptr_t>(rewards) & 15) || (reinterpret_cast<uintptr_t>(flags) & 3) || (reinterpret_cast<uintptr_t>(obs) & 3) ||
        (results && (reinterpret_cast<uintptr_t>(results) & 7)) || (infos && (reinterpret_cast<uintptr_t>(infos) & 15)))
        return fail("buffer alignment: actions/results 8 B, masks/rewards/infos 16 B, flags/obs 4 B (obs 16 B for the bulk-store path)");
    DeviceGuard g(h->device);
    StepArgs A;
    A.state = h->state, A.actions = actions, A.obs = obs, A.masks = masks, A.rewards = rewards, A.flags = flags;
    A.results = results, A.infos = infos, A.num_envs = h->num_envs, A.first_env = h->first_env, A.seed = h->seed;
    A.num_orders = h->num_orders, A.autoreset = autoreset, A.tile_begin = 0, A.wire = nullptr;
    A.prefetch_tiles = h->prefetch_tiles, A.prefetch_tiles_env = h->prefetch_tiles_env;
    A.otab = h->d_otab, A.otab_stride = h->otab_stride, A.rq = h->d_rq;
    if (h->shared) {
        const StepArgs& Ar = A;
        DISPATCH_A(h->shared, launch_ex(fjsp_shared_step_kernel<A>, (unsigned)h->num_tiles, TILE, ShGeo<A>::STEP_SMEM_BYTES, (cudaStream_t)stream,
                                        h->pdl, (const Params)h->P, (const StepArgs)Ar))
    } else {
        DISPATCH_KL(h, (launch_step<K, false, LONG>(h, A, (unsigned)h->num_tiles, (cudaStream_t)stream, h->pdl)))
    }
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

static StepArgs wire_args(FjspHandle* h, const uint8_t* actions, u32* wire, uint8_t* results, int32_t* infos, int autoreset) {
    StepArgs A;
    A.state = h->state, A.actions = actions, A.obs = nullptr, A.masks = nullptr, A.rewards = nullptr, A.flags = nullptr;
    A.results = results, A.infos = infos, A.wire = wire, A.num_envs = h->num_envs, A.first_env = h->first_env, A.seed = h->seed;
    A.num_orders = h->num_orders, A.autoreset = autoreset, A.tile_begin = 0;
    A.prefetch_tiles = h->prefetch_tiles, A.prefetch_tiles_env = h->prefetch_tiles_env;
    A.otab = h->d_otab, A.otab_stride = h->otab_stride, A.rq = h->d_rq;
    return A;
}

int fjsp_step_wire(FjspHandle* h, const uint8_t* actions, uint32_t* wire, uint8_t* results, int32_t* infos, int autoreset, void* stream) {
    if (!h) return fail("handle is NULL");
    NOT_SHARED(h, "fjsp_step_wire")
    if (!actions || !wire) return fail("actions/wire must be device pointers");
    if ((reinterpret_cast<uintptr_t>(actions) & 7) || (reinterpret_cast<uintptr_t>(wire) & 15) ||
        (results && (reinterpret_cast<uintptr_t>(results) & 7)) || (infos && (reinterpret_cast<uintptr_t>(infos) & 15)))
        return fail("buffer alignment: actions/results 8 B, wire/infos 16 B");
    DeviceGuard g(h->device);
    const StepArgs A = wire_args(h, actions, wire, results, infos, autoreset);
    DISPATCH_KL(h, (launch_step<K, true, LONG>(h, A, (unsigned)h->num_tiles, (cudaStream_t)stream, h->pdl)))
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

size_t fjsp_wire_row_bytes(int num_cells) {
    return (num_cells >= 1 && num_cells <= FJSP_MAX_CELLS) ? (size_t)FJSP_WIRE_WORDS_K(num_cells) * 4 : 0;
}

int fjsp_wire_decode(const FjspConfig* cfg, const uint32_t* wire, int64_t n, float* obs, int8_t* masks, float* rewards, uint8_t* flags,
                     int threads) {
    if (!wire || n < 0) return fail("wire is NULL or n < 0");
    FjspConfig c;
    if (cfg) c = *cfg; else default_config(&c);
    Params P;
    if (const char* m = make_params(c, &P)) return fail(m);
    if (threads > 64) threads = 64;
    if (threads <= 1 || n < 2 * threads) {
        wire_decode(c.num_cells, P, wire, 0, n, obs, masks, rewards, flags);
        return 0;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&, t] { wire_decode(c.num_cells, P, wire, n * t / threads, n * (t + 1) / threads, obs, masks, rewards, flags); });
    for (auto& t : th) t.join();
    return 0;
}

static void free_staging(FjspHandle* h) {
    delete h->pool;
    h->pool = nullptr;
    cudaFree(h->d_actions), cudaFree(h->d_wire);
    h->d_actions = nullptr, h->d_wire = nullptr;
    if (h->h_wire) cudaFreeHost(h->h_wire);
    h->h_wire = nullptr;
    for (int i = 0; i < FJSP_HOST_MAX_STREAMS; i++) {
        if (h->hs[i]) cudaStreamDestroy(h->hs[i]);
        if (h->hev[i]) cudaEventDestroy(h->hev[i]);
        h->hs[i] = nullptr, h->hev[i] = nullptr;
    }
    for (int i = 0; i < FJSP_HOST_MAX_CHUNKS; i++) {
        if (h->cev[i]) cudaEventDestroy(h->cev[i]);
        h->cev[i] = nullptr;
    }
    if (h->hin) cudaEventDestroy(h->hin);
    h->hin = nullptr;
    h->staging_ready = false;
}

static int create_staging(FjspHandle* h) {
    const size_t n = (size_t)h->num_envs;
    CK(cudaMalloc(&h->d_actions, n * h->act));
    CK(cudaMalloc(&h->d_wire, n * h->wire_words * sizeof(u32)));
    CK(cudaMallocHost(&h->h_wire, n * h->wire_words * sizeof(u32)));
    h->nstreams = 2;
    if (const char* e = getenv("FJSP_HOST_STREAMS")) h->nstreams = atoi(e);
    if (h->nstreams < 1 || h->nstreams > FJSP_HOST_MAX_STREAMS) h->nstreams = 2;
    for (int i = 0; i < h->nstreams; i++) {
        CK(cudaStreamCreateWithFlags(&h->hs[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->hev[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < FJSP_HOST_MAX_CHUNKS; i++) CK(cudaEventCreateWithFlags(&h->cev[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->hin, cudaEventDisableTiming));
    // decode workers: the CPUs this process may run on (bench.py binds each rank to its GPU's NUMA node), minus the caller
    int workers = (h->decode_threads > 0 ? h->decode_threads : usable_cpus()) - 1;  // the calling thread decodes too
    if (const char* e = getenv("FJSP_DECODE_THREADS")) workers = atoi(e) - 1;
    if (workers > 31) workers = 31;
    if (n < 4096 || workers < 0) workers = 0;  // small batches decode on the calling thread
    h->pool = new (std::nothrow) DecodePool(workers);
    if (!h->pool) return fail("out of host memory");
    return 0;
}

// staging of the host-buffer path, created on first use; a failure half-way releases what was created
static int ensure_staging(FjspHandle* h) {
    if (h->staging_ready) return 0;
    const int rc = create_staging(h);
    if (rc != 0) {
        const std::string keep = g_err;
        free_staging(h);
        g_err = keep;
        return rc;
    }
    h->staging_ready = true;
    return 0;
}

// Host-buffer step.  The batch is cut into tile-aligned chunks that alternate between two internal streams: H2D of the
// chunk's actions, the step kernel writing WIRE ROWS (64 B per env instead of 220 B of float tensors), D2H of the rows
// into the handle's pinned staging.  As soon as a chunk has landed, the handle's host threads decode it into the
// caller's obs/masks/rewards/flags while later chunks are still computing / crossing PCIe; envs are independent, so
// chunks may run in any order.  Ordered after prior work on `stream`, and `stream` is ordered after it on return.
// wire_out != NULL: the rows land in the caller's host buffer and are not decoded (fjsp_step_host_wire)
static int step_host_impl(FjspHandle* h, const uint8_t* actions, u32* wire_out, float* obs, int8_t* masks, float* rewards, uint8_t* flags,
                          int autoreset, void* stream) {
    DeviceGuard g(h->device);
    if (int rc = ensure_staging(h)) return rc;
    cudaStream_t user = (cudaStream_t)stream;
    CK(cudaEventRecord(h->hin, user));
    for (int i = 0; i < h->nstreams; i++) CK(cudaStreamWaitEvent(h->hs[i], h->hin, 0));
    const int64_t tiles = h->num_tiles;
    int64_t nchunks = tiles >= 2048 ? 16 : tiles >= 64 ? 8 : (tiles >= 2 ? 2 : 1);
    static const int env_chunks = getenv("FJSP_HOST_CHUNKS") ? atoi(getenv("FJSP_HOST_CHUNKS")) : 0;
    static const int env_grain = getenv("FJSP_DECODE_GRAIN") ? atoi(getenv("FJSP_DECODE_GRAIN")) : 0;
    static const bool timing = getenv("FJSP_HOST_TIMING") != nullptr;
    if (env_chunks > 0 && env_chunks <= FJSP_HOST_MAX_CHUNKS && tiles >= env_chunks) nchunks = env_chunks;
    const int64_t per = (tiles + nchunks - 1) / nchunks;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_in = timing ? now() : 0.0;
    double t_ev[FJSP_HOST_MAX_CHUNKS + 2] = {0};
    StepArgs A = wire_args(h, h->d_actions, h->d_wire, nullptr, nullptr, autoreset);
    const int64_t na = h->act, ww = h->wire_words;
    int c = 0;
    for (int64_t t0 = 0; t0 < tiles; t0 += per, c++) {
        cudaStream_t st = h->hs[c % h->nstreams];
        const int64_t t1 = t0 + per < tiles ? t0 + per : tiles;
        const int64_t e0 = t0 * TILE, e1 = t1 * TILE < h->num_envs ? t1 * TILE : h->num_envs;
        const size_t n = (size_t)(e1 - e0);
        CK(cudaMemcpyAsync(h->d_actions + e0 * na, actions + e0 * na, n * na, cudaMemcpyHostToDevice, st));
        A.tile_begin = t0;
        DISPATCH_KL(h, (launch_step<K, true, LONG>(h, A, (unsigned)(t1 - t0), st)))
        h->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync((wire_out ? wire_out : h->h_wire) + e0 * ww, h->d_wire + e0 * ww, n * ww * sizeof(u32), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(h->cev[c], st));
    }
    for (int i = 0; i < h->nstreams; i++) {
        CK(cudaEventRecord(h->hev[i], h->hs[i]));
        CK(cudaStreamWaitEvent(user, h->hev[i], 0));
    }
    if (wire_out) {  // no decode: wait for the copies and return
        for (int i = 0; i < h->nstreams; i++) CK(cudaStreamSynchronize(h->hs[i]));
        return 0;
    }
    const int64_t grain = h->pool->size() > 0 ? (env_grain > 0 ? env_grain : 2048) : (int64_t)1 << 40;
    const double t_enq = timing ? now() : 0.0;
    c = 0;
    for (int64_t t0 = 0; t0 < tiles; t0 += per, c++) {
        const int64_t t1 = t0 + per < tiles ? t0 + per : tiles;
        const int64_t e0 = t0 * TILE, e1 = t1 * TILE < h->num_envs ? t1 * TILE : h->num_envs;
        cudaError_t e = cudaEventSynchronize(h->cev[c]);
        if (e != cudaSuccess) {
            h->pool->wait();
            return cuda_fail(e, "cudaEventSynchronize(chunk)");
        }
        if (timing) t_ev[c] = now();
        DecodePool::Job j{h->cells, &h->P, h->h_wire, e0, e1, obs, masks, rewards, flags};
        h->pool->submit(j, grain);
    }
    h->pool->wait();
    if (timing) {
        fprintf(stderr, "[fjsp_step_host] enqueue %.3f ms | chunk landed at", t_enq - t_in);
        for (int i = 0; i < c; i++) fprintf(stderr, " %.2f", t_ev[i] - t_in);
        fprintf(stderr, " | decoded at %.3f ms (%d workers, grain %lld)\n", now() - t_in, h->pool->size(), (long long)grain);
    }
    return 0;
}

int fjsp_step_host(FjspHandle* h, const uint8_t* actions, float* obs, int8_t* masks, float* rewards, uint8_t* flags, int autoreset,
                   void* stream) {
    if (!h) return fail("handle is NULL");
    NOT_SHARED(h, "fjsp_step_host")
    if (!actions || !obs || !masks || !rewards || !flags) return fail("host buffers must not be NULL");
    if (reinterpret_cast<uintptr_t>(masks) & 7) return fail("masks must be 8-byte aligned");
    return step_host_impl(h, actions, nullptr, obs, masks, rewards, flags, autoreset, stream);
}

int fjsp_step_host_wire(FjspHandle* h, const uint8_t* actions, uint32_t* wire, int autoreset, void* stream) {
    if (!h) return fail("handle is NULL");
    NOT_SHARED(h, "fjsp_step_host_wire")
    if (!actions || !wire) return fail("host buffers must not be NULL");
    return step_host_impl(h, actions, wire, nullptr, nullptr, nullptr, nullptr, autoreset, stream);
}

int fjsp_host_stream_write_probe(void* host_buf, size_t bytes, int threads, double* seconds) {
    if (!host_buf || !seconds || bytes == 0) return fail("host_buf / seconds is NULL or bytes == 0");
    if (threads < 1 || threads > 256) return fail("threads must be in 1..256");
    *seconds = host_stream_write_seconds(host_buf, bytes, threads);
    return 0;
}

int fjsp_set_decode_threads(FjspHandle* h, int threads) {
    if (!h) return fail("handle is NULL");
    if (threads < 0 || threads > 64) return fail("threads must be in 0..64 (0 = all CPUs of the process)");
    if (h->staging_ready) return fail("fjsp_set_decode_threads must be called before the first fjsp_step_host");
    h->decode_threads = threads;
    return 0;
}

int fjsp_random_actions(FjspHandle* h, uint64_t seed, uint64_t t, uint8_t* actions, void* stream) {
    if (!h) return fail("handle is NULL");
    if (!actions || (reinterpret_cast<uintptr_t>(actions) & 7)) return fail("actions must be an 8-byte aligned device pointer");
    DeviceGuard g(h->device);
    const int threads = 256;
    const unsigned blocks = (unsigned)((h->num_envs + threads - 1) / threads);
    if (h->shared) {
        DISPATCH_A(h->shared, launch_ex(fjsp_shared_random_actions_kernel<A>, blocks, threads, 0, (cudaStream_t)stream, h->pdl, actions,
                                        (int64_t)h->num_envs, (int64_t)h->first_env, (uint64_t)seed, (uint64_t)t))
        h->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    DISPATCH_K(h->cells, launch_ex(fjsp_random_actions_kernel<K>, blocks, threads, 0, (cudaStream_t)stream, h->pdl, actions,
                                   (int64_t)h->num_envs, (int64_t)h->first_env, (uint64_t)seed, (uint64_t)t))
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

int fjsp_rollout_random(FjspHandle* h, int steps, uint64_t seed, uint64_t t0, uint64_t* stats, void* stream) {
    if (!h) return fail("handle is NULL");
    NOT_SHARED(h, "fjsp_rollout_random")
    if (steps < 1) return fail("steps must be >= 1");
    if (!stats || (reinterpret_cast<uintptr_t>(stats) & 7)) return fail("stats must be an 8-byte aligned device pointer (8 x u64)");
    if (seed != h->seed) return fail("rollout seed must equal the seed of the last fjsp_reset (one Philox key per handle)");
    DeviceGuard g(h->device);
    if (h->d_otab) return fail("fjsp_rollout_random draws Philox orders at auto-reset: not available while explicit order tables are installed");
    DISPATCH_KL(h, (launch_rollout<K, LONG>(h, steps, seed, t0, reinterpret_cast<unsigned long long*>(stats), (cudaStream_t)stream)))
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

size_t fjsp_state_total_bytes(const FjspHandle* h) { return h ? (size_t)h->num_tiles * h->tile_bytes : 0; }

int fjsp_state_save(FjspHandle* h, void* dst_device, size_t bytes, void* stream) {
    if (!h || !dst_device) return fail("NULL argument");
    if (bytes != fjsp_state_total_bytes(h)) return fail("bytes must equal fjsp_state_total_bytes()");
    DeviceGuard g(h->device);
    CK(cudaMemcpyAsync(dst_device, h->state, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int fjsp_state_load(FjspHandle* h, const void* src_device, size_t bytes, void* stream) {
    if (!h || !src_device) return fail("NULL argument");
    if (bytes != fjsp_state_total_bytes(h)) return fail("bytes must equal fjsp_state_total_bytes()");
    DeviceGuard g(h->device);
    CK(cudaMemcpyAsync(h->state, src_device, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int fjsp_export_packed(FjspHandle* h, int64_t env, uint32_t* out_words) {
    if (!h || !out_words) return fail("NULL argument");
    if (env < 0 || env >= h->num_envs) return fail("env index out of range");
    DeviceGuard g(h->device);
    CK(cudaDeviceSynchronize());
    const int words = state_words(h);
    const u32* src = h->state + (env / TILE) * (int64_t)words * TILE + (env % TILE);
    // word w of this env sits TILE words after word w-1: a strided gather of `words` x 4 bytes
    CK(cudaMemcpy2D(out_words, sizeof(u32), src, TILE * sizeof(u32), sizeof(u32), (size_t)words, cudaMemcpyDeviceToHost));
    return 0;
}

int fjsp_export_state_cell(FjspHandle* h, int64_t env, int cell, FjspCanonState* out) {
    if (!out) return fail("out is NULL");
    if (h && (cell < 0 || cell >= (h->shared ? h->shared : h->cells))) return fail("cell (shared floor: AGV) index out of range");
    u32 words[FJSP_STATE_WORDS_LONG_K(FJSP_MAX_CELLS)];
    if (int rc = fjsp_export_packed(h, env, words)) return rc;
    if (h->shared) {
        export_canon_shared(words, h->P, cell, out);
        return 0;
    }
    std::vector<u32> rq;
    if (h->long_streams) {
        rq.resize(FJSP_LONG_READY_FIFO);
        DeviceGuard g(h->device);
        CK(cudaMemcpy(rq.data(), h->d_rq + env * FJSP_LONG_READY_FIFO, FJSP_LONG_READY_FIFO * sizeof(u32), cudaMemcpyDeviceToHost));
    }
    export_canon(words, h->long_streams ? rq.data() : nullptr, h->P, h->cells, h->long_streams, cell, out);
    return 0;
}

int fjsp_export_orders(FjspHandle* h, int64_t env, int first, int count, int32_t* out4, int32_t* order_base) {
    if (!out4 || count < 0) return fail("out4 is NULL or count < 0");
    u32 words[FJSP_STATE_WORDS_LONG_K(FJSP_MAX_CELLS)];
    if (int rc = fjsp_export_packed(h, env, words)) return rc;
    export_orders(words, h->P, h->cells, h->long_streams, first, count, out4);
    if (order_base) {
        FjspCanonState c;
        export_canon(words, nullptr, h->P, h->cells, h->long_streams, 0, &c, order_base);
    }
    return 0;
}

int fjsp_export_state(FjspHandle* h, int64_t env, FjspCanonState* out) { return fjsp_export_state_cell(h, env, 0, out); }

// ---- fused ops of the batched A2C trainer (stateless; run on the current device) ----
int fjsp_a2c_sample(const float* logits, const int8_t* masks, uint8_t* actions, float* logp, int64_t rows, int64_t first_row,
                    uint64_t seed, const uint64_t* counter, uint64_t t_off, void* stream) {
    if (!logits || !masks || !actions) return fail("logits/masks/actions must be device pointers");
    if ((reinterpret_cast<uintptr_t>(logits) & 15) || (reinterpret_cast<uintptr_t>(masks) & 15) ||
        (reinterpret_cast<uintptr_t>(actions) & 7) || (logp && (reinterpret_cast<uintptr_t>(logp) & 15)))
        return fail("alignment: logits/masks/logp 16 B, actions 8 B");
    if (rows <= 0) return 0;
    fjsp_policy_sample_kernel<<<(unsigned)((rows * 8 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        logits, masks, actions, logp, rows, first_row, seed, reinterpret_cast<const unsigned long long*>(counter), t_off);
    CK(cudaGetLastError());
    return 0;
}

int fjsp_cells_pack_actions(const uint8_t* view_actions, uint8_t* actions, int64_t num_envs, int num_cells, void* stream) {
    if (!view_actions || !actions) return fail("NULL argument");
    if (num_cells < 1 || num_cells > FJSP_MAX_CELLS || num_envs <= 0) return fail("num_cells must be in 1..4 and num_envs positive");
    const int act = FJSP_ACT_DIM_K(num_cells);
    const int64_t n = num_envs * act;
    fjsp_cells_pack_actions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(view_actions, actions, num_envs, num_cells, act);
    CK(cudaGetLastError());
    return 0;
}

int fjsp_cells_unpack_views(const float* obs, const int8_t* masks, const float* rewards, const uint8_t* flags, float* v_obs,
                            int8_t* v_masks, float* v_rewards, uint8_t* v_flags, int64_t num_envs, int num_cells, void* stream) {
    if (!obs || !masks || !rewards || !flags || !v_obs || !v_masks || !v_rewards || !v_flags) return fail("NULL argument");
    if (num_cells < 1 || num_cells > FJSP_MAX_CELLS || num_envs <= 0) return fail("num_cells must be in 1..4 and num_envs positive");
    const int64_t n = num_envs * num_cells * 82;
    fjsp_cells_unpack_views_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        obs, masks, rewards, flags, v_obs, v_masks, v_rewards, v_flags, num_envs, num_cells, FJSP_OBS_DIM_K(num_cells),
        FJSP_MASK_DIM_K(num_cells), FJSP_ACT_DIM_K(num_cells));
    CK(cudaGetLastError());
    return 0;
}

int fjsp_a2c_counter_add(uint64_t* counter, uint64_t inc, void* stream) {
    if (!counter) return fail("counter is NULL");
    fjsp_counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(counter), inc);
    CK(cudaGetLastError());
    return 0;
}

int fjsp_a2c_gae(const float* rewards, const float* values, const uint8_t* flags, float* returns, float* advantages, int T, int64_t N,
                 float gamma, float lamb, void* stream) {
    if (!rewards || !values || !flags || !returns || !advantages) return fail("NULL argument");
    if (T <= 0 || N <= 0) return fail("T and N must be positive");
    fjsp_gae_kernel<<<(unsigned)((N * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rewards, values, flags, returns, advantages, T, N,
                                                                                   gamma, lamb);
    CK(cudaGetLastError());
    return 0;
}

int fjsp_a2c_loss_grad(const float* logits, const int8_t* masks, const uint8_t* actions, const float* adv, const float* returns,
                       const float* values, const float* adv_mean, const float* adv_rstd, float entropy_coef, int64_t rows, float* dlogits,
                       float* dvalue, float* sums, void* stream) {
    if (!logits || !masks || !actions || !adv || !returns || !values || !adv_mean || !adv_rstd || !dlogits || !dvalue || !sums)
        return fail("NULL argument");
    if ((reinterpret_cast<uintptr_t>(logits) & 15) || (reinterpret_cast<uintptr_t>(masks) & 15) || (reinterpret_cast<uintptr_t>(actions) & 7) ||
        (reinterpret_cast<uintptr_t>(adv) & 15) || (reinterpret_cast<uintptr_t>(returns) & 15) || (reinterpret_cast<uintptr_t>(dlogits) & 15))
        return fail("alignment: logits/masks/adv/returns/dlogits 16 B, actions 8 B");
    if (rows <= 0) return fail("rows must be positive");
    fjsp_a2c_loss_grad_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        logits, masks, actions, adv, returns, values, adv_mean, adv_rstd, entropy_coef, 1.0f / (float)rows, dlogits, dvalue, sums, rows);
    CK(cudaGetLastError());
    return 0;
}

int fjsp_a2c_gemm(const FjspGemmProb* probs, int nprob, int max_ctas, int a_op, int b_op, int passes, void* stream) {
    if (!probs) return fail("probs is NULL");
    if (nprob < 1 || nprob > 65535 || max_ctas < 1) return fail("nprob must be in 1..65535 and max_ctas positive");
    if (passes != 1 && passes != 3) return fail("passes must be 1 (plain TF32) or 3 (3xTF32, fp32-level accuracy)");
    cudaStream_t st = (cudaStream_t)stream;
    if (a_op < 0 || a_op > 3 || b_op < 0 || b_op > 3) return fail("operand orientation out of range");
    const int combo = a_op * 4 + b_op;
    switch (combo) {
        case umma::OP_KC * 4 + umma::OP_MC: return launch_gemm<umma::OP_KC, umma::OP_MC>(probs, nprob, max_ctas, passes, st);
        case umma::OP_KCS * 4 + umma::OP_MC: return launch_gemm<umma::OP_KCS, umma::OP_MC>(probs, nprob, max_ctas, passes, st);
        case umma::OP_KC * 4 + umma::OP_KC: return launch_gemm<umma::OP_KC, umma::OP_KC>(probs, nprob, max_ctas, passes, st);
        case umma::OP_MC * 4 + umma::OP_MC: return launch_gemm<umma::OP_MC, umma::OP_MC>(probs, nprob, max_ctas, passes, st);
        case umma::OP_KCS * 4 + umma::OP_KCS: return launch_gemm<umma::OP_KCS, umma::OP_KCS>(probs, nprob, max_ctas, passes, st);
        case umma::OP_KC * 4 + umma::OP_PK: return launch_gemm<umma::OP_KC, umma::OP_PK>(probs, nprob, max_ctas, passes, st);
        case umma::OP_KCS * 4 + umma::OP_PK: return launch_gemm<umma::OP_KCS, umma::OP_PK>(probs, nprob, max_ctas, passes, st);
        default: return fail("unsupported operand orientation pair (supported: KC/MC, KCS/MC, KC/KC, MC/MC, KCS/KCS, KC/PK, KCS/PK)");
    }
}

int fjsp_a2c_wgrad_small(const FjspWgradJob* jobs, int njobs, int max_rows, int max_ny, void* stream) {
    if (!jobs) return fail("jobs is NULL");
    if (njobs < 1 || njobs > 65535 || max_rows < 1) return fail("njobs must be in 1..65535 and max_rows positive");
    if (max_ny < 1 || max_ny > 40) return fail("max_ny (widest Y of the jobs) must be in 1..40");
    static_assert(sizeof(FjspWgradJob) == sizeof(WgradJob), "FjspWgradJob mirrors WgradJob");
    const dim3 grid((unsigned)((max_rows + WG_SLAB - 1) / WG_SLAB), (unsigned)njobs);
    const WgradJob* j = reinterpret_cast<const WgradJob*>(jobs);
    static thread_local int attr_dev = -1;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        CK(cudaFuncSetAttribute(fjsp_a2c_wgrad_small_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES));
        CK(cudaFuncSetAttribute(fjsp_a2c_wgrad_small_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES));
        CK(cudaFuncSetAttribute(fjsp_a2c_wgrad_small_kernel<40>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES));
        attr_dev = dev;
    }
    if (max_ny <= 8) fjsp_a2c_wgrad_small_kernel<8><<<grid, 256, WG_SMEM_BYTES, (cudaStream_t)stream>>>(j);
    else if (max_ny <= 16) fjsp_a2c_wgrad_small_kernel<16><<<grid, 256, WG_SMEM_BYTES, (cudaStream_t)stream>>>(j);
    else fjsp_a2c_wgrad_small_kernel<40><<<grid, 256, WG_SMEM_BYTES, (cudaStream_t)stream>>>(j);
    CK(cudaGetLastError());
    return 0;
}

int fjsp_a2c_head_backward(const FjspHeadBwdJob* jobs, int njobs, int max_rows, void* stream) {
    if (!jobs) return fail("jobs is NULL");
    if (njobs < 1 || njobs > 65535 || max_rows < 1) return fail("njobs must be in 1..65535 and max_rows positive");
    static_assert(sizeof(FjspHeadBwdJob) == sizeof(HeadBwdJob), "FjspHeadBwdJob mirrors HeadBwdJob");
    const dim3 grid((unsigned)((max_rows + HB_SLAB - 1) / HB_SLAB), (unsigned)njobs);
    static thread_local int attr_dev = -1;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        CK(cudaFuncSetAttribute(fjsp_a2c_head_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HB_SMEM_BYTES));
        attr_dev = dev;
    }
    fjsp_a2c_head_backward_kernel<<<grid, 256, HB_SMEM_BYTES, (cudaStream_t)stream>>>(reinterpret_cast<const HeadBwdJob*>(jobs));
    CK(cudaGetLastError());
    return 0;
}

int fjsp_a2c_layer1(const FjspLayer1Job* jobs, int njobs, int max_rows, int max_k, void* stream) {
    if (!jobs) return fail("jobs is NULL");
    if (njobs < 1 || njobs > 65535 || max_rows < 1) return fail("njobs must be in 1..65535 and max_rows positive");
    if (max_k < 1 || max_k > L1_KMAX) return fail("max_k (largest k of the jobs) must be in 1..40");
    static_assert(sizeof(FjspLayer1Job) == sizeof(Layer1Job), "FjspLayer1Job mirrors Layer1Job");
    const dim3 grid((unsigned)((max_rows + L1_ROWS - 1) / L1_ROWS), (unsigned)njobs);
    const size_t smem = ((size_t)max_k * L1_NMAX + L1_ROWS * (L1_KMAX + 1)) * sizeof(float);   // <= 46,208 bytes
    fjsp_a2c_layer1_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const Layer1Job*>(jobs), max_k);
    CK(cudaGetLastError());
    return 0;
}

int fjsp_a2c_clip_adam(const FjspOptSeg* segs, int nseg, int max_elems, float* norms_sq, float max_norm, double beta1, double beta2,
                       double eps, void* stream) {
    if (!segs || !norms_sq) return fail("segs / norms_sq is NULL");
    if (nseg < 1 || nseg > 65535 || max_elems < 1) return fail("nseg must be in 1..65535 and max_elems positive");
    static_assert(sizeof(FjspOptSeg) == sizeof(OptSeg), "FjspOptSeg mirrors OptSeg");
    const OptSeg* s = reinterpret_cast<const OptSeg*>(segs);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned bx = (unsigned)std::min(64, (max_elems + 1023) / 1024);
    fjsp_a2c_gradnorm_kernel<<<dim3(bx, (unsigned)nseg), 256, 0, st>>>(s, norms_sq);
    fjsp_a2c_clip_adam_kernel<<<dim3(bx, (unsigned)nseg), 256, 0, st>>>(s, norms_sq, max_norm, beta1, beta2, eps);
    fjsp_a2c_opt_finish_kernel<<<(unsigned)((std::max(nseg, 16) + 127) / 128), 128, 0, st>>>(s, nseg, norms_sq);
    CK(cudaGetLastError());
    return 0;
}

int fjsp_a2c_gemm_pack(const FjspPackJob* jobs, int njobs, void* stream) {
    if (!jobs) return fail("jobs is NULL");
    if (njobs < 1 || njobs > 65535) return fail("njobs must be in 1..65535");
    static_assert(sizeof(FjspPackJob) == sizeof(umma::PackJob), "FjspPackJob mirrors umma::PackJob");
    umma::fjsp_gemm_pack_kernel<<<dim3(16, (unsigned)njobs), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const umma::PackJob*>(jobs));
    CK(cudaGetLastError());
    return 0;
}

}  // extern "C"
