#!/usr/bin/env python
"""bench.py — env agent-steps/sec of the batched FJSP environment step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E]

Workload (config.workload): BASELINE.json configs[3] — default layout (constants.py), 2^20 envs PER GPU,
30 Philox orders per env, uniform-random actions (Philox), auto-reset; one "step" = one lockstep env step of
every env = ONE launch of fjsp_step_kernel.  1 env-step = 8 agent-steps.  Weak scaling: per-GPU work is fixed,
ranks own disjoint global env ranges, there is no collective in the step.

value      device-timed (CUDA events on the launching stream), inputs resident in HBM: the K action buffers are
           generated before the timed region; every step reads a fresh 8 MiB action buffer and the 512 MiB state,
           so the working set exceeds the 126 MB L2 without a flush.
e2e        the same metric through the host-buffer C-ABI call fjsp_step_host: pinned host actions in, pinned host
           float32 obs / int8 masks / float32 rewards / u8 flags out, H2D + D2H copies inside the timed region.  The
           results cross PCIe as 32-byte wire rows (include/fjsp_b200.h) and are decoded into the caller's tensors
           by the library's host threads, pipelined with the copies; d2h_bytes_per_step counts the bytes that cross.
roofline   algorithmic bytes per launch = envs x (8 + 152 + 32 + 32 + 4 + 2 x 512) = envs x 1252 B,
           divided by the mean launch duration, against MEASURED_PEAKS.json hbm_gbs.  traffic = the ncu DRAM bytes per
           launch from profiles/ncu_traffic.json, only when the kernel's SASS hash there matches the running library.
strong_scaling  configs[3] as written: 2^20 envs in total over the N GPUs (the headline keeps 2^20 per GPU).
           e2e.undecoded_wire_rows_variant: the same pipeline delivering the rows as they are (fjsp_step_host_wire);
           e2e.decode_only_ms: the host decode alone.  Reported beside the headline, never instead of it.
cpu_baseline / --impl reference
           the reference is pure Python + SimPy and cannot travel to the GPU box, so the CPU arm is the C port of it
           (oracle/fjsp_oracle.c, kind "port") on all host threads, on a bounded sample of the same workload.
extra      (N = 1) BASELINE configs[1]: 4096 envs, stepwise / CUDA-graph / K-steps-per-launch figures.
scaled_shop (N = 1) BASELINE configs[4]: the 4-cell shop (29 agents) at 2^19 envs with its own roofline fraction.
a2c        BASELINE configs[2]: batched A2C frames/s, 4096 envs per GPU, rollout 32: grouped tcgen05 GEMMs in 3xTF32
           (fp32-level accuracy); single-pass TF32 and the torch / cuBLAS fp32 trainer beside it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "env_agent_steps_per_sec"
UNIT = "agent-steps/s"
BYTES_IO = 8 + 152 + 32 + 32 + 4
STATE_BYTES = 512
BYTES_PER_ENV_STEP = BYTES_IO + 2 * STATE_BYTES  # 1252
SAMPLE_ENVS, SAMPLE_INNER = 8192, 64  # the CPU arm's bounded sample: the same in cpu_baseline and --impl reference
SEED = 20261018
NUM_ORDERS = 30


def workload_name(envs):
    return ("configs[3]: default layout, %d envs per GPU (2^20 default), 30 Philox orders/env, Philox uniform-random "
            "actions, autoreset; 1 step = 1 lockstep env step of all envs" % envs)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region: NVML polled from a thread every ~2 ms
    (nvidia_ml_py), falling back to one `nvidia-smi` query loop when NVML is not importable."""

    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index=0):
        self.rows, self.gpu, self._stop, self.mode = [], gpu_index, False, None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.mode = None
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.mode = "smi"
            self.t = threading.Thread(target=self._read_smi, daemon=True)
            self.t.start()
        except Exception:
            self.mode = None

    def _poll_nvml(self):
        nv = self.nv
        while not self._stop:
            try:
                clk = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.time(), clk, {n for n, b in self.REASONS if bits & b}))
            except Exception:
                pass
            time.sleep(0.002)

    def _read_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                clk, self.max_mhz = float(f[0]), float(f[1])
            except Exception:
                continue
            names = ("hw_slowdown", "sw_thermal_slowdown", "hw_thermal_slowdown", "sw_power_cap")
            self.rows.append((time.time(), clk, {n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")}))

    def stop(self, t0, t1):
        self._stop = True
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        if self.mode == "smi":
            time.sleep(0.05)
            self.proc.terminate()
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        if not inside:  # region shorter than the sampling period: take the samples closest to it
            inside = sorted(self.rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))[:3]
        sm = sorted(r[1] for r in inside)
        reasons = set()
        for r in inside:
            reasons |= r[2]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(inside), "source": self.mode}


def ncu_traffic(envs):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the timed kernel, from the committed ncu capture —
    only if that capture was taken on THIS binary: profiles/ncu_traffic.json stores the SHA-256 of the kernel's SASS
    beside the figure (tools/sass_hash.py), and the running library's kernel is hashed the same way.  Otherwise null."""
    try:
        sys.path.insert(0, os.path.join(REPO, "tools"))
        from sass_hash import sass_sha256

        with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as f:
            rec = json.load(f)["fjsp_step_kernel<1,false,false>"]
        h, _ = sass_sha256()
        if h and h == rec["sass_sha256"] and int(rec["envs"]) == int(envs):
            return float(rec["dram_bytes_per_launch"]), rec.get("source")
    except Exception:
        pass
    return None, None


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_port_rate(sample_envs, steps_per_call, calls, threads, warm_calls=1):
    """Times the C port of the reference (oracle/fjsp_oracle.c) on `threads` host threads.  Returns
    (agent-steps/s, seconds, env-steps)."""
    from oracle.fjsp_oracle import OracleBatch

    b = OracleBatch(sample_envs, SEED, NUM_ORDERS)
    for _ in range(warm_calls):
        b.rollout(steps_per_call, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(calls):
        b.rollout(steps_per_call, nthreads=threads)
    dt = time.perf_counter() - t0
    env_steps = sample_envs * steps_per_call * calls
    return env_steps * 8 / dt, dt, env_steps


def cpu_sample_text(envs, inner, calls, threads, dt=None):
    return ("%d envs x %d consecutive env steps per call x %d calls, env-major%s (bounded sample of the 2^20-env workload: Philox "
            "orders/actions, autoreset), C port of the reference (oracle/fjsp_oracle.c), %d pthreads"
            % (envs, inner, calls, "" if dt is None else " (%.1f s)" % dt, threads))


def cpu_baseline_block(target_seconds=12.0):
    threads = os.cpu_count() or 1
    rate, dt, _ = cpu_port_rate(SAMPLE_ENVS, SAMPLE_INNER, 2, threads, warm_calls=1)  # calibrate
    calls = max(3, min(20000, int(target_seconds * rate / 8 / SAMPLE_ENVS / SAMPLE_INNER)))
    rate, dt, env_steps = cpu_port_rate(SAMPLE_ENVS, SAMPLE_INNER, calls, threads, warm_calls=0)
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": cpu_sample_text(SAMPLE_ENVS, SAMPLE_INNER, calls, threads, dt)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Python+SimPy and does not
    exist on the GPU box, so this is the C port on all host threads; each step = one lockstep step of a bounded
    sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    # one reference "step" = INNER consecutive steps of each of sample_envs envs, env-major, which is the CPU's best
    # mode (an env's 82 KB object graph stays in cache); agent-steps are counted the same way on both arms.
    sample_envs, inner = SAMPLE_ENVS, SAMPLE_INNER  # the same sample the GPU arm's cpu_baseline block times
    from oracle.fjsp_oracle import OracleBatch

    b = OracleBatch(sample_envs, SEED, NUM_ORDERS)
    for _ in range(args.warmup):
        b.rollout(inner, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b.rollout(inner, nthreads=threads)
    dt = time.perf_counter() - t0
    value = sample_envs * inner * 8 * args.steps / dt
    sample = "per step: " + cpu_sample_text(sample_envs, inner, 1, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
        "config": {"workload": workload_name(args.envs), "sample_envs": sample_envs, "inner_steps": inner},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the environment step exists only as sm_100a kernels (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from multi_agent_rl_for_fjsp_b200.dist import bind_to_gpu_numa

    numa_cpus = bind_to_gpu_numa(local_rank)  # before any pinned allocation: first touch on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    E, K, W = args.envs, args.steps, args.warmup
    # the ranks of one host share its cores for the host-buffer path's decode
    decode_threads = max(1, min(32, len(os.sched_getaffinity(0)) // world))
    env = BatchedFJSPEnv(E, device=dev, first_env=rank * E, seed=SEED, num_orders=NUM_ORDERS, autoreset=True,
                         decode_threads=decode_threads)
    env.reset()
    # inputs resident in HBM: one action buffer per step, generated by the Philox policy stand-in
    nbuf = min(W + K, args.max_action_buffers)
    acts = torch.empty((nbuf, E, 8), dtype=torch.uint8, device=dev)
    for t in range(nbuf):
        env.random_actions(t, out=acts[t])
    torch.cuda.synchronize()

    for t in range(W):
        env.step(acts[t % nbuf])
    launches0 = env.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    for t in range(W, W + K):
        env.step(acts[t % nbuf])
    ev1.record()
    barrier()
    wall1 = time.time()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = env.launch_count - launches0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_per_step = ms / K
    value = world * E * 8 * K / (ms * 1e-3)

    # ---- end to end through the host-buffer entry point
    Ke = max(3, min(K, args.e2e_steps))
    host_actions = [acts[i % nbuf].cpu().pin_memory() for i in range(min(4, nbuf))]  # inputs in pinned host memory
    for i in range(2):
        env.step_host(host_actions[i % len(host_actions)])
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        env.step_host(host_actions[i % len(host_actions)])
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * E * 8 * Ke / e2e_s
    # the same pipeline delivering the wire rows undecoded (for callers that consume the integers): reported, not the headline
    for i in range(2):
        env.step_host_wire(host_actions[i % len(host_actions)])
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        env.step_host_wire(host_actions[i % len(host_actions)])
    torch.cuda.synchronize()
    e2e_wire_s = max_over_ranks(time.perf_counter() - t0)
    # the host decode alone (no GPU work): the rows just delivered -> the float32 / int8 tensors, same thread count
    import ctypes as _C
    hb = env.host_buffers()
    _p = lambda t: _C.c_void_p(t.data_ptr())  # noqa: E731
    t0 = time.perf_counter()
    for i in range(5):
        env._L.fjsp_wire_decode(_C.byref(env.cfg), _p(hb["wire"]), E, _p(hb["obs"]), _p(hb["masks"]), _p(hb["rewards"]), _p(hb["flags"]),
                                decode_threads)
    decode_only_s = max_over_ranks((time.perf_counter() - t0) / 5)
    # the box's ceiling for the decode: streaming stores over the same pinned observation buffer, same thread count
    _secs = _C.c_double(0.0)
    _n = hb["obs"].numel() * 4
    best = None
    for i in range(4):
        env._L.fjsp_host_stream_write_probe(_p(hb["obs"]), _n, max(1, decode_threads), _C.byref(_secs))
        best = _secs.value if best is None else min(best, _secs.value)
    host_stream_gbs = _n / best / 1e9
    wire_row = 4 * env.dims["wire_words"]
    # the e2e path's own roofline: this box's pinned D2H bandwidth on the wire rows (64 B/env)
    hb = env.host_buffers()
    for _ in range(2):
        hb["obs"].copy_(env.obs, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        hb["obs"].copy_(env.obs, non_blocking=True)
    torch.cuda.synchronize()
    d2h_gbs = E * 152 * 5 / (time.perf_counter() - t0) / 1e9

    # ---- small-batch point (BASELINE configs[1]: 4096 envs, launch-latency bound), device-timed, reported as extra
    small = None
    if rank == 0 and world == 1 and not args.no_extras:
        es = BatchedFJSPEnv(4096, device=dev, seed=SEED, autoreset=True)
        es.reset()
        sa = [es.random_actions(t, out=torch.empty((4096, 8), dtype=torch.uint8, device=dev)) for t in range(64)]
        for t in range(50):
            es.step(sa[t % 64])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(2000):
            es.step(sa[t % 64])
        e1.record()
        torch.cuda.synchronize()
        sms = e0.elapsed_time(e1)
        st = es.rollout_random(50)
        torch.cuda.synchronize()
        e0.record()
        es.rollout_random(2000)
        e1.record()
        torch.cuda.synchronize()
        rms = e0.elapsed_time(e1)
        small = {"envs": 4096, "step_launch_us": sms / 2000 * 1e3, "agent_steps_per_s_stepwise": 4096 * 8 * 2000 / (sms * 1e-3),
                 "agent_steps_per_s_rollout_2000_steps_per_launch": 4096 * 8 * 2000 / (rms * 1e-3),
                 "note": "configs[1]; L2-resident, launch-latency bound; not the headline"}
        # the same 4096-env stepping replayed from a CUDA graph of 64 captured step launches (no per-launch host work)
        gs = torch.cuda.Stream(device=dev)
        gs.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(gs):
            es.step(sa[0])
        torch.cuda.current_stream(dev).wait_stream(gs)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for t in range(64):
                es.step(sa[t])
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(32):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        small["agent_steps_per_s_cuda_graph_64_steps"] = 4096 * 8 * 64 * 32 / (e0.elapsed_time(e1) * 1e-3)
        small["cuda_graph_step_us"] = e0.elapsed_time(e1) / (64 * 32) * 1e3
        # K-steps-per-launch variant on the full batch (state stays on-chip between steps; not the headline)
        e0.record()
        env.rollout_random(32)
        e1.record()
        torch.cuda.synchronize()
        small["rollout32_full_batch_agent_steps_per_s"] = E * 8 * 32 / (e0.elapsed_time(e1) * 1e-3)

    # ---- scaled shop (BASELINE configs[4]: K = 4 cells = 4 AGVs, 8 machines, 16 packaging stations, 29 agents), device-timed
    scaled = None
    if rank == 0 and world == 1 and not args.no_extras:
        from multi_agent_rl_for_fjsp_b200 import abi as _abi

        cfg4 = _abi.default_config()
        cfg4.num_cells = 4
        n4 = 1 << 19
        e4 = BatchedFJSPEnv(n4, config=cfg4, device=dev, seed=SEED, num_orders=32, autoreset=True)
        e4.reset()
        d4 = e4.dims
        a4 = [e4.random_actions(t, out=torch.empty((n4, d4["act"]), dtype=torch.uint8, device=dev)) for t in range(16)]
        for t in range(5):
            e4.step(a4[t % 16])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(50):
            e4.step(a4[t % 16])
        e1.record()
        torch.cuda.synchronize()
        ms4 = e0.elapsed_time(e1) / 50
        bytes4 = d4["act"] + 4 * d4["obs"] + d4["mask"] + 4 * d4["act"] + 4 + 2 * 4 * d4["state_words"]
        peak4, _ = measured_peak()
        scaled = {"workload": "configs[4] scaled shop: 4 cells (4 AGVs, 4 small + 4 big machines, 16 packaging stations, 29 agents), "
                              "%d envs, 32 Philox orders/env, Philox uniform-random actions, autoreset" % n4,
                  "envs": n4, "agents": d4["agents"], "ms_per_step": ms4, "agent_steps_per_s": n4 * d4["agents"] / (ms4 * 1e-3),
                  "state_bytes_per_env": 4 * d4["state_words"], "algorithmic_bytes_per_env_step": bytes4,
                  "achieved_gbs": n4 * bytes4 / (ms4 * 1e-3) / 1e9, "roofline_frac": n4 * bytes4 / (ms4 * 1e-3) / 1e9 / peak4,
                  "kernel": "fjsp_step_cells_kernel<4,false> (one thread per (env, cell))"}
        e0.record()
        e4.rollout_random(32)
        e1.record()
        torch.cuda.synchronize()
        scaled["rollout32_agent_steps_per_s"] = n4 * d4["agents"] * 32 / (e0.elapsed_time(e1) * 1e-3)  # K-steps-per-launch kernel
        del e4, a4
        # the same shop on the LONG order-stream layout (include/fjsp_b200.h): 8 orders at reset, Philox arrivals up to 200,
        # 500-step episodes — "larger grid, 4x AGVs and machines, long order streams" of configs[4]
        cfgl = _abi.default_config()
        cfgl.num_cells, cfgl.long_streams, cfgl.max_episode_steps = 4, 1, 500
        cfgl.arrival_prob_q16, cfgl.arrival_max_orders = 26214, 200   # 0.4 orders per step
        nl = 1 << 18
        el = BatchedFJSPEnv(nl, config=cfgl, device=dev, seed=SEED, num_orders=8, autoreset=True)
        el.reset()
        dl = el.dims
        al = [el.random_actions(t, out=torch.empty((nl, dl["act"]), dtype=torch.uint8, device=dev)) for t in range(16)]
        el.rollout_random(300)   # away from the empty initial state: arrivals, trays in the FIFOs, orders in process
        for t in range(5):
            el.step(al[t % 16])
        torch.cuda.synchronize()
        e0.record()
        for t in range(50):
            el.step(al[t % 16])
        e1.record()
        torch.cuda.synchronize()
        msl = e0.elapsed_time(e1) / 50
        bytesl = dl["act"] + 4 * dl["obs"] + dl["mask"] + 4 * dl["act"] + 4 + 2 * 4 * dl["state_words"]
        scaled["long_streams"] = {
            "workload": "4 cells, long order streams: 8 orders at reset + Philox arrivals (0.4 per step) up to 200, 500-step episodes, "
                        "%d envs, Philox uniform-random actions, autoreset" % nl,
            "envs": nl, "ms_per_step": msl, "agent_steps_per_s": nl * dl["agents"] / (msl * 1e-3),
            "state_bytes_per_env": 4 * dl["state_words"], "algorithmic_bytes_per_env_step": bytesl,
            "achieved_gbs": nl * bytesl / (msl * 1e-3) / 1e9, "roofline_frac": nl * bytesl / (msl * 1e-3) / 1e9 / peak4,
            "kernel": "fjsp_step_cells_kernel<4,false,true>",
            "note": "the ready FIFOs (4 KB per env in HBM) are touched in at most two words per env-step and are not counted"}
        del el, al
        # the SHARED FLOOR (include/fjsp_b200.h): 4 AGVs on one set of stations with station occupancy — "multi-AGV with
        # collision" of configs[4]; thread-per-env kernel on the reference shop's data path
        cfgs = _abi.default_config()
        cfgs.shared_agvs = 4
        ns = 1 << 20
        esf = BatchedFJSPEnv(ns, config=cfgs, device=dev, seed=SEED, num_orders=30, autoreset=True)
        esf.reset()
        dsf = esf.dims
        asf = [esf.random_actions(t, out=torch.empty((ns, dsf["act"]), dtype=torch.uint8, device=dev)) for t in range(12)]
        for t in range(30):
            esf.step(asf[t % 12])
        torch.cuda.synchronize()
        e0.record()
        for t in range(50):
            esf.step(asf[t % 12])
        e1.record()
        torch.cuda.synchronize()
        mss = e0.elapsed_time(e1) / 50
        bytess = dsf["act"] + 4 * dsf["obs"] + dsf["mask"] + 4 * dsf["act"] + 4 + 2 * 4 * dsf["state_words"]
        scaled["shared_floor"] = {
            "workload": "shared floor: 4 AGVs on one set of stations, one AGV per station position (occupancy resolved in agent "
                        "order), 11 agents, %d envs, 30 Philox orders/env, Philox uniform-random actions, autoreset" % ns,
            "envs": ns, "agents": dsf["agents"], "ms_per_step": mss, "agent_steps_per_s": ns * dsf["agents"] / (mss * 1e-3),
            "state_bytes_per_env": 4 * dsf["state_words"], "algorithmic_bytes_per_env_step": bytess,
            "achieved_gbs": ns * bytess / (mss * 1e-3) / 1e9, "roofline_frac": ns * bytess / (mss * 1e-3) / 1e9 / peak4,
            "kernel": "fjsp_shared_step_kernel<4>"}
        del esf, asf

    # ---- configs[3] as written: 2^20 envs in TOTAL, sharded over the N GPUs (strong scaling).  At N = 8 a rank's share
    #      (131,072 envs = 64 MiB of state) is L2-resident, so this is NOT an HBM-roofline figure; the headline keeps 2^20 per GPU.
    strong = None
    if not args.no_extras:
        total = 1 << 20
        es = total // world
        if world == 1:
            strong = {"total_envs": total, "envs_per_gpu": es, "ms_per_step": ms_per_step, "agent_steps_per_s": value,
                      "note": "N = 1: identical to the headline run"}
        else:
            sv = BatchedFJSPEnv(es, device=dev, first_env=rank * es, seed=SEED, num_orders=NUM_ORDERS, autoreset=True)
            sv.reset()
            sa = [sv.random_actions(t, out=torch.empty((es, 8), dtype=torch.uint8, device=dev)) for t in range(32)]
            for t in range(10):
                sv.step(sa[t % 32])
            barrier()
            ev0.record()
            for t in range(200):
                sv.step(sa[t % 32])
            ev1.record()
            barrier()
            sms = max_over_ranks(ev0.elapsed_time(ev1)) / 200
            strong = {"total_envs": total, "envs_per_gpu": es, "ms_per_step": sms, "agent_steps_per_s": total * 8 / (sms * 1e-3),
                      "state_mib_per_gpu": es * 512 / 2 ** 20,
                      "note": "strong scaling of configs[3]; a rank's state fits the 126 MB L2 from N = 8 on (L2-resident, not "
                              "an HBM figure)"}
            del sv, sa

    # ---- A2C frames/s (second half of BASELINE.json's metric; configs[2]): 4096 envs per GPU, rollout 32, fp32 GEMMs,
    #      CUDA-graph rollout, ONE flat NCCL all-reduce of the gradients per update when N > 1.  All ranks take part.
    a2c = None
    if not args.no_extras:
        from multi_agent_rl_for_fjsp_b200.a2c_batched import BatchedA2C

        def a2c_run(**kw):
            e = BatchedFJSPEnv(4096, device=dev, first_env=rank * 4096, seed=SEED + 1, num_orders=25, autoreset=True)
            t = BatchedA2C(e, rollout_len=32, seed=1, **kw)
            t.train(3)
            barrier()
            _, secs = t.train(20)
            secs = max_over_ranks(secs)
            n = t.net.num_parameters()
            del t, e
            return world * 4096 * 32 * 20 / secs, secs, n

        fps, secs, nparams = a2c_run()
        a2c = {"metric": "a2c_frames_per_sec", "value": fps, "unit": "frames/s",
               "envs_per_gpu": 4096, "rollout_len": 32, "updates": 20, "ms_per_update": secs / 20 * 1e3,
               "params": nparams, "impl": "umma: grouped tcgen05 GEMMs (TMEM accumulators), analytic loss gradients, rollout "
                                           "activations reused by the update, actors' logits layer fused into the second layer's "
                                           "epilogue, critic forward batched after the rollout, packed weight images fed by "
                                           "cp.async.bulk, per-network clipping + Adam as one call, CUDA-graph rollout and update",
               "gemm_precision": "3xTF32 (fp32 operands split into two TF32 terms, three products, fp32 accumulate: fp32-level)",
               "cuda_graph_rollout": True,
               "gradient_allreduce": "nccl, one flat 2.62 MB buffer" if world > 1 else "none (1 GPU)",
               "note": "1 frame = 1 env step consumed by training (rollout + update); reference CPU: 17-33 frames/s (BASELINE.md)"}
        # the same run with single-pass TF32 GEMMs, and with the torch / cuBLAS fp32 statement of the trainer (autograd):
        # reported beside the headline, never instead of it
        a2c["tf32_single_pass_frames_per_sec"] = a2c_run(gemm_passes=1)[0]
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            a2c["torch_cublas_fp32_frames_per_sec"] = a2c_run(impl="torch")[0]
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev_tf32

    if rank == 0:
        peak, peak_src = measured_peak()
        traffic, traffic_src = ncu_traffic(E)
        achieved = E * BYTES_PER_ENV_STEP / (ms_per_step * 1e-3) / 1e9  # GB/s per GPU
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(E), "envs_per_gpu": E, "num_orders": NUM_ORDERS,
                       "state_bytes_per_env": STATE_BYTES, "parallelism": "env-sharded x%d, no collective in the step" % world,
                       "l2": "no flush: each step streams the 512 MiB state and a fresh 8 MiB action buffer (> 126 MB L2)",
                       "action_buffers": nbuf},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "fjsp_step_kernel<1,false,false>",
                         "algorithmic_bytes_per_launch": E * BYTES_PER_ENV_STEP},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": E * 8, "d2h_bytes_per_step": E * wire_row,
                    "steps": Ke, "ms_per_step": e2e_s / Ke * 1e3,
                    "api": "fjsp_step_host (pinned host buffers; float32 obs/rewards, int8 masks, u8 flags delivered)",
                    "wire_row_bytes": wire_row, "decoded_bytes_per_step": E * (152 + 32 + 32 + 4),
                    "undecoded_wire_rows_variant": {"api": "fjsp_step_host_wire (same pipeline, rows delivered as they are)",
                                                    "value": world * E * 8 * Ke / e2e_wire_s, "ms_per_step": e2e_wire_s / Ke * 1e3},
                    "decode_threads": decode_threads, "decode_only_ms": decode_only_s * 1e3,
                    "decode_only_host_gbs": E * (wire_row + 220) / decode_only_s / 1e9,
                    "host_stream_write_gbs": host_stream_gbs,
                    "decode_only_note": "fjsp_wire_decode on the delivered rows, host only, threads spawned per call (the pipelined "
                                        "path uses the handle's persistent workers and overlaps the copies)", "pcie_d2h_gbs_measured": d2h_gbs, "host_cpus_bound": len(numa_cpus),
                    "pcie_bound_frac": (E * wire_row / (d2h_gbs * 1e9)) / (e2e_s / Ke)},
            "gpu_launches": launches, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block()
        if small:
            line["extra"] = small
        if scaled:
            line["scaled_shop"] = scaled
        if strong:
            line["strong_scaling"] = strong
        if a2c:
            line["a2c"] = a2c
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--max-action-buffers", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
