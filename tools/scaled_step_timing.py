"""Step time of the scaled shop for K = 1..4 cells at an HBM-bound batch size (device-timed, random Philox actions):
agent-steps/s and the fraction of the measured HBM peak for the algorithmic bytes of one step."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv, abi

peak = 6551.4
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
state_mb = int(sys.argv[1]) if len(sys.argv) > 1 else 768
for k in (1, 2, 3, 4):
    cfg = abi.default_config()
    cfg.num_cells = k
    d = abi.dims(k)
    n = (state_mb << 20) // (4 * d["state_words"]) // 64 * 64
    env = BatchedFJSPEnv(n, config=cfg, seed=3, num_orders=32)
    env.reset()
    env.rollout_random(40)
    acts = [env.random_actions(100 + t, out=torch.empty((n, d["act"]), dtype=torch.uint8, device=env.device)) for t in range(8)]
    for t in range(3):
        env.step(acts[t])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for t in range(24):
        env.step(acts[t % 8])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 24
    nbytes = d["act"] + 4 * d["obs"] + d["mask"] + 4 * d["act"] + 4 + 8 * d["state_words"]
    print(json.dumps({"cells": k, "agents": d["agents"], "envs": n, "ms_per_step": ms, "agent_steps_per_s": n * d["agents"] / (ms * 1e-3),
                      "bytes_per_env_step": nbytes, "gbs": n * nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": n * nbytes / (ms * 1e-3) / 1e9 / peak}))
    del env, acts
