"""Per-call wall time of fjsp_step_host (host buffers in, decoded float tensors out) over many consecutive calls:
shows the spread the bench's mean hides.   python tools/e2e_probe.py [envs] [calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 60
env = BatchedFJSPEnv(n, seed=1)
env.reset()
acts = [env.random_actions(t).cpu().pin_memory() for t in range(4)]
ts = []
for i in range(calls):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    env.step_host(acts[i % 4])
    ts.append((time.perf_counter() - t0) * 1e3)
ts = np.array(ts)
print("fjsp_step_host, %d envs, %d calls: ms per call: first 5 %s | min %.3f median %.3f mean %.3f p90 %.3f max %.3f" % (
    n, calls, np.round(ts[:5], 2).tolist(), ts.min(), np.median(ts), ts.mean(), np.percentile(ts, 90), ts.max()))
print("per-call ms:", np.round(ts, 2).tolist())
