"""Where an A2C update cycle goes: CUDA-event split rollout / update (graph replay) and a per-kernel table of one
eager cycle from torch.profiler (kineto; nsys is not in the image).  Writes a JSON summary to stdout.

    python tools/profile_a2c.py [--envs 4096] [--rollout 32] [--impl torch|umma]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv
from multi_agent_rl_for_fjsp_b200.a2c_batched import BatchedA2C

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--rollout", type=int, default=32)
ap.add_argument("--impl", default=None)
ap.add_argument("--top", type=int, default=25)
args = ap.parse_args()
dev = torch.device("cuda", 0)
kw = {}
if args.impl:
    kw["impl"] = args.impl


def make(graph):
    env = BatchedFJSPEnv(args.envs, device=dev, seed=11, num_orders=25, autoreset=True)
    return BatchedA2C(env, rollout_len=args.rollout, seed=1, use_cuda_graph=graph, **kw)


def ev():
    return torch.cuda.Event(enable_timing=True)


out = {"envs": args.envs, "rollout_len": args.rollout, "impl": args.impl or "default"}
tr = make(True)
tr.train(4)
torch.cuda.synchronize()
ro, up = 0.0, 0.0
R = 10
for _ in range(R):
    a, b, c = ev(), ev(), ev()
    a.record()
    tr.rollout()
    b.record()
    tr.update()
    tr.obs[0].copy_(tr.obs[tr.T]), tr.masks[0].copy_(tr.masks[tr.T])
    c.record()
    torch.cuda.synchronize()
    ro += a.elapsed_time(b)
    up += b.elapsed_time(c)
out["graph_replay_ms"] = {"rollout": ro / R, "update": up / R, "frames_per_s": args.envs * args.rollout / ((ro + up) / R * 1e-3)}
del tr

tr = make(False)
tr.train(3)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.rollout()
    torch.cuda.synchronize()
    tr.update()
    torch.cuda.synchronize()
rows = []
total = 0.0
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0.0)
    if t <= 0:
        continue
    rows.append((t, e.count, e.key))
    total += t
rows.sort(reverse=True)
out["eager_cycle_device_ms"] = total / 1e3
out["kernels"] = [{"ms": t / 1e3, "share": t / total, "launches": n, "name": k[:140]} for t, n, k in rows[:args.top]]
out["kernel_count"] = sum(n for _, n, _ in rows)
print(json.dumps(out, indent=1))
