"""Step-kernel time vs batch size (CUDA-graph replay of 32 steps, so host launch overhead is excluded)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

for n in (4096, 9472, 16384, 47360, 65536, 131072, 262144, 524288, 1048576, 4194304):
    env = BatchedFJSPEnv(n, seed=1)
    env.reset()
    acts = [env.random_actions(t, out=torch.empty((n, 8), dtype=torch.uint8, device=env.device)) for t in range(32)]
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        env.step(acts[0])
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for t in range(32):
            env.step(acts[t])
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 320 * 1e3
    gbs = n * 1252 / (us * 1e-6) / 1e9
    print("envs %8d  tiles %6d  step %8.2f us  %7.1f GB/s algorithmic  (%.2f of 6551)  %.3e agent-steps/s" % (
        n, (n + 63) // 64, us, gbs, gbs / 6551.4, n * 8 / (us * 1e-6)))
    del env, g
