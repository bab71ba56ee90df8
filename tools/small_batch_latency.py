"""Step latency of small batches (BASELINE configs[1] = 4096 envs): one lockstep step replayed from a CUDA graph of 64
captured launches, for a few batch sizes.  Run once per setting of FJSP_PDL (0 = ordinary launches, 1 = programmatic
dependent launches; unset = the library's default) — the variable is read when the env is created.

    python tools/small_batch_latency.py [sizes...]      # JSON lines
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

sizes = [int(x) for x in sys.argv[1:]] or [32, 64, 2048, 4096, 8192, 9472, 18944]
dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in sizes:
    env = BatchedFJSPEnv(n, device=dev, seed=5, autoreset=True)
    env.reset()
    acts = [env.random_actions(t, out=torch.empty((n, 8), dtype=torch.uint8, device=dev)) for t in range(64)]
    for t in range(64):
        env.step(acts[t])
    torch.cuda.synchronize()
    e0.record()
    for t in range(1024):
        env.step(acts[t % 64])
    e1.record()
    torch.cuda.synchronize()
    stepwise_us = e0.elapsed_time(e1) / 1024 * 1e3
    gs = torch.cuda.Stream(device=dev)
    gs.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(gs):
        env.step(acts[0])
    torch.cuda.current_stream(dev).wait_stream(gs)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for t in range(64):
            env.step(acts[t])
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(32):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    graph_us = e0.elapsed_time(e1) / (64 * 32) * 1e3
    print(json.dumps({"envs": n, "tiles": (n + 63) // 64, "fjsp_pdl": os.environ.get("FJSP_PDL", "default"),
                      "stepwise_us": round(stepwise_us, 3), "graph_step_us": round(graph_us, 3),
                      "agent_steps_per_s_graph": n * 8 / (graph_us * 1e-6)}), flush=True)
    del graph, env, acts
