import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv
from multi_agent_rl_for_fjsp_b200.a2c_batched import BatchedA2C
tr = BatchedA2C(BatchedFJSPEnv(4096, seed=3, num_orders=25, autoreset=True), rollout_len=32, seed=2, use_cuda_graph=False)
tr.rollout()
eng = tr.engine
def timed(fn, n=50):
    g = torch.cuda.CUDAGraph()
    fn(); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * n)
l1, l2 = eng._fwd_actors[3]
print("layer1 us", timed(l1.launch), "layer2+heads us", timed(l2.launch), "both", timed(lambda: (l1.launch(), l2.launch())))
print("critic batch us", timed(eng.forward_critic, 10))
