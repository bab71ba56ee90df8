"""torch.profiler breakdown of one rollout + update of the batched A2C trainer (top CUDA kernels by time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv
from multi_agent_rl_for_fjsp_b200.a2c_batched import BatchedA2C

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
graph = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
if len(sys.argv) > 4 and sys.argv[4] == "tf32":
    torch.backends.cuda.matmul.allow_tf32 = True
env = BatchedFJSPEnv(envs, seed=11, num_orders=25)
tr = BatchedA2C(env, rollout_len=T, seed=1, use_cuda_graph=graph)
tr.train(3)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
torch.cuda.synchronize()
ev[0].record(); tr.rollout(); ev[1].record(); tr.update(); ev[2].record(); torch.cuda.synchronize()
print("envs %d T %d graph %s: rollout %.2f ms, update %.2f ms" % (envs, T, graph, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.rollout(); tr.update(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
