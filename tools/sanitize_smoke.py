"""Tiny run for compute-sanitizer: ragged batch, reset (explicit + Philox), steps with autoreset, rollout, host step,
fused A2C ops.  Keep it small: the sanitizer slows kernels ~50x."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv, abi

env = BatchedFJSPEnv(200, seed=5, with_infos=True)
env.reset()
for t in range(230):   # crosses the truncation/auto-reset boundary at step 201
    env.step(env.random_actions(t))
env.rollout_random(25)
env.step_host(env.random_actions(1).cpu().numpy())
orders = np.zeros((200, 32), np.uint32); orders[:, :3] = abi.order_rec(5, 1, 1)
env.reset(orders=orders, num_orders=3, env_mask=(np.arange(200) % 2).astype(np.uint8))
env.step(env.random_actions(2))
s = env.export_state(199)
L = abi.lib(); dev = env.device
logits = torch.randn(200, 32, device=dev); acts = torch.zeros(200, 8, dtype=torch.uint8, device=dev); lp = torch.zeros(200, 8, device=dev)
p = lambda x: C.c_void_p(x.data_ptr()); st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
assert L.fjsp_a2c_sample(p(logits), p(env.masks), p(acts), p(lp), 200, 0, 1, None, 0, st) == 0
T = 5
rew = torch.randn(T, 200, 8, device=dev); val = torch.randn(T + 1, 200, device=dev); fl = torch.zeros(T, 200, 4, dtype=torch.uint8, device=dev)
ret = torch.zeros_like(rew); adv = torch.zeros_like(rew)
assert L.fjsp_a2c_gae(p(rew), p(val), p(fl), p(ret), p(adv), T, 200, 0.99, 0.95, st) == 0
torch.cuda.synchronize()
print("sanitize smoke ok", int(s["current_step"]))
