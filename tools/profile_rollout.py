"""Small driver for ncu: a few launches of the K-steps-per-launch rollout kernel (compute-bound view of step_env)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cells = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = None
if cells > 1:
    from multi_agent_rl_for_fjsp_b200 import abi
    cfg = abi.default_config()
    cfg.num_cells = cells
env = BatchedFJSPEnv(n, config=cfg, seed=3)
env.reset()
env.rollout_random(60)      # get away from the all-idle initial state
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    e0.record(); env.rollout_random(k); e1.record(); torch.cuda.synchronize()
    print("rollout %d envs x %d steps: %.3f ms -> %.3e agent-steps/s" % (n, k, e0.elapsed_time(e1), n * k * (1 + 7 * cells) / (e0.elapsed_time(e1) * 1e-3)))
