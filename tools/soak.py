"""Soak: many envs x many steps through the K-steps-per-launch kernel, Philox random policy, auto-reset; prints the
exact integer statistics (env-steps, episodes, orders completed, products packaged, faults).  The comparison of sampled
envs with the CPU restatement after such a run lives in tests/test_gpu_parity.py::test_gpu_soak_sampled_envs_match_restatement
(the oracle is test infrastructure: tools do not import it).  python tools/soak.py [envs] [steps] [cells]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv, abi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cells = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = abi.default_config()
cfg.num_cells = cells
seed = 99
env = BatchedFJSPEnv(n, config=cfg, seed=seed, num_orders=30, autoreset=True)
env.reset()
t0 = time.time()
done = 0
while done < steps:
    k = min(250, steps - done)
    st = env.rollout_random(k, t0=done)
    done += k
st = st.cpu().numpy()
secs = time.time() - t0
out = {"envs": n, "cells": cells, "steps": steps, "env_steps": int(st[0]), "episodes": int(st[1]), "orders_completed": int(st[2]),
       "products_packaged": int(st[3]), "faults": int(st[4]), "reward_units": int(np.int64(st[5])), "seconds": round(secs, 2)}
print(json.dumps(out))
