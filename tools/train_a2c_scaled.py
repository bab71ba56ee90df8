"""Short A2C training run on the scaled shop (K cells, the reference's 8 networks shared by the cells through
CellViewEnv): learning curve as JSON lines — mean team reward per env-step (all 1 + 7K agents) and the fraction of
episodes that end by completing every order.  Evidence that the trainer trains on the extension; not a benchmark."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv, CellViewEnv, abi
from multi_agent_rl_for_fjsp_b200.a2c_batched import BatchedA2C

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=2048)
ap.add_argument("--cells", type=int, default=4)
ap.add_argument("--rollout", type=int, default=32)
ap.add_argument("--updates", type=int, default=300)
ap.add_argument("--every", type=int, default=25)
ap.add_argument("--orders", type=int, default=25)
args = ap.parse_args()
cfg = abi.default_config()
cfg.num_cells = args.cells
env = BatchedFJSPEnv(args.envs, config=cfg, seed=123, num_orders=args.orders, autoreset=True)
tr = BatchedA2C(CellViewEnv(env), rollout_len=args.rollout, seed=7)
N, K, T = args.envs, args.cells, args.rollout
acc = torch.zeros(4, device=env.device)  # env-steps, team reward, episodes, terminated episodes
for u in range(1, args.updates + 1):
    tr.rollout()
    r = tr.rewards.view(T, N, K, 8)
    team = r[..., 1:].sum((-1, -2)) + r[:, :, 0, 0]              # the cells' 7K agents + the pickup station once
    f = tr.flags.view(T, N, K, 4)[:, :, 0]
    done = (f[..., 0:3] != 0).any(-1)
    acc += torch.stack([torch.tensor(float(T * N), device=env.device), team.sum(), done.sum().float(), (f[..., 0] != 0).sum().float()])
    tr.update()
    tr.obs[0].copy_(tr.obs[tr.T]), tr.masks[0].copy_(tr.masks[tr.T])
    if u % args.every == 0:
        a = acc.cpu().tolist()
        print(json.dumps({"updates": u, "frames": u * N * T, "cells": K, "mean_team_reward_per_env_step": a[1] / a[0],
                          "episodes": int(a[2]), "terminated_fraction": a[3] / max(1.0, a[2]),
                          "critic_loss": float(tr.stats["critic_loss"]) if tr.stats else None}), flush=True)
        acc.zero_()
