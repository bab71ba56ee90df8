"""Short A2C training run on the CUDA env: per-episode learning curve as JSON lines (mean episode return summed over
the 8 agents, products packaged and orders completed per episode).  Evidence that the batched trainer trains; not a
benchmark (per-rollout bookkeeping below adds host work)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv
from multi_agent_rl_for_fjsp_b200.a2c_batched import BatchedA2C

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--rollout", type=int, default=32)
ap.add_argument("--updates", type=int, default=600)
ap.add_argument("--every", type=int, default=50)
ap.add_argument("--orders", type=int, default=25)
ap.add_argument("--tf32", action="store_true")
args = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
env = BatchedFJSPEnv(args.envs, seed=123, num_orders=args.orders, autoreset=True)
tr = BatchedA2C(env, rollout_len=args.rollout, seed=7)
N, dev = args.envs, env.device
ep_ret = torch.zeros(N, device=dev)
ep_glob = torch.zeros(N, device=dev)
win = torch.zeros(4, device=dev)  # episodes, sum return, sum global reward, terminated episodes
for u in range(1, args.updates + 1):
    tr.rollout()
    team = tr.rewards.sum(-1)                                   # [T,N]
    # shared part g/8 of every agent's reward: a packaging_blue_2 that never has a queue only ever receives g/8
    glob = tr.rewards.min(-1).values.clamp_min(-0.125) * 8       # lower bound of g from the least-rewarded agent
    done = (tr.flags[..., 0:3] != 0).any(-1)
    for t in range(args.rollout):
        ep_ret += team[t]
        ep_glob += (tr.rewards[t, :, 5] * 8 + 1.0).clamp_min(0)  # 100*orders + 10*products seen through blue_2's g/8
        d = done[t]
        if d.any():
            win += torch.stack([d.sum(), ep_ret[d].sum(), ep_glob[d].sum(), (tr.flags[t, :, 0] != 0).sum()]).float()
            ep_ret[d] = 0
            ep_glob[d] = 0
    tr.update()
    tr.obs[0].copy_(tr.obs[tr.T]), tr.masks[0].copy_(tr.masks[tr.T])
    tr.frames += 0
    if u % args.every == 0:
        w = win.cpu().tolist()
        if w[0] > 0:
            print(json.dumps({"updates": u, "frames": u * N * args.rollout, "episodes": int(w[0]),
                              "mean_episode_return": w[1] / w[0], "mean_episode_global_reward": w[2] / w[0],
                              "terminated_fraction": w[3] / w[0], "critic_loss": float(tr.stats["critic_loss"])
                              if tr.stats else None}))
        win.zero_()
