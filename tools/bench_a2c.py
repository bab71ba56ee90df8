"""A2C frames/sec (BASELINE.json configs[2] / [4]): batched CTDE A2C on E envs per GPU, rollout length T.
1 frame = 1 env step consumed by training (rollout + update amortised).  Run plain for 1 GPU or under torchrun.
Prints one JSON line on rank 0."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv, dist as fdist
from multi_agent_rl_for_fjsp_b200.a2c_batched import BatchedA2C

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--rollout", type=int, default=32)
ap.add_argument("--updates", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--num-orders", type=int, default=25)
ap.add_argument("--cells", type=int, default=1, help="scaled shop: K cells per env, the 8 networks shared by the cells (CellViewEnv)")
ap.add_argument("--tf32", action="store_true", help="TF32 tensor-core GEMMs (fp32 storage/accumulate); default is full fp32")
ap.add_argument("--no-graph", action="store_true")
args = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
rank, local_rank, world = fdist.world_info()
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
fdist.init(device=dev)
cfg = None
if args.cells > 1:
    from multi_agent_rl_for_fjsp_b200 import CellViewEnv, abi
    cfg = abi.default_config()
    cfg.num_cells = args.cells
env = BatchedFJSPEnv(args.envs, config=cfg, device=dev, first_env=rank * args.envs, seed=11, num_orders=args.num_orders, autoreset=True)
tr = BatchedA2C(CellViewEnv(env) if args.cells > 1 else env, rollout_len=args.rollout, seed=1, use_cuda_graph=not args.no_graph)
tr.train(args.warmup)
l0 = env.launch_count
fdist.barrier(dev)
fps, secs = tr.train(args.updates)
secs = fdist.max_over_ranks(secs, dev)
frames = world * args.envs * args.rollout * args.updates
if rank == 0:
    print(json.dumps({"metric": "a2c_frames_per_sec", "value": frames / secs, "unit": "frames/s", "n_gpus": world,
                      "envs_per_gpu": args.envs, "cells": args.cells, "agents": 1 + 7 * args.cells, "rollout_len": args.rollout, "updates": args.updates,
                      "ms_per_update": secs / args.updates * 1e3, "env_step_launches": args.rollout * args.updates,
                      "mean_step_reward": tr.mean_reward(), "params": tr.net.num_parameters(), "gemm_precision": "tf32" if args.tf32 else "fp32", "cuda_graph_rollout": not args.no_graph,
                      "critic_loss": float(tr.stats["critic_loss"]), "update_graph": tr._ugraph is not None,
                      "update_graph_error": tr.update_graph_error}))
sys.stdout.flush()
if world > 1:
    # every rank is done (max_over_ranks above is a collective).  Tearing the NCCL communicator down while CUDA graphs that
    # captured its all-reduces are still alive was seen to hang at exit (2 ranks, scaled shop), so: drop the graphs, sync,
    # barrier, and leave without the communicator teardown.
    del tr
    import gc

    gc.collect()
    torch.cuda.synchronize(dev)
    dist.barrier()
    os._exit(0)
