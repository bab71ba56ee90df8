"""Small driver for ncu: a few launches of each instantiation of the step kernel at HBM-bound sizes —
fjsp_step_kernel<1,false> (float tensors, the bench headline), <1,true> (wire rows, the host-buffer path) and
<4,false> (scaled shop) — after a warm-up rollout that takes the envs away from the all-idle initial state."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv, abi

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
n4 = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
reps = int(os.environ.get("PROFILE_REPS", "4"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(label, fn, envs, agents, nbytes):
    for _ in range(int(os.environ.get("PROFILE_WARM", "2"))):
        fn(0)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%s: %d envs, %.3f ms/launch, %.3e agent-steps/s, %.0f GB/s algorithmic" % (
        label, envs, ms, envs * agents / (ms * 1e-3), envs * nbytes / (ms * 1e-3) / 1e9))


env = BatchedFJSPEnv(n1, seed=3)
env.reset()
env.rollout_random(60)
acts = [env.random_actions(100 + t, out=torch.empty((n1, 8), dtype=torch.uint8, device=env.device)) for t in range(reps)]
timed("step<1,float>", lambda i: env.step(acts[i]), n1, 8, 1252)
wire = torch.zeros((n1, env.dims["wire_words"]), dtype=torch.int32, device=env.device)
timed("step<1,wire>", lambda i: env.step_wire(acts[i], wire), n1, 8, 8 + 64 + 1024)
del env, acts, wire

cfg = abi.default_config()
cfg.num_cells = 4
e4 = BatchedFJSPEnv(n4, config=cfg, seed=3, num_orders=32)
e4.reset()
e4.rollout_random(60)
d = e4.dims
a4 = [e4.random_actions(100 + t, out=torch.empty((n4, d["act"]), dtype=torch.uint8, device=e4.device)) for t in range(reps)]
b4 = d["act"] + 4 * d["obs"] + d["mask"] + 4 * d["act"] + 4 + 8 * d["state_words"]
timed("step<4,float>", lambda i: e4.step(a4[i]), n4, d["agents"], b4)
del e4, a4

cfgs = abi.default_config()
cfgs.shared_agvs = 4
es = BatchedFJSPEnv(n1, config=cfgs, seed=3, num_orders=30)
es.reset()
ds = es.dims
acs = [es.random_actions(100 + t, out=torch.empty((n1, ds["act"]), dtype=torch.uint8, device=es.device)) for t in range(reps)]
for t in range(60):
    es.step(es.random_actions(t))
bs = ds["act"] + 4 * ds["obs"] + ds["mask"] + 4 * ds["act"] + 4 + 8 * ds["state_words"]
timed("shared_step<4>", lambda i: es.step(acs[i]), n1, ds["agents"], bs)
