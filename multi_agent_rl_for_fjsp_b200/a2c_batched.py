"""Batched CTDE multi-agent A2C on the tensor API (BASELINE.json configs[2] and [4]).

The caller of the hot path, rebuilt for N-env lockstep rollouts.  Same algorithm as the reference's trainer
(/root/reference/a2c.py:24-116,168-252,647-731, networks.py:22-61, transition_memory.py:45-105), per-sample loops replaced
by batched tensor ops:

  * 8 actors (obs_i -> 256 -> 256 -> A_i, softmax) and one centralised critic (38 -> 256 -> 256 -> 128 -> 1),
    654,366 fp32 parameters, nn.Linear default init.  The six 3->256->256->3 actors run as one ``torch.bmm`` chain.
  * action selection: probs * action_mask, renormalise, uniform-over-valid fallback, Categorical sample (a2c.py:217-247);
  * returns and GAE per agent with the SHARED critic value, bootstrap V(s_T) when a rollout is cut, 0.0 when the episode
    ended — including time-limit truncation (a2c.py:325-332,357; transition_memory.py:83-105);
  * actor loss  -(norm_adv * logp).mean() - entropy_coef * H(unmasked probs + 1e-10), advantages normalised per agent
    with the unbiased std (a2c.py:694-731); critic loss = MSE of the value against every agent's returns (a2c.py:675-692);
  * clip_grad_norm 0.5 per network, one Adam per network (lr 3e-4 actors / 1e-3 critic).

Two implementations of the same algorithm: ``impl="umma"`` (default on CUDA) runs the nine networks forward and
backward as grouped tcgen05 GEMMs with TMEM accumulators (a2c_umma.py, csrc/fjsp_umma.cuh; 3xTF32 = fp32-level accuracy),
reuses the rollout's activations in the update and computes the loss gradients analytically; ``impl="torch"`` is the
plain torch / autograd statement (cuBLAS fp32) that the tensor-core path is tested against.
Data parallelism: every rank owns an env shard; ONE flat all-reduce (654,366 fp32 = 2.62 MB, NCCL over NVLink) of the
gradients per update, plus one tiny all-reduce of the advantage moments so normalisation is over the GLOBAL batch.
The env step is one kernel launch per step that writes observations straight into the rollout buffer (``step_into``).

Scaled shop (K cells): wrap the env in ``env.CellViewEnv`` — every (env, cell) becomes one row of the
reference's 8-agent layout, so this trainer runs unchanged and the 8 actors + critic are shared by the cells; the pickup
station acts through the row of cell 0 (elsewhere its mask allows action 0 only: constant log-probability, no gradient).
"""
from __future__ import annotations

import math
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .env import MASK_OFFSETS, N_ACTIONS, OBS_SLICES

HID = 256


def _linear_init(out_f, in_f, lead=(), device=None, gen=None):
    """nn.Linear's default init (kaiming_uniform(a=sqrt(5)) -> U(-1/sqrt(in), 1/sqrt(in)) for weight and bias)."""
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(*lead, in_f, out_f, device=device, generator=gen) * 2 - 1) * bound  # stored [in, out] for x @ w
    b = (torch.rand(*lead, 1, out_f, device=device, generator=gen) * 2 - 1) * bound
    return torch.nn.Parameter(w), torch.nn.Parameter(b)


class _StackedLinear(torch.autograd.Function):
    """y[l] = x[l] @ w[l] + b[l] for L stacked layers, with a split-K weight gradient.

    dW[l] = x[l]^T @ dy[l] contracts over the batch (K = B up to 10^6); cuBLAS runs the batched form with a few long
    K loops (9 ms of a 27 ms update at B = 131k).  Splitting the batch into S slices gives L*S independent GEMMs
    that fill the GPU, then one small reduction."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return torch.baddbmm(b, x, w)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        L, B, I = x.shape
        O = dy.shape[-1]
        dx = torch.bmm(dy, w.transpose(1, 2)) if ctx.needs_input_grad[0] else None
        S = 1
        while S < 32 and B % (2 * S) == 0 and B // (2 * S) >= 2048:
            S *= 2
        xs = x.reshape(L * S, B // S, I).transpose(1, 2)
        dw = torch.bmm(xs, dy.reshape(L * S, B // S, O)).reshape(L, S, I, O).sum(1)
        db = dy.sum(1, keepdim=True)
        return dx, dw, db


class ActorCritic(torch.nn.Module):
    """The reference's 8 ActorNetworks + CentralizedCriticNetwork as batched parameter groups."""

    def __init__(self, device=None, seed: int | None = None):
        super().__init__()
        gen = None
        if seed is not None:
            gen = torch.Generator(device=device)
            gen.manual_seed(seed)
        mk = lambda o, i, lead=(): _linear_init(o, i, lead, device, gen)  # noqa: E731
        # pickup_station (7 -> 3), agv (13 -> 8): networks.py:22-38
        self.ps = torch.nn.ParameterList([p for pair in (mk(HID, 7), mk(HID, HID), mk(3, HID)) for p in pair])
        self.agv = torch.nn.ParameterList([p for pair in (mk(HID, 13), mk(HID, HID), mk(8, HID)) for p in pair])
        # small/big machine + 4 packaging stations: six identical 3 -> 256 -> 256 -> 3 actors, one bmm chain
        self.six = torch.nn.ParameterList([p for pair in (mk(HID, 3, (6,)), mk(HID, HID, (6,)), mk(3, HID, (6,))) for p in pair])
        # centralised critic 38 -> 256 -> 256 -> 128 -> 1: networks.py:41-61
        self.critic = torch.nn.ParameterList(
            [p for pair in (mk(HID, 38), mk(HID, HID), mk(HID // 2, HID), mk(1, HID // 2)) for p in pair])

    @staticmethod
    def _mlp(x, params, last_act=None):
        n = len(params) // 2
        for i in range(n):
            w, b = params[2 * i], params[2 * i + 1]
            if w.dim() == 3:
                x = _StackedLinear.apply(x.contiguous(), w, b) if torch.is_grad_enabled() else torch.baddbmm(b, x, w)
            else:
                x = torch.addmm(b[0], x, w)
            if i < n - 1:
                x = torch.relu(x)
        return x

    def actor_logits(self, obs):
        """obs [B,38] -> (pickup [B,3], agv [B,8], six [6,B,3]) pre-softmax outputs of the 8 actors."""
        z_ps = self._mlp(obs[:, 0:7], self.ps)
        z_agv = self._mlp(obs[:, 7:20], self.agv)
        x6 = obs[:, 20:38].reshape(-1, 6, 3).transpose(0, 1)  # [6,B,3]
        return z_ps, z_agv, self._mlp(x6, self.six)

    def logits32(self, obs):
        """Pre-softmax logits in the [B,32] layout of the env's mask block (3 | 8 | 6x3 | 3 zero pad)."""
        z_ps, z_agv, z6 = self.actor_logits(obs)
        pad = torch.zeros(obs.shape[0], 3, device=obs.device, dtype=obs.dtype)
        return torch.cat([z_ps, z_agv, z6.transpose(0, 1).reshape(-1, 18), pad], dim=1).contiguous()

    def probs32(self, obs):
        """Softmax per agent (networks.py:33), same layout."""
        z_ps, z_agv, z6 = self.actor_logits(obs)
        pad = torch.zeros(obs.shape[0], 3, device=obs.device, dtype=obs.dtype)
        return torch.cat([torch.softmax(z_ps, -1), torch.softmax(z_agv, -1),
                          torch.softmax(z6, -1).transpose(0, 1).reshape(-1, 18), pad], dim=1)

    def value(self, obs):
        return self._mlp(obs, self.critic).squeeze(-1)

    def networks(self):
        """(name, parameter list, per-agent leading dim or None) for per-network clipping / optimisers."""
        return [("pickup_station", list(self.ps), None), ("agv", list(self.agv), None), ("six", list(self.six), 6),
                ("critic", list(self.critic), None)]

    def num_parameters(self):
        return sum(p.numel() for p in self.parameters())

    # ---- checkpoint compatibility with the reference (a2c.py:733-775): same dict layout, nn.Sequential key names
    def _groups(self):
        from .env import AGENT_IDS

        return [(AGENT_IDS[0], self.ps, None), (AGENT_IDS[1], self.agv, None)] + [(AGENT_IDS[2 + i], self.six, i) for i in range(6)]

    @staticmethod
    def _to_state_dict(params, idx):
        sd = {}
        for k in range(len(params) // 2):
            w, b = params[2 * k].detach(), params[2 * k + 1].detach()
            if idx is not None:
                w, b = w[idx], b[idx]
            sd["net.%d.weight" % (2 * k)] = w.t().contiguous().cpu().clone()  # nn.Linear stores [out, in]
            sd["net.%d.bias" % (2 * k)] = b.reshape(-1).cpu().clone()
        return sd

    @staticmethod
    def _from_state_dict(params, idx, sd):
        with torch.no_grad():
            for k in range(len(params) // 2):
                w = torch.as_tensor(sd["net.%d.weight" % (2 * k)]).t()
                b = torch.as_tensor(sd["net.%d.bias" % (2 * k)]).reshape(1, -1)
                (params[2 * k] if idx is None else params[2 * k][idx]).copy_(w)
                (params[2 * k + 1] if idx is None else params[2 * k + 1][idx]).copy_(b)

    def reference_checkpoint(self):
        """The dict ``MultiAgentA2C.save_model`` writes (a2c.py:745-752), with plain-int dims."""
        from .env import AGENT_IDS, N_ACTIONS, OBS_SLICES

        return {
            "actor_nets": {name: self._to_state_dict(list(p), i) for name, p, i in self._groups()},
            "critic_net": self._to_state_dict(list(self.critic), None),
            "obs_dims": {a: int(hi - lo) for a, (lo, hi) in zip(AGENT_IDS, OBS_SLICES)},
            "act_dims": {a: int(n) for a, n in zip(AGENT_IDS, N_ACTIONS)},
            "global_obs_dim": 38,
            "possible_agents": list(AGENT_IDS),
        }

    def save_reference_checkpoint(self, path):
        torch.save(self.reference_checkpoint(), path)

    def load_reference_checkpoint(self, path_or_dict):
        """Load a checkpoint in the reference's layout.  Files are unpickled with ``weights_only=True`` plus an allow-list
        for the NumPy scalar types the reference's ``act_dims`` contain (a2c.py:80: ``act_space.n`` is ``np.int64``);
        the file is third-party content, so nothing else is allowed to execute."""
        ckpt = path_or_dict
        if not isinstance(ckpt, dict):
            import numpy as np

            import numpy._core.multiarray as ma  # pickles written under NumPy 1.x name it numpy.core.multiarray

            allow = [np.dtype, np.int64, np.ndarray, type(np.dtype(np.int64)), ma.scalar, ma._reconstruct,
                     (ma.scalar, "numpy.core.multiarray.scalar"), (ma._reconstruct, "numpy.core.multiarray._reconstruct")]
            with torch.serialization.safe_globals(allow):
                ckpt = torch.load(path_or_dict, weights_only=True, map_location="cpu")
        for name, params, idx in self._groups():
            self._from_state_dict(list(params), idx, ckpt["actor_nets"][name])
        self._from_state_dict(list(self.critic), None, ckpt["critic_net"])
        return ckpt


_SEG = [(MASK_OFFSETS[i], N_ACTIONS[i]) for i in range(8)]


def masked_policy(probs32, masks):
    """probs * mask, renormalised per agent; uniform over valid actions when the masked mass is 0 (a2c.py:217-231).
    Returns q [B,32] with each agent's segment summing to 1."""
    m = masks.to(probs32.dtype)
    pm = probs32 * m
    out = torch.zeros_like(pm)
    for off, n in _SEG:
        seg, mseg = pm[:, off:off + n], m[:, off:off + n]
        s = seg.sum(-1, keepdim=True)
        uni = mseg / mseg.sum(-1, keepdim=True).clamp_min(1.0)
        out[:, off:off + n] = torch.where(s > 0, seg / s.clamp_min(1e-38), uni)
    return out


def sample_actions(q, generator=None):
    """Categorical sample per agent from q [B,32] -> uint8 [B,8] (inverse CDF on one uniform per agent)."""
    B = q.shape[0]
    u = torch.rand(B, 8, device=q.device, generator=generator)
    acts = torch.empty(B, 8, dtype=torch.uint8, device=q.device)
    for i, (off, n) in enumerate(_SEG):
        cdf = q[:, off:off + n].cumsum(-1)
        a = (u[:, i:i + 1] * cdf[:, -1:] >= cdf).sum(-1).clamp_max(n - 1)
        # never pick a zero-probability action because of rounding at the end of the cdf
        valid = q[:, off:off + n] > 0
        a = torch.where(valid.gather(1, a[:, None])[:, 0], a, valid.float().argmax(-1))
        acts[:, i] = a.to(torch.uint8)
    return acts


_SEG_OFF = {}


def _seg_offsets(device):
    """Cached [8] tensor of the agents' segment offsets (built once per device: no H2D copy inside graph capture)."""
    key = str(device)
    if key not in _SEG_OFF:
        _SEG_OFF[key] = torch.tensor([o for o, _ in _SEG], device=device)
    return _SEG_OFF[key]


def log_prob_of(q, actions):
    """log q[a] per agent, with Categorical's probability clamp (torch.distributions: eps = finfo.eps)."""
    eps = torch.finfo(q.dtype).eps
    idx = actions.long() + _seg_offsets(q.device)
    return torch.log(q.gather(1, idx).clamp(eps, 1 - eps))


def entropy_unmasked(probs32):
    """-sum p log(p + 1e-10) per agent over the UNMASKED probabilities (a2c.py:694-710) -> [B,8]."""
    e = -(probs32 * torch.log(probs32 + 1e-10))
    return torch.stack([e[:, off:off + n].sum(-1) for off, n in _SEG], dim=1)


def gae_and_returns(rewards, values, dones, gamma, lamb):
    """rewards [T,N,8], values [T+1,N] (values[T] = bootstrap), dones [T,N] bool -> (returns, advantages) [T,N,8].
    Reverse scan of transition_memory.py:83-105; an episode end zeroes the bootstrap and restarts the accumulators."""
    T = rewards.shape[0]
    ret = values[T].unsqueeze(-1).expand_as(rewards[0]).clone()
    nxt = values[T].unsqueeze(-1).expand_as(rewards[0]).clone()
    gae = torch.zeros_like(rewards[0])
    returns, advs = torch.empty_like(rewards), torch.empty_like(rewards)
    for t in range(T - 1, -1, -1):
        nd = (~dones[t]).to(rewards.dtype).unsqueeze(-1)
        v = values[t].unsqueeze(-1)
        ret = rewards[t] + gamma * ret * nd
        td = rewards[t] + gamma * nxt * nd - v
        gae = td + gamma * lamb * gae * nd
        returns[t], advs[t] = ret, gae
        nxt = v.expand_as(gae)
    return returns, advs


def _ptr(t):
    import ctypes

    return None if t is None else ctypes.c_void_p(t.data_ptr())


class BatchedA2C:
    def __init__(self, env, rollout_len=32, gamma=0.99, lamb=0.95, lr_actor=3e-4, lr_critic=1e-3, entropy_coef=0.01,
                 max_grad_norm=0.5, seed=0, global_adv_norm=True, use_cuda_graph=True, fused_ops=True, impl=None, gemm_passes=3,
                 fused_optimizer=True):
        """impl: "umma" (CUDA default) — forward and backward of the 9 networks as grouped tcgen05 GEMMs (a2c_umma.py:
        3xTF32 = fp32-level accuracy; gemm_passes=1: plain TF32), the rollout's activations reused by the update, analytic
        loss gradients, no autograd; "torch" — the same algorithm with torch ops and autograd (the fp32 reference the umma
        path is tested against; the only path off-GPU).  fused_optimizer (umma only): per-network clipping + Adam as one
        C-ABI call on the optimizer's own state tensors (a2c_umma.ClipAdam) instead of torch's ~30 launches."""
        self.env, self.T = env, int(rollout_len)
        self.seed = int(seed)
        self.gamma, self.lamb, self.entropy_coef, self.max_grad_norm = gamma, lamb, entropy_coef, max_grad_norm
        self.device = env.device
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.global_adv_norm = global_adv_norm
        self.net = ActorCritic(device=self.device, seed=seed)  # same seed on every rank -> identical replicas
        groups = self.net.networks()
        fused = self.device.type == "cuda"
        self.opt = torch.optim.Adam(
            [{"params": g[1], "lr": lr_critic if g[0] == "critic" else lr_actor} for g in groups], fused=fused,
            capturable=fused)
        N, T, dev = env.num_envs, self.T, self.device
        self.obs = torch.zeros(T + 1, N, 38, device=dev)
        self.masks = torch.zeros(T + 1, N, 32, dtype=torch.int8, device=dev)
        self.actions = torch.zeros(T, N, 8, dtype=torch.uint8, device=dev)
        self.rewards = torch.zeros(T, N, 8, device=dev)
        self.flags = torch.zeros(T, N, 4, dtype=torch.uint8, device=dev)
        self.values = torch.zeros(T + 1, N, device=dev)
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed * 1000003 + (dist.get_rank() if self.world > 1 else 0))
        # fused CUDA ops (libfjsp_b200.so: fjsp_a2c_sample / fjsp_a2c_gae) and CUDA-graph replay of the rollout
        self.fused = bool(fused_ops) and self.device.type == "cuda"
        self.use_graph = bool(use_cuda_graph) and self.fused
        self._graph = None
        self._ugraph, self._ugraph_tries = None, 0
        _seg_offsets(self.device)
        self.use_update_graph = self.use_graph
        self.update_graph_error = None
        if self.fused:
            from . import abi

            self._L = abi.lib()
            self._ctr = torch.zeros(1, dtype=torch.int64, device=dev)  # Philox time counter, advanced on the device
            self._ret = torch.zeros(T, N, 8, device=dev)
            self._adv = torch.zeros(T, N, 8, device=dev)
        if impl is None:
            impl = "umma" if self.fused else "torch"
        if impl not in ("umma", "torch") or (impl == "umma" and not self.fused):
            raise ValueError("impl must be 'umma' (CUDA, fused ops) or 'torch'")
        self.impl = impl
        self.engine, self.clip_adam = None, None
        if impl == "umma":
            from .a2c_umma import UmmaEngine

            self.engine = UmmaEngine(self.net, self.obs, self.masks, self.actions, self.values, N, T, passes=gemm_passes)
            if fused_optimizer:
                from .a2c_umma import ClipAdam

                self.clip_adam = ClipAdam(self.net, self.opt, self.max_grad_norm)
        self.frames = 0
        self.stats = {}
        o, m = env.reset()
        self.obs[0].copy_(o), self.masks[0].copy_(m)

    # ------------------------------------------------------------------ rollout
    def _stream(self):
        import ctypes

        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _rollout_steps(self):
        """T x (policy forward, masked Categorical sample, ONE env-step launch).  Static buffers: capturable."""
        T, env = self.T, self.env
        if self.engine is not None:
            eng = self.engine
            eng.repack()   # packed weight images of this cycle's weights (the optimizer step changed them)
            for t in range(T):
                eng.forward_actors(t)  # logits -> eng.logits[t]; activations kept for the update
                rc = self._L.fjsp_a2c_sample(_ptr(eng.logits[t]), _ptr(self.masks[t]), _ptr(self.actions[t]), None, env.num_envs,
                                             env.first_env, self.seed, _ptr(self._ctr), t, self._stream())
                assert rc == 0, self._L.fjsp_last_error()
                env.step_into(self.actions[t], self.obs[t + 1], self.masks[t + 1], self.rewards[t], self.flags[t])
            eng.forward_critic()   # values[0..T] (incl. the bootstrap value) in one batched pass: the policy never reads them
            rc = self._L.fjsp_a2c_counter_add(_ptr(self._ctr), T, self._stream())
            assert rc == 0
            return
        for t in range(T):
            o, m = self.obs[t], self.masks[t]
            self.values[t].copy_(self.net.value(o))
            if self.fused:
                z = self.net.logits32(o)
                rc = self._L.fjsp_a2c_sample(_ptr(z), _ptr(m), _ptr(self.actions[t]), None, env.num_envs, env.first_env,
                                             self.seed, _ptr(self._ctr), t, self._stream())
                assert rc == 0, self._L.fjsp_last_error()
            else:
                q = masked_policy(self.net.probs32(o), m)
                self.actions[t] = sample_actions(q, self.gen)
            # the step kernel writes the next observation / mask straight into the rollout buffer
            env.step_into(self.actions[t], self.obs[t + 1], self.masks[t + 1], self.rewards[t], self.flags[t])
        self.values[T].copy_(self.net.value(self.obs[T]))
        if self.fused:
            rc = self._L.fjsp_a2c_counter_add(_ptr(self._ctr), T, self._stream())
            assert rc == 0

    @torch.no_grad()
    def rollout(self):
        if self.use_graph:
            if self._graph is None:
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):  # warm-up on a side stream (cuBLAS workspaces, lazy init) before capture
                    self._rollout_steps()
                torch.cuda.current_stream(self.device).wait_stream(side)
                torch.cuda.synchronize(self.device)
                # that eager pass IS this call's rollout (its buffers are consistent); capturing executes nothing, it only
                # records the same launch sequence over the same static buffers for every later call
                state, ctr = self.env.save_state(), self._ctr.clone()
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._rollout_steps()
                self.env.load_state(state), self._ctr.copy_(ctr)  # capture must not have side effects
            else:
                self._graph.replay()
        else:
            self._rollout_steps()
        self.frames += self.T * self.env.num_envs

    # ------------------------------------------------------------------ update
    def _allreduce_(self, t):
        if self.world > 1:
            dist.all_reduce(t)
        return t

    def update(self):
        """One A2C update from the rollout buffers.  On CUDA the whole update (GAE, forward, backward, gradient all-reduce,
        clipping, Adam) is captured in a CUDA graph after two eager warm-up updates and replayed afterwards."""
        if not self.use_update_graph:
            return self._update_impl()
        if self._ugraph is not None:
            self._ugraph.replay()
            return None
        self._ugraph_tries += 1
        if self._ugraph_tries <= 2:
            return self._update_impl()
        torch.cuda.synchronize(self.device)
        try:
            g = torch.cuda.CUDAGraph()
            if self.engine is None:
                self.opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(g):
                self._update_impl()
            self._ugraph = g
            g.replay()  # capture executed nothing: this replay IS this call's update
        except Exception as e:  # noqa: BLE001 - fall back to the eager update, keep training
            self.update_graph_error = repr(e)
            self.use_update_graph = False
            torch.cuda.synchronize(self.device)
            if self.engine is None:
                self.opt.zero_grad(set_to_none=False)
            return self._update_impl()
        return None

    def _update_impl(self):
        self._compute_grads()
        if self.clip_adam is not None:
            self.clip_adam.step()
            return
        self._clip()
        self.opt.step()

    def _compute_grads(self):
        """GAE, advantage moments, losses and their gradients (all-reduced over the ranks) into ``p.grad``."""
        T, N = self.T, self.env.num_envs
        if self.fused:
            rc = self._L.fjsp_a2c_gae(_ptr(self.rewards), _ptr(self.values), _ptr(self.flags), _ptr(self._ret), _ptr(self._adv),
                                      T, N, self.gamma, self.lamb, self._stream())
            assert rc == 0, self._L.fjsp_last_error()
            returns, advs = self._ret, self._adv
        else:
            dones = (self.flags[:, :, 0:3] != 0).any(-1)
            returns, advs = gae_and_returns(self.rewards, self.values, dones, self.gamma, self.lamb)
        B = T * N
        obs, masks = self.obs[:T].reshape(B, 38), self.masks[:T].reshape(B, 32)
        acts, returns, advs = self.actions.reshape(B, 8), returns.reshape(B, 8), advs.reshape(B, 8)
        # advantage normalisation per agent over the (global) batch, unbiased std (a2c.py:727-729)
        mom = torch.stack([torch.full((8,), float(B), device=self.device), advs.sum(0), (advs * advs).sum(0)])
        if self.global_adv_norm:
            self._allreduce_(mom)
        cnt, mean = mom[0], mom[1] / mom[0]
        var = (mom[2] - cnt * mean * mean) / (cnt - 1).clamp_min(1.0)
        if self.engine is not None:  # tensor-core path: analytic loss gradients + grouped tcgen05 GEMMs, no autograd
            eng = self.engine
            eng.adv_mean.copy_(mean), eng.adv_rstd.copy_(1.0 / (var.clamp_min(0).sqrt() + 1e-8))
            eng.backward(self._adv, self._ret, self.entropy_coef)
            if self.world > 1:  # ONE all-reduce of the flat gradient buffer (2.62 MB), then the mean over ranks
                dist.all_reduce(eng.grad_flat)
                eng.grad_flat /= self.world
            self.stats = eng.stats(self.entropy_coef)
            return
        adv_n = (advs - mean) / (var.clamp_min(0).sqrt() + 1e-8)

        probs = self.net.probs32(obs)
        q = masked_policy(probs, masks)
        logp = log_prob_of(q, acts)
        ent = entropy_unmasked(probs).mean(0)
        actor_loss = -(adv_n * logp).mean(0) - self.entropy_coef * ent  # [8], one loss per agent (separate networks)
        v = self.net.value(obs)
        critic_loss = F.mse_loss(v.unsqueeze(-1).expand_as(returns), returns)
        if not torch.cuda.is_current_stream_capturing() if self.device.type == "cuda" else True:
            self.opt.zero_grad(set_to_none=True)
        (actor_loss.sum() + critic_loss).backward()
        if self.world > 1:  # ONE flat all-reduce of all gradients (2.62 MB), then the mean over ranks
            grads = [p.grad for p in self.net.parameters()]
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat)
            flat /= self.world
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        self.stats = {"actor_loss": actor_loss.detach(), "critic_loss": critic_loss.detach(), "entropy": ent.detach()}

    def _clip(self):
        """clip_grad_norm_(max_grad_norm) per NETWORK: each of the 8 actors and the critic separately (a2c.py:668,686)."""
        for _, params, lead in self.net.networks():
            if lead is None:
                torch.nn.utils.clip_grad_norm_(params, self.max_grad_norm)
            else:
                sq = sum((p.grad.reshape(lead, -1) ** 2).sum(-1) for p in params)
                scale = (self.max_grad_norm / (sq.sqrt() + 1e-6)).clamp_max(1.0)
                for p in params:
                    p.grad.mul_(scale.view(lead, *([1] * (p.grad.dim() - 1))))

    # ------------------------------------------------------------------ driver
    def train(self, updates: int):
        """`updates` x (rollout of T steps + one update).  Returns frames/s of this rank (device-timed when on CUDA)."""
        cuda = self.device.type == "cuda"
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(self.device)
            e0.record()
        t0 = time.perf_counter()
        f0 = self.frames
        for _ in range(updates):
            self.rollout()
            self.update()
            # next rollout continues from the last observation
            self.obs[0].copy_(self.obs[self.T]), self.masks[0].copy_(self.masks[self.T])
        if cuda:
            e1.record()
            torch.cuda.synchronize(self.device)
            secs = e0.elapsed_time(e1) * 1e-3
        else:
            secs = time.perf_counter() - t0
        return (self.frames - f0) / secs, secs

    def mean_reward(self):
        return float(self.rewards.sum(-1).mean())
