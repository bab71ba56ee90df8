"""Tensor-core engine of the batched A2C trainer: the reference's 8 actors + centralised critic
(/root/reference/networks.py:22-61) forward and backward as grouped tcgen05 GEMMs (``fjsp_a2c_gemm``, 3xTF32 =
fp32-level accuracy) plus one analytic loss-gradient kernel (``fjsp_a2c_loss_grad``) — no autograd graph, no library GEMM.

A2C is on-policy: the update differentiates the very networks that produced the rollout.  The rollout's forward passes
therefore ARE the update's forward pass: every rollout step stores its hidden activations (post-ReLU) and pre-softmax
outputs in slice ``t`` of ``[net][T+1][N][256]`` buffers (actors: two launches per step, the logits layer fused into the
second layer's epilogue; critic: three launches over all ``(T+1)*N`` rows once the rollout is over — the policy does not
read it), and the update only runs the backward chain over the ``T*N`` rows:

    loss_grad            dlogits [B,32], dvalue [B]  (+ last-layer bias gradients, loss statistics)
    head_backward x9     dH2 = (dlogits_i W3_i^T) * (H2 > 0) with the bias gradient of layer 2 and dW3_i = H2^T dlogits_i in the same
                         pass over H2 (fp32 FMAs); the critic's value head dH3 = (dvalue w4^T) * (H3 > 0), dw4 likewise
    (KC,KC)   x1         critic: dH2 = (dH3 Wc3^T) * (H2 > 0)
    (KC,KC)   x9         dH1 = (dH2 W2^T) * (H1 > 0)
    (MC,MC)   x11        the 256 x 256 (and the critic's 256 x 128 and 38 x 256) weight gradients, split-K over the batch, atomic
                         accumulation
    wgrad_small x8       the narrow first layers (3..13 x 256): fp32 FMAs, one pass
                         over the wide operand (as GEMMs each costs as much as a 256 x 256 product)
    bias gradients       column sums in the epilogues of the dH GEMMs (same pass)

Gradients land in ONE flat fp32 buffer (``p.grad`` are views of it), so data parallelism is a single NCCL all-reduce of
that buffer without gather / scatter copies.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import abi, umma
from .env import MASK_OFFSETS, N_ACTIONS, OBS_SLICES

HID = 256
LS_DB, LS_ACTOR, LS_ENT, LS_CRITIC, LS_DV, LS_WORDS = 0, 32, 40, 48, 49, 64


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class UmmaEngine:
    def __init__(self, net, obs, masks, actions, values, num_envs, rollout_len, passes=3):
        """net: ``ActorCritic``; obs [T+1,N,38], masks [T+1,N,32], actions [T,N,8], values [T+1,N]: the trainer's rollout buffers."""
        self.net, self.N, self.T, self.passes = net, int(num_envs), int(rollout_len), int(passes)
        self.obs, self.masks, self.actions, self.values = obs, masks, actions, values
        dev = self.dev = obs.device
        N, T = self.N, self.T
        self.B = T * N
        z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
        self.h1, self.h2, self.h3 = z(9, T + 1, N, HID), z(9, T + 1, N, HID), z(T + 1, N, 128)
        self.logits = z(T + 1, N, 32)
        self.dh1, self.dh2, self.dh3 = z(9, self.B, HID), z(9, self.B, HID), z(self.B, 128)
        self.dlogits, self.dvalue = z(self.B, 32), z(self.B)
        self.sums = z(LS_WORDS)
        self.adv_mean, self.adv_rstd = z(8), z(8)
        # one flat gradient buffer; p.grad are views of it
        params = list(net.parameters())
        self.grad_flat = z(sum(p.numel() for p in params))
        off = 0
        for p in params:
            p.grad = self.grad_flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self._L = abi.lib()
        self._nets = self._describe()
        # The 256-wide weight matrices as PACKED (hi, lo) images (fjsp_a2c_gemm_pack, FJSP_OP_PK), one per orientation they are
        # read in: a K chunk of an image is the B tile of a pipeline stage byte for byte, so the GEMM kernel's B loader moves it
        # with one bulk copy (cp.async.bulk) and no thread touches the weights.  Re-packed once per cycle (``repack``: 13 us).
        self._pk = umma.PackTable(dev)
        self._pk_fwd2, self._pk_dx2 = [], []
        for d in self._nets:
            (w2, o2) = self._par(d, 2)
            self._pk_fwd2.append(self._pk.add(w2, umma.OP_MC, HID, HID, HID, o2))    # y = h1 W2: B(n, k) = W2[k, n]
            self._pk_dx2.append(self._pk.add(w2, umma.OP_KCS, HID, HID, HID, o2))    # dh1 = dh2 W2^T: B(n, k) = W2[n, k]
        w3c = self._nets[8]["p"][4]
        self._pk_fwd3c = self._pk.add(w3c, umma.OP_MC, 128, 128, HID)                # critic 256 -> 128
        self._pk_dx3c = self._pk.add(w3c, umma.OP_KCS, 128, HID, 128)                # dh2 = dh3 W3^T
        self._pk.finalize()
        self._fwd_actors = [self._actor_tables(t) for t in range(T)]
        self._fwd_critic = self._critic_tables(0, T + 1)
        self._fwd_critic_step = {}
        self._bwd = self._backward_tables()

    # ------------------------------------------------------------------ the nine networks
    def _describe(self):
        n, nets = self.net, []
        ps, agv, six, cr = list(n.ps), list(n.agv), list(n.six), list(n.critic)
        nets.append(dict(lo=0, k1=7, p=ps, i=None, nact=3, zoff=MASK_OFFSETS[0]))
        nets.append(dict(lo=7, k1=13, p=agv, i=None, nact=8, zoff=MASK_OFFSETS[1]))
        for i in range(6):
            nets.append(dict(lo=OBS_SLICES[2 + i][0], k1=3, p=six, i=i, nact=N_ACTIONS[2 + i], zoff=MASK_OFFSETS[2 + i]))
        nets.append(dict(lo=0, k1=38, p=cr, i=None, nact=128, zoff=None))
        return nets

    @staticmethod
    def _par(d, j):
        """(tensor, element offset) of parameter j (w1,b1,w2,b2,w3,b3[,w4,b4]) of a network (one slice of a stacked one)."""
        t = d["p"][j]
        return (t, 0) if d["i"] is None else (t, d["i"] * t[0].numel())

    def _grad(self, d, j):
        t = d["p"][j].grad
        return (t, 0) if d["i"] is None else (t, d["i"] * t[0].numel())

    def _hoff(self, net, t=0):
        return ((net * (self.T + 1) + t) * self.N) * HID

    # ------------------------------------------------------------------ forward
    def _actor_tables(self, t):
        """Rollout step t, the eight actors: layer 1 from the observation slices (K = 3..13: fp32 FMAs), layer 2 with the 256 -> 3..8 logits layer as
        a fused head in its epilogue (fp32 FMAs on the values it stores) — TWO grouped launches.  As a third GEMM launch the
        heads cost a step as much as the 256 x 256 layer: 16 K chunks of a latency-bound pipeline for 1-3 % of its flops."""
        N, dev, ps = self.N, self.dev, self.passes
        l1 = umma.Layer1Table(dev)   # K = 3..13: fp32 FMAs (fjsp_a2c_layer1), not a one-chunk tensor-core launch
        l2 = umma.GemmTable(dev, umma.OP_KC, umma.OP_PK, ps)
        for k, d in enumerate(self._nets[:8]):
            (w1, o1), (b1, ob1), (w2, o2), (b2, ob2), (w3, o3), (b3, ob3) = (self._par(d, j) for j in range(6))
            l1.add(self.obs, w1, self.h1, N, HID, d["k1"], ldx=38, ldy=HID, x_off=t * N * 38 + d["lo"], w_off=o1,
                   y_off=self._hoff(k, t), bias=b1, bias_off=ob1, relu=True)
            l2.add(self.h1, self._pk.image, self.h2, N, HID, HID, lda=HID, ldb=0, csm=HID, a_off=self._hoff(k, t), b_off=self._pk_fwd2[k],
                   c_off=self._hoff(k, t), bias=b2, bias_off=ob2, relu=True, rowdot_w=w3, rowdot_w_off=o3, rowdot_bias=b3,
                   rowdot_bias_off=ob3, rowdot_out=self.logits, rowdot_out_off=t * N * 32 + d["zoff"], head_n=d["nact"], head_ld=32)
        return [x.finalize() for x in (l1, l2)]

    def _critic_tables(self, t0, steps):
        """The critic over the rows of rollout steps t0 .. t0 + steps - 1 as ONE problem per layer (the buffers are contiguous
        in t): 38 -> 256 -> 256 -> 128 (ReLU) with the 128 -> 1 head fused into the third layer's epilogue."""
        N, dev, ps = self.N, self.dev, self.passes
        M = steps * N
        d = self._nets[8]
        (w1, o1), (b1, ob1), (w2, o2), (b2, ob2), (w3, o3), (b3, ob3) = (self._par(d, j) for j in range(6))
        w4, b4 = d["p"][6], d["p"][7]
        l1 = umma.GemmTable(dev, umma.OP_KCS, umma.OP_MC, ps)
        l2 = umma.GemmTable(dev, umma.OP_KC, umma.OP_PK, ps)
        l3 = umma.GemmTable(dev, umma.OP_KC, umma.OP_PK, ps)
        l1.add(self.obs, w1, self.h1, M, HID, 38, lda=38, ldb=HID, csm=HID, a_off=t0 * N * 38, b_off=o1, c_off=self._hoff(8, t0),
               bias=b1, bias_off=ob1, relu=True)
        l2.add(self.h1, self._pk.image, self.h2, M, HID, HID, lda=HID, ldb=0, csm=HID, a_off=self._hoff(8, t0), b_off=self._pk_fwd2[8],
               c_off=self._hoff(8, t0), bias=b2, bias_off=ob2, relu=True)
        l3.add(self.h2, self._pk.image, self.h3, M, 128, HID, lda=HID, ldb=0, csm=128, a_off=self._hoff(8, t0), b_off=self._pk_fwd3c, c_off=t0 * N * 128,
               bias=b3, bias_off=ob3, relu=True, rowdot_w=w4, rowdot_out=self.values, rowdot_out_off=t0 * N, rowdot_bias=b4)
        return [x.finalize() for x in (l1, l2, l3)]

    def repack(self):
        """Rebuild the packed weight images from the current weights (one launch).  Call after the weights changed (optimizer
        step, checkpoint load) and before the next forward / backward."""
        self._pk.launch()

    def forward_actors(self, t):
        """Logits of rollout step t into ``self.logits[t]`` (two grouped launches); hidden activations kept for the update."""
        for tab in self._fwd_actors[t]:
            tab.launch()

    def forward_critic(self):
        """Values of ALL T + 1 observations of the rollout (``values[0..T]``, the last one being the bootstrap value) in three
        launches over (T + 1) N rows.  The policy does not read the critic, so nothing of it has to run inside the rollout's
        step-by-step chain, where a launch is one wave of latency-bound CTAs; here the same work is 33 times as many rows."""
        for tab in self._fwd_critic:
            tab.launch()

    def forward(self, t):
        """Everything of rollout step t in one call (tests, single steps): the actors' logits (t < T) and the critic value."""
        self.repack()
        if t < self.T:
            self.forward_actors(t)
        if t not in self._fwd_critic_step:
            self._fwd_critic_step[t] = self._critic_tables(t, 1)
        for tab in self._fwd_critic_step[t]:
            tab.launch()

    # ------------------------------------------------------------------ backward (one update)
    def _backward_tables(self):
        B, dev, ps = self.B, self.dev, self.passes
        hb = umma.HeadBwdTable(dev)
        bc = umma.GemmTable(dev, umma.OP_KC, umma.OP_PK, ps)
        b2 = umma.GemmTable(dev, umma.OP_KC, umma.OP_PK, ps)
        dw = umma.GemmTable(dev, umma.OP_MC, umma.OP_MC, ps)
        ws = umma.WgradTable(dev)
        sk = max(1, min(64, B // 2048))
        for k, d in enumerate(self._nets):
            (w1, o1), _, (w2, o2), _, (w3, o3), _ = (self._par(d, j) for j in range(6))
            (gw1, g1), (gb1, gb1o), (gw2, g2), (gb2, gb2o), (gw3, g3), (gb3, gb3o) = (self._grad(d, j) for j in range(6))
            ho, go = self._hoff(k), k * B * HID
            if k < 8:
                na, zo = d["nact"], d["zoff"]
                # backward through the logits layer in one pass over H2: dH2, the bias gradient of layer 2 and the head's own
                # weight gradient (was: a K <= 8 tensor-core launch that is all epilogue + a narrow-gradient job, H2 read twice)
                hb.add(self.dlogits, w3, self.h2, self.dh2, gw3, gb2, B, HID, na, 32, dl_off=zo, w_off=o3, h_off=ho, dh_off=go, gw_off=g3,
                       gb_off=gb2o)
            else:
                w4, gw4 = d["p"][6], d["p"][6].grad
                hb.add(self.dvalue, w4, self.h3, self.dh3, gw4, gb3, B, 128, 1, 1, gb_off=gb3o)   # the value head, the same way
                bc.add(self.dh3, self._pk.image, self.dh2, B, HID, 128, lda=128, ldb=0, csm=HID, b_off=self._pk_dx3c, c_off=go, mask=self.h2,
                       mask_off=ho, colsum=gb2, colsum_off=gb2o)
                dw.add(self.h2, self.dh3, gw3, HID, 128, B, lda=HID, ldb=128, csm=128, a_off=ho, c_off=g3, atomic=True, splitk=sk)
            b2.add(self.dh2, self._pk.image, self.dh1, B, HID, HID, lda=HID, ldb=0, csm=HID, a_off=go, b_off=self._pk_dx2[k], c_off=go, mask=self.h1,
                   mask_off=ho, colsum=gb1, colsum_off=gb1o)
            dw.add(self.h1, self.dh2, gw2, HID, HID, B, lda=HID, ldb=HID, csm=HID, a_off=ho, b_off=go, c_off=g2, atomic=True, splitk=sk)
            # dW1[i, j] = sum_rows obs[row, lo + i] * dH1[row, j]: X = dH1 (wide), Y = the observation slice
            if d["k1"] > 16:   # the critic's 38 x 256: 38 FMAs per streamed element make the FMA kernel compute-bound (0.21 ms);
                #                   one 128-row tile of the tensor-core GEMM (A = the observation rows, 38 of 128 used) is cheaper
                dw.add(self.obs, self.dh1, gw1, d["k1"], HID, B, lda=38, ldb=HID, csm=HID, a_off=d["lo"], b_off=go, c_off=g1, atomic=True,
                       splitk=sk)
            else:
                ws.add(self.dh1, self.obs, gw1, B, HID, d["k1"], ldx=HID, ldy=38, gsi=1, gsj=HID, x_off=go, y_off=d["lo"], g_off=g1)
        return [x.finalize() for x in (hb, bc, b2, dw, ws)]

    def backward(self, adv, returns, entropy_coef):
        """Gradients of the update's losses into ``grad_flat`` (local-batch means; the caller all-reduces and averages).
        ``adv`` / ``returns`` [T,N,8]; ``self.adv_mean`` / ``self.adv_rstd`` must hold the (global) advantage moments."""
        B = self.B
        self.repack()   # (13 us; the rollout packed the same weights already — kept so that backward() stands on its own)
        self.grad_flat.zero_()
        self.sums.zero_()
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        rc = self._L.fjsp_a2c_loss_grad(_p(self.logits), _p(self.masks), _p(self.actions), _p(adv), _p(returns), _p(self.values),
                                        _p(self.adv_mean), _p(self.adv_rstd), float(entropy_coef), B, _p(self.dlogits), _p(self.dvalue),
                                        _p(self.sums), st)
        if rc:
            abi.check(rc)
        for tab in self._bwd:
            tab.launch()
        # last-layer bias gradients = column sums of dlogits / dvalue (from the loss kernel)
        n = self.net
        n.ps[5].grad.view(-1).copy_(self.sums[0:3])
        n.agv[5].grad.view(-1).copy_(self.sums[3:11])
        n.six[5].grad.view(-1).copy_(self.sums[11:29])
        n.critic[7].grad.view(-1).copy_(self.sums[LS_DV:LS_DV + 1])

    def stats(self, entropy_coef):
        inv_b = 1.0 / self.B
        ent = self.sums[LS_ENT:LS_ENT + 8] * inv_b
        return {"actor_loss": self.sums[LS_ACTOR:LS_ACTOR + 8] * inv_b - entropy_coef * ent,
                "critic_loss": self.sums[LS_CRITIC] * (inv_b / 8), "entropy": ent}


class ClipAdam:
    """``clip_grad_norm_(max_norm)`` per network followed by ``Adam.step()`` (/root/reference/a2c.py:668,686-690) as ONE
    C-ABI call (``fjsp_a2c_clip_adam``: three launches) over a device table of parameter segments, on the state tensors of
    the trainer's own ``torch.optim.Adam`` (``exp_avg``, ``exp_avg_sq``, ``step``) — ``opt.state_dict()``, checkpoints and
    a later ``opt.step()`` see exactly what torch's step would have left there.  A segment is one (tensor, network)
    pair: the six small actors are slices of stacked tensors; each network has its own norm."""

    def __init__(self, net, opt, max_norm):
        import numpy as np

        self.max_norm = float(max_norm)
        g0 = opt.param_groups[0]
        self.beta1, self.beta2, self.eps = float(g0["betas"][0]), float(g0["betas"][1]), float(g0["eps"])
        lr_of = {id(p): float(g["lr"]) for g in opt.param_groups for p in g["params"]}
        for g in opt.param_groups:
            assert not g.get("amsgrad") and not g.get("weight_decay") and not g.get("maximize"), "plain Adam only"
            assert tuple(g["betas"]) == (self.beta1, self.beta2) and g["eps"] == self.eps
        rows, net_idx, dev = [], 0, None
        for _, params, lead in net.networks():
            for p in params:
                dev = p.device
                st = opt.state[p]
                if len(st) == 0:  # what Adam._init_group creates at the first step (capturable: the step lives on the device)
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                assert st["step"].is_cuda and st["step"].dtype == torch.float32 and p.grad is not None and p.is_contiguous()
                parts = 1 if lead is None else int(lead)
                n = p.numel() // parts
                for i in range(parts):
                    r = np.zeros((), dtype=abi.OPT_SEG_DT)
                    for key, t in (("param", p.data), ("grad", p.grad), ("m", st["exp_avg"]), ("v", st["exp_avg_sq"])):
                        r[key] = t.data_ptr() + 4 * i * n
                    r["step"], r["n"], r["net"], r["lr"], r["bump"] = st["step"].data_ptr(), n, net_idx + i, lr_of[id(p)], int(i == 0)
                    rows.append(r)
            net_idx += 1 if lead is None else int(lead)
        assert net_idx <= 16
        self.nseg, self.max_elems = len(rows), max(int(r["n"]) for r in rows)
        self._keep = opt
        self.table = torch.from_numpy(np.stack(rows).view(np.uint8).reshape(-1).copy()).to(dev)
        self.norms_sq = torch.zeros(16, device=dev)
        self._L = abi.lib()

    def step(self):
        st = C.c_void_p(torch.cuda.current_stream(self.table.device).cuda_stream)
        rc = self._L.fjsp_a2c_clip_adam(_p(self.table), self.nseg, self.max_elems, _p(self.norms_sq), self.max_norm, self.beta1,
                                        self.beta2, self.eps, st)
        if rc:
            abi.check(rc)
