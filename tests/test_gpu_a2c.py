"""Numerics of the fused A2C device ops against their plain PyTorch fp32 references, and a short training run."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(t.data_ptr())


def test_policy_sample_kernel_vs_torch_reference():
    from multi_agent_rl_for_fjsp_b200 import a2c_batched as A, abi
    from multi_agent_rl_for_fjsp_b200.env import MASK_OFFSETS, N_ACTIONS

    L = abi.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    B = 1 << 16
    logits = torch.zeros(B, 32, device=dev)
    logits[:, :29] = torch.randn(B, 29, device=dev, generator=g) * 3
    masks = (torch.rand(B, 32, device=dev, generator=g) < 0.6).to(torch.int8)
    for off in MASK_OFFSETS[:8]:
        masks[:, off] = 1
    logits[:64, 3] = 200.0   # all softmax mass on AGV action 0 ...
    masks[:64, 3] = 0        # ... masked out -> uniform-over-valid fallback
    masks[:64, 4:6] = 1
    acts = torch.zeros(B, 8, dtype=torch.uint8, device=dev)
    logp = torch.zeros(B, 8, device=dev)
    ctr = torch.tensor([5], dtype=torch.int64, device=dev)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.fjsp_a2c_sample(_p(logits), _p(masks), _p(acts), _p(logp), B, 1000, 42, _p(ctr), 3, s) == 0
    probs = torch.zeros(B, 32, device=dev)
    for off, n in zip(MASK_OFFSETS, N_ACTIONS):
        probs[:, off:off + n] = torch.softmax(logits[:, off:off + n], -1)
    q = A.masked_policy(probs, masks)
    idx = acts.long() + torch.tensor(MASK_OFFSETS[:8], device=dev)
    assert (masks.gather(1, idx) == 1).all(), "sampled a masked action"
    ref_logp = A.log_prob_of(q, acts)
    assert torch.allclose(logp, ref_logp, rtol=1e-4, atol=1e-5), float((logp - ref_logp).abs().max())
    # same counter -> same draw; next counter -> different draw; distribution follows q
    acts2 = torch.zeros_like(acts)
    assert L.fjsp_a2c_sample(_p(logits), _p(masks), _p(acts2), None, B, 1000, 42, _p(ctr), 3, s) == 0
    assert torch.equal(acts, acts2)
    assert L.fjsp_a2c_sample(_p(logits), _p(masks), _p(acts2), None, B, 1000, 42, _p(ctr), 4, s) == 0
    assert not torch.equal(acts, acts2)
    row = logits[100:101].repeat(B, 1).contiguous()
    mrow = masks[100:101].repeat(B, 1).contiguous()
    assert L.fjsp_a2c_sample(_p(row), _p(mrow), _p(acts2), None, B, 0, 7, None, 0, s) == 0
    freq = torch.bincount(acts2[:, 1].long(), minlength=8).float() / B
    assert torch.allclose(freq, q[100, 3:11], atol=0.01)


def test_gae_kernel_vs_torch_reference():
    from multi_agent_rl_for_fjsp_b200 import a2c_batched as A, abi

    L = abi.lib()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    T, N = 32, 5000
    rewards, values = torch.randn(T, N, 8, device=dev), torch.randn(T + 1, N, device=dev)
    flags = torch.zeros(T, N, 4, dtype=torch.uint8, device=dev)
    flags[:, :, 0] = (torch.rand(T, N, device=dev) < 0.02).to(torch.uint8)
    flags[:, :, 1] = (torch.rand(T, N, device=dev) < 0.02).to(torch.uint8)
    flags[:, :, 3] = 1  # was_reset must not count as an episode end
    ret, adv = torch.zeros_like(rewards), torch.zeros_like(rewards)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.fjsp_a2c_gae(_p(rewards), _p(values), _p(flags), _p(ret), _p(adv), T, N, 0.99, 0.95, s) == 0
    dones = (flags[:, :, 0:3] != 0).any(-1)
    r_ref, a_ref = A.gae_and_returns(rewards, values, dones, 0.99, 0.95)
    assert torch.allclose(ret, r_ref, rtol=1e-5, atol=1e-5) and torch.allclose(adv, a_ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("graph", [False, True])
def test_training_runs_and_fused_path_matches_torch_path(graph):
    """A few updates on 512 envs: finite losses, parameters move, frames counted; the CUDA-graph rollout equals the
    eager fused rollout bit for bit (same Philox counters)."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv
    from multi_agent_rl_for_fjsp_b200.a2c_batched import BatchedA2C

    def run(use_graph):
        env = BatchedFJSPEnv(512, seed=3, num_orders=25, autoreset=True)
        tr = BatchedA2C(env, rollout_len=8, seed=2, use_cuda_graph=use_graph, impl="torch")  # deterministic reductions
        p0 = torch.cat([p.detach().reshape(-1) for p in tr.net.parameters()]).clone()
        tr.train(3)
        p1 = torch.cat([p.detach().reshape(-1) for p in tr.net.parameters()])
        assert torch.isfinite(p1).all() and not torch.equal(p0, p1)
        assert torch.isfinite(tr.stats["critic_loss"]) and tr.frames == 3 * 8 * 512
        return p1, tr.actions.clone(), tr.rewards.clone()

    a = run(graph)
    b = run(False)
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert torch.allclose(a[0], b[0], rtol=1e-5, atol=1e-6)


def test_reference_layout_checkpoint_round_trip_on_device(tmp_path):
    """SURVEY §8f rank 4 on the device: a file written by ``save_reference_checkpoint`` (the dict layout of the
    reference's ``MultiAgentA2C.save_model``, a2c.py:745-752) is read back by plain ``nn.Sequential`` networks shaped like
    the reference's (networks.py:22-61) and by a second ``ActorCritic`` on the GPU; all three agree on device."""
    from multi_agent_rl_for_fjsp_b200.a2c_batched import ActorCritic
    from multi_agent_rl_for_fjsp_b200.env import AGENT_IDS, MASK_OFFSETS, N_ACTIONS, OBS_SLICES

    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    src = ActorCritic(device=dev, seed=123)
    path = str(tmp_path / "model.pt")
    src.save_reference_checkpoint(path)
    ckpt = torch.load(path, weights_only=True, map_location="cpu")
    assert set(ckpt) == {"actor_nets", "critic_net", "obs_dims", "act_dims", "global_obs_dim", "possible_agents"}
    assert ckpt["possible_agents"] == list(AGENT_IDS) and ckpt["global_obs_dim"] == 38

    def sequential(sd, dims, softmax):
        layers = []
        for i in range(len(dims) - 1):
            layers.append(torch.nn.Linear(dims[i], dims[i + 1]))
            if i < len(dims) - 2:
                layers.append(torch.nn.ReLU())
        net = torch.nn.Sequential(*layers)
        net.load_state_dict({k[len("net."):]: v for k, v in sd.items()})
        return net.to(dev), softmax

    g = torch.Generator(device=dev).manual_seed(1)
    obs = torch.randint(0, 6, (512, 38), device=dev, generator=g).float()
    dst = ActorCritic(device=dev, seed=7)
    dst.load_reference_checkpoint(path)
    with torch.no_grad():
        p_src, p_dst = src.probs32(obs), dst.probs32(obs)
        v_src, v_dst = src.value(obs), dst.value(obs)
        assert torch.equal(p_src, p_dst) and torch.equal(v_src, v_dst)
        for i, a in enumerate(AGENT_IDS):
            lo, hi = OBS_SLICES[i]
            net, _ = sequential(ckpt["actor_nets"][a], [hi - lo, 256, 256, N_ACTIONS[i]], True)
            ref = torch.softmax(net(obs[:, lo:hi]), -1)
            assert torch.allclose(p_src[:, MASK_OFFSETS[i]:MASK_OFFSETS[i] + N_ACTIONS[i]], ref, rtol=1e-5, atol=1e-6), a
        critic, _ = sequential(ckpt["critic_net"], [38, 256, 256, 128, 1], False)
        assert torch.allclose(v_src, critic(obs).squeeze(-1), rtol=1e-5, atol=1e-5)


def _pair(n_envs, T, passes=3):
    """Two trainers with identical weights on identical envs: the tensor-core engine and the torch / autograd statement."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv
    from multi_agent_rl_for_fjsp_b200.a2c_batched import BatchedA2C

    torch.backends.cuda.matmul.allow_tf32 = False
    mk = lambda impl: BatchedA2C(BatchedFJSPEnv(n_envs, seed=3, num_orders=25, autoreset=True), rollout_len=T, seed=2,  # noqa: E731
                                 use_cuda_graph=False, impl=impl, gemm_passes=passes)
    return mk("umma"), mk("torch")


def test_umma_forward_equals_torch_fp32_forward():
    """Rollout forward of the 8 actors + critic as grouped tcgen05 GEMMs vs the same networks in torch fp32 (rtol 1e-5)."""
    a, b = _pair(1000, 4)
    a.rollout()
    with torch.no_grad():
        for t in range(a.T):
            z = b.net.logits32(a.obs[t])
            v = b.net.value(a.obs[t])
            assert torch.allclose(a.engine.logits[t], z, rtol=1e-5, atol=2e-5), float((a.engine.logits[t] - z).abs().max())
            assert torch.allclose(a.values[t], v, rtol=1e-5, atol=2e-5), float((a.values[t] - v).abs().max())
        assert torch.allclose(a.values[a.T], b.net.value(a.obs[a.T]), rtol=1e-5, atol=2e-5)
        h2 = torch.relu(torch.addmm(b.net.agv[3][0], torch.relu(torch.addmm(b.net.agv[1][0], a.obs[1][:, 7:20], b.net.agv[0])), b.net.agv[2]))
        assert torch.allclose(a.engine.h2[1, 1], h2, rtol=1e-5, atol=2e-5)


def _fp64_grads(tr, relu_masks=None):
    """The update's gradients from the trainer's rollout buffers in float64 (torch autograd): the yardstick for both paths.

    relu_masks = (h1 [9,B,256], h2 [9,B,256], h3 [B,128]) booleans: the ReLU gates are taken from there instead of from the
    float64 pre-activations.  A unit whose pre-activation is within rounding distance of zero is open in one
    implementation and shut in another; its forward contribution is ~0 either way, but its GRADIENT contribution is not
    small — a property of ReLU networks under any change of summation order, not an accuracy defect.  With the gates
    pinned, what is left is the arithmetic of the backward pass."""
    import copy

    from multi_agent_rl_for_fjsp_b200 import a2c_batched as A
    from multi_agent_rl_for_fjsp_b200.env import MASK_OFFSETS, N_ACTIONS, OBS_SLICES

    net = copy.deepcopy(tr.net).double()
    for p in net.parameters():
        p.grad = None
    T, N = tr.T, tr.env.num_envs
    B = T * N
    dones = (tr.flags[:, :, 0:3] != 0).any(-1)
    ret, adv = A.gae_and_returns(tr.rewards.double(), tr.values.double(), dones, tr.gamma, tr.lamb)
    obs, masks, acts = tr.obs[:T].reshape(B, 38).double(), tr.masks[:T].reshape(B, 32), tr.actions.reshape(B, 8)
    ret, adv = ret.reshape(B, 8), adv.reshape(B, 8)
    adv_n = (adv - adv.mean(0)) / (adv.std(0) + 1e-8)
    if relu_masks is None:
        probs, v = net.probs32(obs), net.value(obs)
    else:
        m1, m2, m3 = (m.double() for m in relu_masks)

        def mlp(x, params, k, idx=None):
            w = [p if idx is None else p[idx] for p in params]
            h = (x @ w[0] + w[1]) * m1[k]
            h = (h @ w[2] + w[3]) * m2[k]
            return h, w

        z = torch.zeros(B, 32, dtype=torch.float64, device=obs.device)
        nets = [(0, net.ps, None), (1, net.agv, None)] + [(2 + i, net.six, i) for i in range(6)]
        parts = []
        for k, params, idx in nets:
            lo, hi = OBS_SLICES[k]
            h, w = mlp(obs[:, lo:hi], list(params), k, idx)
            parts.append(torch.softmax(h @ w[4] + w[5], -1))
        probs = torch.cat(parts + [torch.zeros(B, 3, dtype=torch.float64, device=obs.device)], dim=1)
        h, w = mlp(obs, list(net.critic), 8)
        h3 = (h @ w[4] + w[5]) * m3
        v = (h3 @ w[6] + w[7]).squeeze(-1)
    logp = A.log_prob_of(A.masked_policy(probs, masks), acts)
    # Categorical's clamp uses the fp32 epsilon in both trainers; in float64 it would be 2.2e-16: same gradients (zero
    # where the clamp is active in either precision: q = 1 exactly), log-probabilities equal to 1.2e-7
    loss = (-(adv_n * logp).mean(0) - tr.entropy_coef * A.entropy_unmasked(probs).mean(0)).sum()
    loss = loss + torch.nn.functional.mse_loss(v.unsqueeze(-1).expand_as(ret), ret)
    loss.backward()
    return [p.grad for p in net.parameters()]


@pytest.mark.parametrize("n_envs,T", [(512, 8), (1000, 5)])
def test_umma_update_equals_autograd_update(n_envs, T):
    """One update from the SAME rollout: analytic loss gradients + tcgen05 backward GEMMs vs torch autograd.
    (1) Against float64 autograd with the ReLU gates pinned to the ones the tensor-core forward produced, the tensor-core
    gradients are fp32-accurate (3xTF32).  (2) The gates themselves differ from the fp32 library forward in a vanishing
    fraction of units.  (3) Losses agree with the torch / cuBLAS fp32 trainer, gradients agree to the level gate flips
    allow, and both take the same Adam step wherever the gradient is not numerically zero."""
    a, b = _pair(n_envs, T)
    a.rollout()
    for name in ("obs", "masks", "actions", "rewards", "flags", "values"):
        getattr(b, name).copy_(getattr(a, name))
    eng, B = a.engine, a.T * a.env.num_envs
    gates = (eng.h1[:, :a.T].reshape(9, B, 256) > 0, eng.h2[:, :a.T].reshape(9, B, 256) > 0, eng.h3[:a.T].reshape(B, 128) > 0)
    g64 = _fp64_grads(a, gates)
    a._compute_grads()
    b._compute_grads()
    for k in ("actor_loss", "critic_loss", "entropy"):
        assert torch.allclose(a.stats[k], b.stats[k], rtol=2e-4, atol=1e-5), (k, a.stats[k], b.stats[k])
    for (na, pa), g in zip(a.net.named_parameters(), g64):
        scale = g.abs().max().item() + 1e-12
        err = (pa.grad.double() - g).abs().max().item()
        # (3xTF32 products are exact to ~2^-22; what is left is the tensor core's own fp32 accumulation over K, which cuts
        # rather than rounds: ~2e-5 of the largest gradient here, against ~1e-3 for single-pass TF32)
        assert err <= 1e-4 * scale, (na, err, scale)
    # gates: the fp32 library forward of the same weights
    with torch.no_grad():
        obs = a.obs[:a.T].reshape(B, 38)
        h1 = torch.relu(torch.addmm(b.net.agv[1][0], obs[:, 7:20], b.net.agv[0]))
        h2 = torch.relu(torch.addmm(b.net.agv[3][0], h1, b.net.agv[2]))
        flips = ((h1 > 0) != gates[0][1]).float().mean().item() + ((h2 > 0) != gates[1][1]).float().mean().item()
        assert flips < 1e-4, flips
    for (na, pa), (nb, pb) in zip(a.net.named_parameters(), b.net.named_parameters()):
        scale = pb.grad.abs().max().item() + 1e-12
        assert (pa.grad - pb.grad).abs().max().item() <= 2e-3 * scale, (na, (pa.grad - pb.grad).abs().max().item(), scale)
    a._clip(), a.opt.step()
    b._clip(), b.opt.step()
    for (na, pa), (nb, pb) in zip(a.net.named_parameters(), b.net.named_parameters()):
        # Adam's first step moves every weight by lr * g / (|g| + 1e-8): where the gradient is clearly non-zero the two
        # updates must coincide; where it is not, the step's sign is noise in both implementations
        sure = pb.grad.abs() > 1e-2 * pb.grad.abs().max()
        assert torch.allclose(pa[sure], pb[sure], rtol=1e-5, atol=2e-6), na


def test_umma_training_with_graphs_and_tf32_variant():
    """The tensor-core trainer end to end (CUDA-graph rollout and update): finite, parameters move, frames counted;
    plain-TF32 GEMMs (gemm_passes=1) also train."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv
    from multi_agent_rl_for_fjsp_b200.a2c_batched import BatchedA2C

    for passes in (3, 1):
        tr = BatchedA2C(BatchedFJSPEnv(512, seed=3, num_orders=25, autoreset=True), rollout_len=8, seed=2, gemm_passes=passes)
        assert tr.impl == "umma"
        p0 = torch.cat([p.detach().reshape(-1) for p in tr.net.parameters()]).clone()
        tr.train(5)
        p1 = torch.cat([p.detach().reshape(-1) for p in tr.net.parameters()])
        assert torch.isfinite(p1).all() and not torch.equal(p0, p1)
        assert torch.isfinite(tr.stats["critic_loss"]) and tr.frames == 5 * 8 * 512
        assert tr._ugraph is not None, tr.update_graph_error


def test_clip_adam_call_equals_torch_clipping_and_adam():
    """``fjsp_a2c_clip_adam`` (per-network clip_grad_norm_ 0.5 + Adam on the optimizer's own state, a2c.py:668,686-690) against
    ``clip_grad_norm_`` + ``torch.optim.Adam.step()`` on the same gradients, five steps: some networks clipped, some not, one
    with an all-zero gradient.  Parameters, both moments, the step counters and the clipped gradients must agree."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv
    from multi_agent_rl_for_fjsp_b200.a2c_batched import BatchedA2C

    a, b2 = (BatchedA2C(BatchedFJSPEnv(256, seed=3, num_orders=25, autoreset=True), rollout_len=4, seed=2, use_cuda_graph=False,
                        impl="umma", fused_optimizer=f) for f in (True, False))
    assert a.clip_adam is not None and b2.clip_adam is None
    b2.net.load_state_dict(a.net.state_dict())
    gen = torch.Generator(device="cuda").manual_seed(5)
    sizes = [sum(p.numel() for p in params) for _, params, _ in a.net.networks()]
    for it in range(5):
        g = torch.randn(a.engine.grad_flat.numel(), device="cuda", generator=gen)
        # per-network scales: pickup station far above the clip threshold, agv below it, the six mixed, critic above
        off = 0
        for sz, sc in zip(sizes, (1.0, 1e-5, 1e-3 * (it + 1), 0.1)):
            g[off:off + sz] *= sc
            off += sz
        a.engine.grad_flat.copy_(g), b2.engine.grad_flat.copy_(g)
        if it == 2:  # one stacked actor with a zero gradient: the clip coefficient is max_norm / 1e-6, clamped to 1
            for p in a.net.six:
                p.grad[3].zero_()
            for p in b2.net.six:
                p.grad[3].zero_()
        a.clip_adam.step()
        b2._clip(), b2.opt.step()
        assert torch.allclose(a.engine.grad_flat, b2.engine.grad_flat, rtol=1e-5, atol=1e-12), it
        for (na, pa), (_, pb) in zip(a.net.named_parameters(), b2.net.named_parameters()):
            sa, sb = a.opt.state[pa], b2.opt.state[pb]
            assert float(sa["step"]) == float(sb["step"]) == it + 1, (na, sa["step"], sb["step"])
            # (m = m + w (g - m) cancels where g ~ -m: absolute tolerances scaled to the tensor)
            for key in ("exp_avg", "exp_avg_sq"):
                ref = sb[key]
                assert torch.allclose(sa[key], ref, rtol=1e-5, atol=1e-6 * ref.abs().max().item()), (na, it, key)
            assert torch.allclose(pa, pb, rtol=1e-6, atol=2e-8), (na, it, (pa - pb).abs().max().item())
    assert float(a.clip_adam.norms_sq.abs().max()) == 0.0
    # the state is the optimizer's own: a torch step after ours continues from it
    sd = a.opt.state_dict()
    assert len(sd["state"]) == len(list(a.net.parameters()))


def test_single_step_forward_equals_the_rollout_passes():
    """``UmmaEngine.forward(t)`` (actors of step t + the critic on that step's rows only) reproduces what the rollout computed
    step by step (actors) and in one batched pass over all steps (critic): same kernels on the same rows, bit for bit."""
    a, _ = _pair(1000, 4)
    a.rollout()
    eng = a.engine
    for t in (0, 2, a.T):
        z, v = eng.logits[t].clone(), a.values[t].clone()
        if t < a.T:
            eng.logits[t].zero_()
        a.values[t].zero_()
        eng.forward(t)
        torch.cuda.synchronize()
        assert torch.equal(a.values[t], v), t
        if t < a.T:
            assert torch.equal(eng.logits[t], z), t
