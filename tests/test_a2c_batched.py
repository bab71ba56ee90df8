"""Batched A2C trainer (multi_agent_rl_for_fjsp_b200/a2c_batched.py) — torch logic on CPU:
unit properties, equivalence with the reference's a2c.py losses/gradients on identical data (needs the reference
tree), and 2-rank data parallelism over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_agent_rl_for_fjsp_b200 import a2c_batched as A
from multi_agent_rl_for_fjsp_b200.env import MASK_OFFSETS, N_ACTIONS


def test_parameter_count_matches_reference_checkpoint():
    assert A.ActorCritic(seed=0).num_parameters() == 654366  # checkpoints/model.pt


def test_masked_policy_and_sampling():
    g = torch.Generator().manual_seed(0)
    B = 4000
    probs = torch.zeros(B, 32)
    for off, n in zip(MASK_OFFSETS, N_ACTIONS):
        probs[:, off:off + n] = torch.softmax(torch.randn(B, n, generator=g), -1)
    masks = (torch.rand(B, 32, generator=g) < 0.6).to(torch.int8)
    for off in MASK_OFFSETS[:8]:
        masks[:, off] = 1  # action 0 is always valid in the env
    probs[:10, 3:11] = torch.tensor([1.0] + [0.0] * 7)  # all mass on action 0 ...
    masks[:10, 3] = 0                                   # ... which we mask out -> uniform-over-valid fallback
    masks[:10, 4:6] = 1
    q = A.masked_policy(probs, masks)
    for off, n in zip(MASK_OFFSETS, N_ACTIONS):
        assert torch.allclose(q[:, off:off + n].sum(-1), torch.ones(B), atol=1e-6)
        assert (q[:, off:off + n][masks[:, off:off + n] == 0] == 0).all()
    nvalid = masks[:10, 3:11].sum(-1, keepdim=True).float()
    assert torch.allclose(q[:10, 3:11], masks[:10, 3:11].float() / nvalid)
    acts = A.sample_actions(q, g)
    idx = acts.long() + torch.tensor(MASK_OFFSETS[:8])
    assert (masks.gather(1, idx) == 1).all(), "sampled a masked action"
    # empirical frequencies follow q for the AGV
    big = q[:1].repeat(20000, 1)
    a = A.sample_actions(big, g)[:, 1]
    freq = torch.bincount(a.long(), minlength=8).float() / 20000
    assert torch.allclose(freq, q[0, 3:11], atol=0.02)
    lp = A.log_prob_of(q, acts)
    assert torch.isfinite(lp).all() and (lp <= 0).all()


def test_gae_matches_python_loops_with_episode_ends():
    torch.manual_seed(1)
    T, N, gamma, lamb = 13, 3, 0.99, 0.95
    rewards, values = torch.randn(T, N, 8), torch.randn(T + 1, N)
    dones = torch.zeros(T, N, dtype=torch.bool)
    dones[4, 0] = dones[12, 1] = dones[7, 2] = dones[8, 2] = True
    ret, adv = A.gae_and_returns(rewards, values, dones, gamma, lamb)
    for n in range(N):
        for ag in range(8):
            # transition_memory.py:83-105 applied per episode segment
            r_ref, a_ref = [0.0] * T, [0.0] * T
            nxt_ret, nxt_val, gae = float(values[T, n]), float(values[T, n]), 0.0
            for t in range(T - 1, -1, -1):
                if dones[t, n]:
                    nxt_ret, nxt_val, gae = 0.0, 0.0, 0.0
                r = float(rewards[t, n, ag])
                nxt_ret = r + gamma * nxt_ret
                td = r + gamma * nxt_val - float(values[t, n])
                gae = td + gamma * lamb * gae
                r_ref[t], a_ref[t] = nxt_ret, gae
                nxt_val = float(values[t, n])
            assert np.allclose(ret[:, n, ag].numpy(), r_ref, atol=1e-5)
            assert np.allclose(adv[:, n, ag].numpy(), a_ref, atol=1e-5)


def _copy_weights_into_reference(net, ma2c, agent_ids):
    def put(seq, params, idx=None):
        lins = [m for m in seq if isinstance(m, torch.nn.Linear)]
        for k, lin in enumerate(lins):
            w, b = params[2 * k], params[2 * k + 1]
            if idx is not None:
                w, b = w[idx], b[idx]
            lin.weight.data.copy_(w.detach().t()), lin.bias.data.copy_(b.detach().reshape(-1))
    put(ma2c.actor_nets[agent_ids[0]].net, net.ps)
    put(ma2c.actor_nets[agent_ids[1]].net, net.agv)
    for i in range(6):
        put(ma2c.actor_nets[agent_ids[2 + i]].net, net.six, i)
    put(ma2c.critic_net.net, net.critic)


@pytest.mark.reference
def test_losses_and_gradients_equal_reference_a2c():
    """Same weights, same trajectory: per-agent actor losses, critic loss and gradients of the batched trainer equal
    those of the reference's MultiAgentA2C (a2c.py:647-731) computed by its own code."""
    import sys

    from oracle import canon, refload

    ns = refload.load_reference()
    sys.modules.setdefault("visualization", type(sys)("visualization")).GridVisualizer = object
    import importlib

    a2c_ref = importlib.import_module("a2c")
    torch.manual_seed(3)
    np.random.seed(3)
    env = ns.FJSPParallelEnv()
    ma = a2c_ref.MultiAgentA2C(env, batch_size=10**9, gamma=0.99, lamb=0.95, entropy_coef=0.01)  # train.py:63-74 values
    net = A.ActorCritic(seed=5)
    ids = env.possible_agents
    _copy_weights_into_reference(net, ma, ids)
    T = 48
    with refload.quiet():
        obs, _ = env.reset(options={"num_orders": 25})
    O, M, ACT, R = [], [], [], []
    for t in range(T):
        o, m = canon.flatten_reference_obs(obs)
        actions, logprobs, values = ma.predict(obs, env.agents, train_returns=True)
        with refload.quiet():
            nobs, rew, te, tr, _ = env.step(actions)
        ma.memory.put(obs, actions, rew, logprobs, values)
        O.append(o), M.append(m), ACT.append([actions[a] for a in ids]), R.append([rew[a] for a in ids])
        obs = nobs
    o_last, _ = canon.flatten_reference_obs(obs)
    boot = ma.critic_net(torch.FloatTensor(ma._get_global_state(obs, env.agents))).item()
    ma.memory.finish_trajectory({a: boot for a in ids})
    # reference losses / grads (no optimiser step)
    ref_actor, ref_grad0 = [], []
    for a in ids:
        obs_lst, _, _, logprob_lst, return_lst, value_lst, adv_lst = ma.memory.get(a)
        loss = ma.calc_actor_loss(logprob_lst, adv_lst) - ma.entropy_coef * ma._calculate_entropy(a, obs_lst)
        ma.actor_nets[a].zero_grad()
        loss.backward(retain_graph=True)
        ref_actor.append(loss.item())
        ref_grad0.append(ma.actor_nets[a].net[0].weight.grad.clone())
    all_v, all_r = [], []
    for a in ids:
        _, _, _, _, return_lst, value_lst, _ = ma.memory.get(a)
        all_v.extend(value_lst), all_r.extend(return_lst)
    closs = ma.calc_critic_loss(all_v, all_r)
    ma.critic_net.zero_grad()
    closs.backward()
    ref_cgrad = ma.critic_net.net[0].weight.grad.clone()

    # batched trainer on the same data (N = 1)
    rewards = torch.tensor(R, dtype=torch.float32).reshape(T, 1, 8)
    obs_t = torch.tensor(np.array(O + [o_last]), dtype=torch.float32).reshape(T + 1, 1, 38)
    masks = torch.tensor(np.array(M), dtype=torch.int8)
    acts = torch.tensor(ACT, dtype=torch.uint8)
    with torch.no_grad():
        values = net.value(obs_t.reshape(T + 1, 38)).reshape(T + 1, 1)
    assert abs(float(values[T, 0]) - boot) < 1e-5
    ret, adv = A.gae_and_returns(rewards, values, torch.zeros(T, 1, dtype=torch.bool), 0.99, 0.95)
    advs = adv.reshape(T, 8)
    adv_n = (advs - advs.mean(0)) / (advs.std(0) + 1e-8)
    probs = net.probs32(obs_t[:T].reshape(T, 38))
    q = A.masked_policy(probs, masks)
    logp = A.log_prob_of(q, acts)
    actor_loss = -(adv_n * logp).mean(0) - 0.01 * A.entropy_unmasked(probs).mean(0)
    critic_loss = torch.nn.functional.mse_loss(net.value(obs_t[:T].reshape(T, 38)).unsqueeze(-1).expand(T, 8), ret.reshape(T, 8))
    (actor_loss.sum() + critic_loss).backward()
    assert np.allclose(actor_loss.detach().numpy(), ref_actor, rtol=2e-4, atol=2e-5), (actor_loss, ref_actor)
    assert abs(critic_loss.item() - closs.item()) <= 2e-4 * abs(closs.item())
    mine0 = [net.ps[0].grad.t(), net.agv[0].grad.t()] + [net.six[0].grad[i].t() for i in range(6)]
    for g_m, g_r in zip(mine0, ref_grad0):
        assert torch.allclose(g_m, g_r, rtol=1e-3, atol=1e-5)
    assert torch.allclose(net.critic[0].grad.t(), ref_cgrad, rtol=1e-3, atol=1e-5)


def _dp_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo")
    from tests.fake_tensor_env import FakeTensorEnv

    torch.set_num_threads(1)
    env = FakeTensorEnv(2, first_env=2 * rank, seed=9)
    tr = A.BatchedA2C(env, rollout_len=6, seed=4)
    tr.gen.manual_seed(100 + rank)
    tr.rollout()
    shard = {k: getattr(tr, k).clone() for k in ("obs", "masks", "actions", "rewards", "flags", "values")}
    tr.update()
    params = torch.cat([p.detach().reshape(-1) for p in tr.net.parameters()])
    gathered = [None] * world
    dist.all_gather_object(gathered, (shard, params))
    if rank == 0:  # plain numpy through the queue (torch tensors would travel as fds of a process that exits)
        q.put([({k: v.numpy() for k, v in sh.items()}, pr.numpy()) for sh, pr in gathered])
    dist.destroy_process_group()


def test_two_rank_data_parallel_update_equals_single_process_on_joint_batch():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (s0, p0), (s1, p1) = [({k: torch.from_numpy(v) for k, v in sh.items()}, torch.from_numpy(pr)) for sh, pr in gathered]
    assert torch.equal(p0, p1), "replicas diverged"

    class Joint:
        num_envs, device = 4, torch.device("cpu")

        def reset(self):
            return torch.zeros(4, 38), torch.zeros(4, 32, dtype=torch.int8)

    tr = A.BatchedA2C(Joint(), rollout_len=6, seed=4)
    for k in s0:
        getattr(tr, k).copy_(torch.cat([s0[k], s1[k]], dim=1))
    tr.update()
    pj = torch.cat([p.detach().reshape(-1) for p in tr.net.parameters()])
    # one Adam step moves a parameter by ~lr = 3e-4; summation-order noise on near-zero gradients stays far below that
    assert torch.allclose(pj, p0, rtol=1e-4, atol=2e-5), float((pj - p0).abs().max())


def test_checkpoint_roundtrip_in_reference_layout(tmp_path):
    a, b = A.ActorCritic(seed=1), A.ActorCritic(seed=2)
    a.save_reference_checkpoint(str(tmp_path / "m.pt"))
    ck = b.load_reference_checkpoint(str(tmp_path / "m.pt"))
    assert ck["global_obs_dim"] == 38 and ck["obs_dims"]["agv"] == 13 and ck["act_dims"]["agv"] == 8
    o = torch.randn(7, 38)
    assert torch.equal(a.probs32(o), b.probs32(o)) and torch.equal(a.value(o), b.value(o))
    assert set(ck["actor_nets"]["agv"]) == {"net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias", "net.4.weight", "net.4.bias"}
    assert ck["actor_nets"]["agv"]["net.0.weight"].shape == (256, 13)


@pytest.mark.reference
def test_loads_reference_checkpoint_and_matches_reference_networks():
    """checkpoints/model.pt of the reference (654,366 fp32 params) loads with weights_only=True and the batched
    networks reproduce the reference's nn.Sequential outputs."""
    import importlib
    import sys

    from oracle import refload

    refload.load_reference()
    networks = importlib.import_module("networks")
    path = refload.REFERENCE_ROOT + "/checkpoints/model.pt"
    net = A.ActorCritic(seed=0)
    ck = net.load_reference_checkpoint(path)
    assert list(ck["possible_agents"])[1] == "agv" and int(ck["global_obs_dim"]) == 38
    o = torch.rand(16, 38) * 3
    pr = net.probs32(o)
    agv = networks.ActorNetwork(13, 8)
    agv.load_state_dict(ck["actor_nets"]["agv"])
    red = networks.ActorNetwork(3, 3)
    red.load_state_dict(ck["actor_nets"]["packaging_red"])
    crit = networks.CentralizedCriticNetwork(38)
    crit.load_state_dict(ck["critic_net"])
    assert torch.allclose(pr[:, 3:11], agv(o[:, 7:20]), atol=1e-6)
    assert torch.allclose(pr[:, 23:26], red(o[:, 32:35]), atol=1e-6)
    assert torch.allclose(net.value(o), crit(o).squeeze(-1), atol=1e-5)
