"""Scaled shop (K = 2..4 cells, include/fjsp_b200.h) on the CUDA path, through the C ABI, against the C restatement.

The reference has no K-cell shop, so the anchors are the ones of tests/test_scaled_shop.py: K = 1 is the reference shop
bit for bit (every other GPU test), the packed-state core == the oracle for K = 2..4 on the CPU, and here the kernels ==
the oracle on the same seeded Philox streams (observations, masks, flags, action results, canonical state of every cell
bit-exact; rewards within REL_TOL), the K-steps-per-launch kernel == the step kernel on the whole packed state, and the
host-buffer path == the device path.
"""
import numpy as np
import pytest
import torch

from oracle import canon
from oracle.fjsp_oracle import OracleBatch, OracleEnv, default_config, dims, philox_actions, philox_orders
from tests.test_gpu_parity import _abi_cfg
from tests.util import REL_TOL

pytestmark = pytest.mark.gpu


def ocfg_k(k, **kw):
    c = default_config()
    c.num_cells = k
    for key, v in kw.items():
        setattr(c, key, v)
    return c


@pytest.mark.parametrize("k,n_envs,first_env", [(2, 200, 7), (3, 129, 0), (4, 64, 0), (4, 1000, 999)])
def test_gpu_scaled_vs_oracle_random_batch(k, n_envs, first_env):
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    seed, num_orders, steps = 0xABCDEF + k, 30, 420
    ocfg = ocfg_k(k)
    env = BatchedFJSPEnv(n_envs, config=_abi_cfg(ocfg), first_env=first_env, seed=seed, num_orders=num_orders, autoreset=True,
                         with_infos=True)
    d = dims(k)
    assert (env.act_dim, env.obs_dim, env.mask_dim) == (d["act"], d["obs"], d["mask"])
    assert env.state_bytes_per_env == 4 * (64 + 64 * k + 20 * (k - 1)) and len(env.agent_ids) == 1 + 7 * k
    obs0, masks0 = env.reset()
    obs0, masks0 = obs0.cpu().numpy(), masks0.cpu().numpy()
    rs = np.random.RandomState(n_envs + k)
    sample = sorted(set([0, n_envs - 1] + rs.randint(0, n_envs, size=12).tolist()))
    oracles, episodes = {}, {}
    for i in sample:
        o = OracleEnv(ocfg)
        oo, om = o.reset(philox_orders(seed, first_env + i, 0, num_orders))
        assert np.array_equal(oo, obs0[i]) and np.array_equal(om, masks0[i])
        oracles[i], episodes[i] = o, 0
    for t in range(steps):
        acts = env.random_actions(t)
        obs, rew, term, trunc, masks = env.step(acts)
        h_act, obs, rew, masks, flags, results = (x.cpu().numpy() for x in (acts, obs, rew, masks, env.flags, env.results))
        for i in sample:
            o = oracles[i]
            assert np.array_equal(h_act[i], philox_actions(seed, first_env + i, t, cells=k)), "action stream"
            oo, om, orw, of = o.step(h_act[i])
            assert tuple(of[:3]) == tuple(flags[i][:3]), (i, t, of, flags[i])
            assert np.all(np.abs(rew[i] - orw) <= REL_TOL * np.abs(orw)), (i, t, rew[i], orw)
            assert np.array_equal(o.results, results[i]), (i, t)
            if of[0] or of[1] or of[2]:
                assert flags[i][3] == 1
                episodes[i] += 1
                oo, om = o.reset(philox_orders(seed, first_env + i, episodes[i], num_orders))
            assert np.array_equal(oo, obs[i]), (i, t, np.flatnonzero(oo != obs[i]))
            assert np.array_equal(om, masks[i]), (i, t, np.flatnonzero(om != masks[i]))
        if t % 60 == 59:
            for i in sample[:4]:
                for c in range(k):
                    df = canon.diff(oracles[i].export(c), env.export_state(i, c))
                    assert not df, (i, t, c, df[:4])
    assert max(episodes.values()) >= 2


@pytest.mark.parametrize("k", [2, 4])
def test_gpu_scaled_rollout_matches_stepwise_and_oracle(k):
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    n, seed, num_orders, steps = 640 + 17, 31 + k, 30, 230
    ocfg = ocfg_k(k)
    a = BatchedFJSPEnv(n, config=_abi_cfg(ocfg), seed=seed, num_orders=num_orders, autoreset=True)
    b = BatchedFJSPEnv(n, config=_abi_cfg(ocfg), seed=seed, num_orders=num_orders, autoreset=True)
    a.reset(), b.reset()
    for t in range(steps):
        a.step(a.random_actions(t))
    b.rollout_random(70, t0=0)
    stats = b.rollout_random(steps - 70, t0=70).cpu().numpy()
    assert torch.equal(a.save_state(), b.save_state())
    ob = OracleBatch(n, seed, num_orders, cfg=ocfg)
    ostats = ob.rollout(steps, nthreads=4)
    assert stats[:6].tolist() == ostats[:6].astype(np.int64).tolist(), (stats, ostats)


def test_gpu_scaled_heuristic_cells_complete_orders():
    """A competent per-cell policy on the device tensors: trays flow through every cell, orders complete, episodes
    terminate at different steps (partial-ballot cooperative resets) — sampled envs followed by the oracle."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv
    from tests.test_gpu_parity import _torch_heuristic

    k, n, seed, num_orders, steps = 4, 700, 2468, 6, 300
    ocfg = ocfg_k(k)
    env = BatchedFJSPEnv(n, config=_abi_cfg(ocfg), seed=seed, num_orders=num_orders, autoreset=True)
    obs, masks = env.reset()
    gen = torch.Generator(device=env.device).manual_seed(11)
    sample = sorted(set([0, 31, 32, 63, 64, n - 1] + np.random.RandomState(2).randint(0, n, size=20).tolist()))
    oracles, episodes = {}, {}
    for i in sample:
        o = OracleEnv(ocfg)
        o.reset(philox_orders(seed, i, 0, num_orders))
        oracles[i], episodes[i] = o, 0
    terminated = 0
    for t in range(steps):
        acts = torch.zeros(n, env.act_dim, dtype=torch.uint8, device=env.device)
        for c in range(k):  # the single-shop policy applied to each cell's view (pickup station + the cell's 7 agents)
            o_c = torch.cat([obs[:, :7], obs[:, 7 + 31 * c:38 + 31 * c]], dim=1)
            m_c = torch.zeros(n, 32, dtype=torch.int8, device=env.device)
            m_c[:, :3], m_c[:, 3:29] = masks[:, :3], masks[:, 3 + 26 * c:29 + 26 * c]
            a_c = _torch_heuristic(o_c, m_c, gen)
            if c == 0:
                acts[:, 0] = a_c[:, 0]
            acts[:, 1 + 7 * c:8 + 7 * c] = a_c[:, 1:]
        obs, rew, term, trunc, masks = env.step(acts)
        h_act, h_obs, h_rew, h_masks, h_flags = (x.cpu().numpy() for x in (acts, obs, rew, masks, env.flags))
        terminated += int(h_flags[:, 0].sum())
        for i in sample:
            o = oracles[i]
            oo, om, orw, of = o.step(h_act[i])
            assert tuple(of[:3]) == tuple(h_flags[i][:3]), (i, t, of, h_flags[i])
            assert np.all(np.abs(h_rew[i] - orw) <= REL_TOL * np.abs(orw)), (i, t)
            if of[0] or of[1] or of[2]:
                episodes[i] += 1
                oo, om = o.reset(philox_orders(seed, i, episodes[i], num_orders))
            assert np.array_equal(oo, h_obs[i]) and np.array_equal(om, h_masks[i]), (i, t)
    assert terminated > n // 2, terminated
    for i in sample:
        for c in range(k):
            df = canon.diff(oracles[i].export(c), env.export_state(i, c))
            assert not df, (i, c, df[:4])


@pytest.mark.parametrize("k,n", [(3, 333), (4, 20001), (2, 4097)])
def test_gpu_scaled_host_buffer_step_matches_device_step(k, n):
    """Wire rows of a K-cell shop through the pipelined host path (ragged last tile; with and without decode workers)."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    seed = 5
    cfg = _abi_cfg(ocfg_k(k))
    a = BatchedFJSPEnv(n, config=cfg, seed=seed)
    b = BatchedFJSPEnv(n, config=cfg, seed=seed)
    a.reset(), b.reset()
    for t in range(30):
        acts = a.random_actions(t)
        obs, rew, term, trunc, masks = a.step(acts)
        hobs, hmasks, hrew, hflags = b.step_host(acts.cpu().numpy())
        assert np.array_equal(obs.cpu().numpy(), hobs) and np.array_equal(masks.cpu().numpy(), hmasks)
        assert np.array_equal(rew.cpu().numpy(), hrew) and np.array_equal(a.flags.cpu().numpy(), hflags)


def test_gpu_scaled_shard_map_invariance_at_size():
    """2^17 four-cell shops (190 MB of state): env g of the full batch == env g - first of a shard created with first_env."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    k, n, seed, steps = 4, 1 << 17, 77, 40
    cfg = _abi_cfg(ocfg_k(k))
    full = BatchedFJSPEnv(n, config=cfg, seed=seed)
    full.reset()
    stats = full.rollout_random(steps, t0=0).cpu().numpy()
    assert stats[0] == n * steps
    first = n // 2 + 321
    shard = BatchedFJSPEnv(512, config=cfg, first_env=first, seed=seed)
    shard.reset()
    for t in range(steps):
        shard.step(shard.random_actions(t))
    for g in (first, first + 77, first + 511):
        assert np.array_equal(full.export_packed(g), shard.export_packed(g - first))


def test_gpu_cell_views_match_plain_slicing():
    """fjsp_cells_unpack_views / fjsp_cells_pack_actions against plain torch slicing of the same tensors."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv, CellViewEnv

    for k in (2, 4):
        n = 257
        env = BatchedFJSPEnv(n, config=_abi_cfg(ocfg_k(k)), seed=3)
        view = CellViewEnv(env)
        vo, vm = view.reset()
        gen = torch.Generator(device=env.device).manual_seed(k)
        for t in range(12):
            va = (torch.rand(n * k, 8, device=env.device, generator=gen) * torch.tensor([3, 8, 3, 3, 3, 3, 3, 3], device=env.device)).to(torch.uint8)
            vo, vr, term, trunc, vm = view.step(va)
            a = view._actions.cpu().numpy()
            van = va.cpu().numpy().reshape(n, k, 8)
            assert np.array_equal(a[:, 0], van[:, 0, 0]) and (a[:, 1 + 7 * k:] == 0).all()
            o, m, r, f = (x.cpu().numpy() for x in (env.obs, env.masks, env.rewards, env.flags))
            vo_n, vm_n, vr_n, vf_n = (x.cpu().numpy() for x in (vo, vm, vr, view.flags))
            for c in range(k):
                assert np.array_equal(a[:, 1 + 7 * c:8 + 7 * c], van[:, c, 1:])
                rows = np.arange(n) * k + c
                assert np.array_equal(vo_n[rows, :7], o[:, :7]) and np.array_equal(vo_n[rows, 7:], o[:, 7 + 31 * c:38 + 31 * c])
                want_ps = m[:, :3] if c == 0 else np.tile(np.array([1, 0, 0], np.int8), (n, 1))
                assert np.array_equal(vm_n[rows, :3], want_ps) and np.array_equal(vm_n[rows, 3:29], m[:, 3 + 26 * c:29 + 26 * c])
                assert (vm_n[rows, 29:] == 0).all()
                assert np.array_equal(vr_n[rows, 0], r[:, 0]) and np.array_equal(vr_n[rows, 1:], r[:, 1 + 7 * c:8 + 7 * c])
                assert np.array_equal(vf_n[rows], f)


def test_gpu_a2c_trains_on_the_scaled_shop():
    """BatchedA2C on a 4-cell shop through CellViewEnv: CUDA-graph rollout + update run, losses finite, the dummy pickup
    rows only ever idle, and the networks are the reference's (654,366 parameters, shared by the cells)."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv, CellViewEnv
    from multi_agent_rl_for_fjsp_b200.a2c_batched import BatchedA2C

    k, n, T = 4, 512, 16
    env = BatchedFJSPEnv(n, config=_abi_cfg(ocfg_k(k)), seed=9, num_orders=25, autoreset=True)
    tr = BatchedA2C(CellViewEnv(env), rollout_len=T, seed=2)
    assert tr.net.num_parameters() == 654366
    before = [p.detach().clone() for p in tr.net.parameters()]
    fps, secs = tr.train(5)
    torch.cuda.synchronize()
    assert tr.update_graph_error is None
    assert torch.isfinite(tr.stats["critic_loss"]).all() and torch.isfinite(tr.stats["actor_loss"]).all()
    acts = tr.actions.reshape(T, n, k, 8)
    assert (acts[:, :, 1:, 0] == 0).all() and (acts[:, :, 0, 0] != 0).any()
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, tr.net.parameters()))
    assert env.launch_count >= T   # eager warm-up pass; later rollouts replay the captured graph


@pytest.mark.parametrize("name", ["k2_mixed", "k3_pack_cap3", "k4_heuristic"])
def test_gpu_replays_scaled_vectors(name):
    """The scaled-shop regression vectors (tests/golden/scaled/) through fjsp_reset / fjsp_step / fjsp_export_state_cell."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv
    from tests.test_scaled_golden import replay_scaled

    class GpuSingleK:
        def __init__(self, ocfg):
            self.env = BatchedFJSPEnv(1, config=_abi_cfg(ocfg), autoreset=False)

        def reset(self, orders):
            obs, masks = self.env.reset(orders=np.asarray(orders)[None, :, :])
            return obs[0].cpu().numpy(), masks[0].cpu().numpy()

        def step(self, actions):
            a = torch.as_tensor(np.asarray(actions, dtype=np.uint8)[None, :], device=self.env.device)
            obs, rew, term, trunc, masks = self.env.step(a)
            return obs[0].cpu().numpy(), masks[0].cpu().numpy(), rew[0].cpu().numpy(), self.env.flags[0].cpu().numpy()

        def export(self, cell=0):
            return self.env.export_state(0, cell)

    assert replay_scaled(name, lambda cfg: GpuSingleK(cfg), exact_rewards=False) > 300


def test_gpu_random_config_fuzz_all_cell_counts():
    """Random valid configurations (layouts with multi-step AGV moves, timings, capacities, episode lengths, few trays)
    x K in {1, 2, 3, 4}: a batch of 96 envs with Philox orders / actions and auto-reset on the GPU, sampled envs followed
    by the restatement — observations, masks, flags bit-exact, rewards within REL_TOL, canonical state of every cell."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    rs = np.random.RandomState(77)
    checked = 0
    for trial in range(10):
        k = [1, 2, 3, 4][trial % 4]
        step = int(rs.choice([5, 10, 20]))
        cells = set()
        while len(cells) < 5:
            cells.add((int(rs.randint(0, 12)), int(rs.randint(0, 30))))
        ocfg = ocfg_k(k, step_size=step, proc_small=step * int(rs.randint(1, 8)), proc_big=step * int(rs.randint(1, 13)),
                      proc_pack=step * int(rs.randint(1, 5)), agv_speed=int(rs.randint(1, 3)),
                      max_episode_steps=int(rs.randint(40, 241)), storage_capacity=int(rs.randint(0, 6)),
                      pack_capacity=int(rs.randint(1, 32)), num_trays=int(rs.choice([3, 40, 1000])))
        for i, (r, c) in enumerate(sorted(cells, key=lambda _: rs.rand())):
            ocfg.pos[i][0], ocfg.pos[i][1] = r, c
        n, seed, num_orders, steps = 96, 1000 + trial, int(rs.randint(1, 33)), 300
        env = BatchedFJSPEnv(n, config=_abi_cfg(ocfg), seed=seed, num_orders=num_orders, autoreset=True, first_env=trial * 1000)
        obs0, masks0 = env.reset()
        sample = [0, 31, 64, 95]
        oracles, episodes = {}, {}
        for i in sample:
            o = OracleEnv(ocfg)
            oo, om = o.reset(philox_orders(seed, trial * 1000 + i, 0, num_orders))
            assert np.array_equal(oo, obs0[i].cpu().numpy()) and np.array_equal(om, masks0[i].cpu().numpy()), trial
            oracles[i], episodes[i] = o, 0
        for t in range(steps):
            acts = env.random_actions(t)
            obs, rew, term, trunc, masks = env.step(acts)
            h_act, obs, rew, masks, flags = (x.cpu().numpy() for x in (acts, obs, rew, masks, env.flags))
            for i in sample:
                o = oracles[i]
                oo, om, orw, of = o.step(h_act[i])
                assert tuple(of[:3]) == tuple(flags[i][:3]), (trial, k, i, t, of, flags[i])
                assert np.all(np.abs(rew[i] - orw) <= REL_TOL * np.abs(orw)), (trial, k, i, t)
                if of[0] or of[1] or of[2]:
                    episodes[i] += 1
                    oo, om = o.reset(philox_orders(seed, trial * 1000 + i, episodes[i], num_orders))
                assert np.array_equal(oo, obs[i]) and np.array_equal(om, masks[i]), (trial, k, i, t)
                checked += 1
        for i in sample:
            for c in range(k):
                df = canon.diff(oracles[i].export(c), env.export_state(i, c))
                assert not df, (trial, k, i, c, df[:4])
    assert checked == 10 * 300 * 4
