"""SHARED FLOOR on the device (include/fjsp_b200.h "shared floor"): fjsp_shared_step_kernel<A> through the
C ABI against the CPU restatement and the regression vectors of tests/golden/shared/ (CPU side: tests/test_shared_floor.py)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import canon
from oracle.fjsp_oracle import OracleEnv, default_config, philox_actions_shared, philox_orders
from tests.test_shared_floor import SHARED_FILES, digest, load_shared
from tests.util import REL_TOL

pytestmark = pytest.mark.gpu


def _abi_cfg(ocfg):
    from multi_agent_rl_for_fjsp_b200 import abi

    a = abi.FjspConfig()
    C.memmove(C.addressof(a), C.addressof(ocfg), C.sizeof(a))
    return a


@pytest.mark.parametrize("agvs,n_envs,num_orders", [(2, 700, 30), (3, 333, 25), (4, 1000, 32), (4, 65, 6)])
def test_gpu_shared_batch_follows_the_restatement(agvs, n_envs, num_orders):
    """Philox orders and actions, auto-reset, ragged batch sizes: sampled envs are followed by the restatement step by step
    — observations, masks, rewards, flags, action results, every AGV's canonical record."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    ocfg = default_config()
    ocfg.shared_agvs = agvs
    seed, first_env = 0xFACE0000 + agvs, 7000
    env = BatchedFJSPEnv(n_envs, config=_abi_cfg(ocfg), first_env=first_env, seed=seed, num_orders=num_orders, autoreset=True, with_infos=True)
    assert env.dims["agents"] == 7 + agvs and len(env.agent_ids) == 7 + agvs and env.state_bytes_per_env == 528
    obs0, masks0 = env.reset()
    rs = np.random.RandomState(n_envs)
    sample = sorted(set([0, n_envs - 1, 63, 64] + rs.randint(0, n_envs, size=10).tolist()))
    oracles, episodes = {}, {}
    obs0, masks0 = obs0.cpu().numpy(), masks0.cpu().numpy()
    for i in sample:
        o = OracleEnv(ocfg)
        oo, om = o.reset(philox_orders(seed, first_env + i, 0, num_orders))
        assert np.array_equal(oo, obs0[i]) and np.array_equal(om, masks0[i])
        oracles[i], episodes[i] = o, 0
    A = 7 + agvs
    for t in range(450):
        acts = env.random_actions(t)
        ah = acts.cpu().numpy()
        obs, rew, term, trunc, masks = env.step(acts)
        obs, rew, masks, flags, infos = (x.cpu().numpy() for x in (obs, rew, masks, env.flags, env.infos))
        results = env.results.cpu().numpy()
        for i in sample:
            o = oracles[i]
            assert np.array_equal(ah[i], philox_actions_shared(seed, first_env + i, t, agvs)), "action stream"
            oo, om, orw, of = o.step(ah[i])
            assert tuple(of[:3]) == tuple(flags[i][:3]), (i, t, of, flags[i])
            assert np.all(np.abs(rew[i][:A] - orw[:A]) <= REL_TOL * np.abs(orw[:A])) and not rew[i][A:].any(), (i, t)
            assert np.array_equal(o.results, results[i]), (i, t)
            if of[0] or of[1] or of[2]:
                assert flags[i][3] == 1
                episodes[i] += 1
                oo, om = o.reset(philox_orders(seed, first_env + i, episodes[i], num_orders))
            else:
                assert flags[i][3] == 0 and infos[i][0] == int(o.export()["current_step"])
            assert np.array_equal(oo, obs[i]), (i, t, np.flatnonzero(oo != obs[i]))
            assert np.array_equal(om, masks[i]), (i, t)
        if t % 50 == 0 or t == 449:
            for i in sample[:6]:
                for j in range(agvs):
                    d = canon.diff(oracles[i].export(j), env.export_state(i, j))
                    assert not d, (i, t, j, d[:4])
    assert min(episodes.values()) >= 2   # auto-reset happened (200-step episodes)


@pytest.mark.parametrize("name", SHARED_FILES)
def test_gpu_replays_shared_vectors(name):
    """tests/golden/shared/ through fjsp_reset / fjsp_step / fjsp_export_state_cell: every episode of a file is one env of
    a batch, stepped in lockstep with explicit order tables."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    g, ocfg = load_shared(name)
    agvs = int(g["agvs"])
    starts = g["ep_start"].tolist() + [g["actions"].shape[0]]
    E = len(starts) - 1
    env = BatchedFJSPEnv(E, config=_abi_cfg(ocfg), autoreset=False, with_infos=True)
    # one reset per distinct order count (the handle takes one num_orders per call; masked resets for the others)
    for no in sorted(set(int(x) for x in g["ep_norders"])):
        sel = np.array([int(g["ep_norders"][e]) == no for e in range(E)])
        orders = np.zeros((E, 32), dtype=np.uint32)
        for e in range(E):
            t = g["ep_orders"][e]
            orders[e] = t[:, 0] | (t[:, 1] << 8) | (t[:, 2] << 16)
            orders[e, int(g["ep_norders"][e]):] = 0
        obs, masks = env.reset(num_orders=no, orders=orders, env_mask=torch.as_tensor(sel, device=env.device))
    obs, masks = obs.cpu().numpy(), masks.cpu().numpy()
    last_no = int(sorted(set(int(x) for x in g["ep_norders"]))[-1])
    for e in range(E):
        if int(g["ep_norders"][e]) == last_no:
            assert np.array_equal(obs[e], g["ep_obs0"][e]) and np.array_equal(masks[e], g["ep_masks0"][e])
    lens = [starts[e + 1] - starts[e] for e in range(E)]
    A = 7 + agvs
    for k in range(max(lens)):
        acts = np.zeros((E, env.act_dim), dtype=np.uint8)
        live = [e for e in range(E) if k < lens[e]]
        for e in live:
            acts[e] = g["actions"][starts[e] + k]
        obs, rew, term, trunc, masks = env.step(torch.as_tensor(acts, device=env.device))
        obs, rew, masks, flags, results = (x.cpu().numpy() for x in (obs, rew, masks, env.flags, env.results))
        for e in live:
            t = starts[e] + k
            assert np.array_equal(obs[e], g["obs"][t]), (name, e, k, np.flatnonzero(obs[e] != g["obs"][t]))
            assert np.array_equal(masks[e], g["masks"][t]), (name, e, k)
            assert np.all(np.abs(rew[e][:A] - g["rewards"][t][:A]) <= REL_TOL * np.abs(g["rewards"][t][:A])), (name, e, k)
            assert tuple(flags[e][:3]) == tuple(g["flags"][t]), (name, e, k)
            assert np.array_equal(results[e], g["results"][t]), (name, e, k)
            if k % 25 == 0 or k == lens[e] - 1:
                assert [digest(env.export_state(e, j)) for j in range(agvs)] == g["hashes"][t].tolist(), (name, e, k)


def test_gpu_shared_entry_points_and_state_round_trip():
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    ocfg = default_config()
    ocfg.shared_agvs = 3
    a = BatchedFJSPEnv(300, config=_abi_cfg(ocfg), seed=5, num_orders=20, autoreset=True)
    b = BatchedFJSPEnv(300, config=_abi_cfg(ocfg), seed=5, num_orders=20, autoreset=True)
    a.reset(), b.reset()
    for t in range(40):
        a.step(a.random_actions(t))
    snap = a.save_state()
    b.load_state(snap)
    assert np.array_equal(a.export_packed(7), b.export_packed(7)) and a.export_packed(7).shape == (132,)
    for t in range(40, 60):
        acts = a.random_actions(t)
        oa, ra, _, _, ma = a.step(acts)
        ob, rb, _, _, mb = b.step(acts)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(ma, mb)
    assert torch.equal(a.save_state(), b.save_state())
    # what the shared floor does not offer refuses loudly
    with pytest.raises(RuntimeError, match="shared floor"):
        a.rollout_random(4)
    with pytest.raises(RuntimeError, match="shared floor"):
        a.step_host(np.zeros((300, a.act_dim), np.uint8))
    with pytest.raises(Exception):
        a.export_state(0, 3)   # AGV index out of range


@pytest.mark.skipif(not __import__("os").environ.get("FJSP_SOAK"), reason="long run: set FJSP_SOAK=1 (evidence: profiles/r02_soak_shared_long.jsonl)")
@pytest.mark.parametrize("mode", ["shared4", "shared2", "long4"])
def test_gpu_soak_new_layouts_follow_the_restatement(mode):
    """2^16 envs x 10,000 steps with auto-reset (6.6*10^8 env-steps per mode), 48 sampled envs followed by the restatement
    at EVERY step: the shared floor (A = 4, 2) and the 4-cell shop on long order streams with Philox arrivals."""
    import json
    import os

    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv
    from oracle.fjsp_oracle import philox_actions

    ocfg = default_config()
    if mode.startswith("shared"):
        ocfg.shared_agvs = int(mode[-1])
        cells, num_orders, table_len = 1, 30, 30
        acts_of = lambda seed, g, t: philox_actions_shared(seed, g, t, ocfg.shared_agvs)  # noqa: E731
    else:
        ocfg.num_cells, ocfg.long_streams, ocfg.max_episode_steps, ocfg.arrival_prob_q16, ocfg.arrival_max_orders = 4, 1, 700, 20000, 220
        cells, num_orders, table_len = 4, 8, 220
        acts_of = lambda seed, g, t: philox_actions(seed, g, t, 4)  # noqa: E731
    n, steps, seed, first_env = 1 << 16, 10000, 0xC0FFEE, 123456
    env = BatchedFJSPEnv(n, config=_abi_cfg(ocfg), first_env=first_env, seed=seed, num_orders=num_orders, autoreset=True, with_infos=True)
    obs0, masks0 = env.reset()
    rs = np.random.RandomState(7)
    sample = sorted(set([0, n - 1] + rs.randint(0, n, size=46).tolist()))
    idx = torch.as_tensor(sample, device=env.device)
    oracles, episodes = {}, {}

    def fresh(o, i, ep):
        orders = philox_orders(seed, first_env + i, ep, table_len)
        return o.reset_stream(orders, num_orders, seed, first_env + i, ep) if ocfg.long_streams else o.reset(orders)

    for i in sample:
        o = OracleEnv(ocfg)
        oo, om = fresh(o, i, 0)
        assert np.array_equal(oo, obs0[i].cpu().numpy()) and np.array_equal(om, masks0[i].cpu().numpy())
        oracles[i], episodes[i] = o, 0
    A = env.dims["agents"]
    faults = ends = 0
    for t in range(steps):
        acts = env.random_actions(t)
        obs, rew, term, trunc, masks = env.step(acts)
        ho, hr, hm, hf, ha = (x[idx].cpu().numpy() for x in (obs, rew, masks, env.flags, acts))
        for j, i in enumerate(sample):
            o = oracles[i]
            assert np.array_equal(ha[j], acts_of(seed, first_env + i, t))
            oo, om, orw, of = o.step(ha[j])
            assert tuple(of[:3]) == tuple(hf[j][:3]), (mode, i, t, of, hf[j])
            assert np.all(np.abs(hr[j][:A] - orw[:A]) <= REL_TOL * np.abs(orw[:A])), (mode, i, t)
            if of[0] or of[1] or of[2]:
                ends += 1
                faults += int(of[2] != 0)
                episodes[i] += 1
                oo, om = fresh(o, i, episodes[i])
            assert np.array_equal(oo, ho[j]) and np.array_equal(om, hm[j]), (mode, i, t)
    for i in sample[:8]:
        for c in range(ocfg.shared_agvs if ocfg.shared_agvs >= 2 else cells):
            assert not canon.diff(oracles[i].export(c), env.export_state(i, c)), (mode, i, c)
    rec = {"mode": mode, "envs": n, "steps": steps, "env_steps": n * steps, "sampled_envs": len(sample), "sampled_episode_ends": ends,
           "sampled_faults": faults, "mismatches": 0}
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "r02_soak_shared_long.jsonl"), "a") as f:
        f.write(json.dumps(rec) + "\n")
