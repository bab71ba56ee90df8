"""Long order streams on the device (include/fjsp_b200.h "long order streams", BASELINE configs[4]): the kernels of the
second packed layout through the C ABI against the CPU restatement (long build).  The reference's own behaviour with
150-300 orders / 500-700 steps is pinned by the `long_*` goldens, which tests/test_gpu_parity.py replays like all others."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import canon
from oracle.fjsp_oracle import OracleEnv, default_config, philox_actions, philox_orders
from tests.util import REL_TOL, cfg_from_dict, load_golden

pytestmark = pytest.mark.gpu


def _cfgs(cells=1, steps=600, arr=0, arr_max=0):
    from multi_agent_rl_for_fjsp_b200 import abi

    o = default_config()
    o.num_cells, o.long_streams, o.max_episode_steps, o.arrival_prob_q16, o.arrival_max_orders = cells, 1, steps, arr, arr_max
    a = abi.FjspConfig()
    C.memmove(C.addressof(a), C.addressof(o), C.sizeof(a))
    return o, a


@pytest.mark.parametrize("cells,n_envs,num_orders,steps,arr,arr_max", [
    (1, 700, 120, 450, 0, 0), (1, 333, 4, 400, 16000, 90), (2, 200, 100, 300, 0, 0), (4, 130, 6, 260, 20000, 150)])
def test_gpu_long_batch_follows_the_restatement(cells, n_envs, num_orders, steps, arr, arr_max):
    """Philox orders, Philox actions, Philox ARRIVALS, auto-reset, ragged batch sizes: sampled envs are followed by the
    restatement step by step — observations, masks, rewards, flags, action results, canonical state, per-order records."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    ocfg, acfg = _cfgs(cells, steps, arr, arr_max)
    seed, first_env = 0xBEEF0000 + cells, 5000
    env = BatchedFJSPEnv(n_envs, config=acfg, first_env=first_env, seed=seed, num_orders=num_orders, autoreset=True, with_infos=True)
    obs0, masks0 = env.reset()
    rs = np.random.RandomState(n_envs)
    sample = sorted(set([0, n_envs - 1, 63, 64] + rs.randint(0, n_envs, size=10).tolist()))
    table_len = max(num_orders, arr_max)
    oracles, episodes = {}, {}
    obs0, masks0 = obs0.cpu().numpy(), masks0.cpu().numpy()
    for i in sample:
        o = OracleEnv(ocfg)
        oo, om = o.reset_stream(philox_orders(seed, first_env + i, 0, table_len), num_orders, seed, first_env + i, 0)
        assert np.array_equal(oo, obs0[i]) and np.array_equal(om, masks0[i])
        oracles[i], episodes[i] = o, 0
    total_T = steps + 40
    for t in range(total_T):
        acts = env.random_actions(t)
        ah = acts.cpu().numpy()
        obs, rew, term, trunc, masks = env.step(acts)
        obs, rew, masks, flags, infos = (x.cpu().numpy() for x in (obs, rew, masks, env.flags, env.infos))
        results = env.results.cpu().numpy()
        for i in sample:
            o = oracles[i]
            assert np.array_equal(ah[i], philox_actions(seed, first_env + i, t, cells)), "action stream"
            oo, om, orw, of = o.step(ah[i])
            assert tuple(of[:3]) == tuple(flags[i][:3]), (i, t, of, flags[i])
            A = 1 + 7 * cells
            assert np.all(np.abs(rew[i][:A] - orw[:A]) <= REL_TOL * np.abs(orw[:A])), (i, t)
            assert np.array_equal(o.results[:A], results[i][:A]), (i, t)
            if of[0] or of[1] or of[2]:
                assert flags[i][3] == 1
                episodes[i] += 1
                oo, om = o.reset_stream(philox_orders(seed, first_env + i, episodes[i], table_len), num_orders, seed, first_env + i, episodes[i])
            else:
                assert flags[i][3] == 0 and infos[i][0] == int(o.export()["current_step"])
            assert np.array_equal(oo, obs[i]), (i, t, np.flatnonzero(oo != obs[i]))
            assert np.array_equal(om, masks[i]), (i, t)
        if t % 97 == 0 or t == total_T - 1:
            for i in sample[:6]:
                for c in range(cells):
                    d = canon.diff(oracles[i].export(c), env.export_state(i, c))
                    assert not d, (i, t, c, d[:4])
                n_now = int(oracles[i].export()["num_orders"])
                assert np.array_equal(oracles[i].export_orders(0, n_now), env.export_orders(i, 0, n_now)), (i, t)
    assert max(episodes.values()) >= 1   # auto-reset happened


@pytest.mark.parametrize("cells", [1, 3])
def test_gpu_long_rollout_kernel_equals_stepwise(cells):
    """K-steps-per-launch kernel of the long layout == its single-step kernel: the whole packed state bit for bit, and the
    canonical records (which read the ready FIFOs) of sampled envs."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    _, acfg = _cfgs(cells, 300, 12000, 80)
    n = 1000 + 13
    a = BatchedFJSPEnv(n, config=acfg, seed=9, num_orders=10, autoreset=True)
    b = BatchedFJSPEnv(n, config=acfg, seed=9, num_orders=10, autoreset=True)
    a.reset(), b.reset()
    T = 350
    for t in range(T):
        a.step(a.random_actions(t))
    stats = b.rollout_random(T, t0=0).cpu().numpy()
    assert stats[0] == n * T and stats[1] >= n   # every env truncated at least once
    assert torch.equal(a.save_state(), b.save_state())
    for i in (0, 64, n - 1):
        assert not canon.diff(a.export_state(i), b.export_state(i))


def test_gpu_compact_golden_through_the_long_layout():
    """<= 32 orders, <= 240 steps, arrivals off: the long layout's kernels give the reference's trajectory too."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv, abi

    g, cfgd = load_golden("default_heuristic")
    ocfg = cfg_from_dict(cfgd)
    ocfg.long_streams = 1
    acfg = abi.FjspConfig()
    C.memmove(C.addressof(acfg), C.addressof(ocfg), C.sizeof(acfg))
    starts = g["ep_start"].tolist() + [g["actions"].shape[0]]
    E = len(starts) - 1
    keep = [e for e in range(E) if int(g["ep_norders"][e]) == 30]
    env = BatchedFJSPEnv(len(keep), config=acfg, autoreset=False)
    orders = np.zeros((len(keep), 30), dtype=np.uint32)
    for i, e in enumerate(keep):
        t = g["ep_orders"][e][:30]
        orders[i] = t[:, 0] | (t[:, 1] << 8) | (t[:, 2] << 16)
    env.reset(num_orders=30, orders=orders)
    lens = [starts[e + 1] - starts[e] for e in keep]
    for k in range(max(lens)):
        acts = np.zeros((len(keep), 8), dtype=np.uint8)
        live = [i for i in range(len(keep)) if k < lens[i]]
        for i in live:
            acts[i] = g["actions"][starts[keep[i]] + k]
        obs, rew, term, trunc, masks = env.step(torch.as_tensor(acts, device=env.device))
        obs, rew, masks, flags = obs.cpu().numpy(), rew.cpu().numpy(), masks.cpu().numpy(), env.flags.cpu().numpy()
        for i in live:
            t = starts[keep[i]] + k
            assert np.array_equal(obs[i], g["obs"][t]) and np.array_equal(masks[i], g["masks"][t]), (i, k)
            assert np.all(np.abs(rew[i] - g["rewards"][t]) <= REL_TOL * np.abs(g["rewards"][t])), (i, k)
            assert tuple(flags[i][:3]) == (g["flags"][t][0], g["flags"][t][1], 0), (i, k)


def test_gpu_long_host_path_equals_device_path():
    """fjsp_step_host (wire rows + host decode) on the long layout == fjsp_step, bit for bit."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    _, acfg = _cfgs(1, 400, 9000, 60)
    n = 5000
    a = BatchedFJSPEnv(n, config=acfg, seed=4, num_orders=8, autoreset=True)
    b = BatchedFJSPEnv(n, config=acfg, seed=4, num_orders=8, autoreset=True)
    a.reset(), b.reset()
    for t in range(60):
        acts = a.random_actions(t)
        obs, rew, _, _, masks = a.step(acts)
        hobs, hmasks, hrew, hflags = b.step_host(acts.cpu().numpy())
        assert np.array_equal(obs.cpu().numpy(), hobs) and np.array_equal(masks.cpu().numpy(), hmasks)
        assert np.array_equal(rew.cpu().numpy(), hrew) and np.array_equal(a.flags.cpu().numpy(), hflags)
