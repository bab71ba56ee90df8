"""The drop-in facade (dropin/FJSPParallelEnvWrapper.py) on the real device, against the oracle, driven the way
a2c.py drives the reference env (reset(options), `while env.agents:` with dict actions, unwrapped.simulation)."""
import importlib
import os
import sys

import numpy as np
import pytest

from oracle import canon
from oracle.fjsp_oracle import OracleEnv

pytestmark = pytest.mark.gpu
DROPIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi_agent_rl_for_fjsp_b200", "dropin")


@pytest.fixture
def FJSPParallelEnv(monkeypatch):
    monkeypatch.syspath_prepend(DROPIN)
    sys.modules.pop("FJSPParallelEnvWrapper", None)
    mod = importlib.import_module("FJSPParallelEnvWrapper")
    yield mod.FJSPParallelEnv
    sys.modules.pop("FJSPParallelEnvWrapper", None)


def test_dict_api_episode_matches_oracle(FJSPParallelEnv):
    env = FJSPParallelEnv()
    rs = np.random.RandomState(9)
    for ep, no in enumerate((25, 5)):
        obs, infos = env.reset(seed=40 + ep, options={"num_orders": no})
        orc = OracleEnv()
        o_o, m_o = orc.reset(env._orders)
        of, mf = canon.flatten_reference_obs(obs)
        assert np.array_equal(of, o_o) and np.array_equal(mf, m_o)
        steps = 0
        while env.agents:
            a = np.array([rs.randint(n) for n in (3, 8, 3, 3, 3, 3, 3, 3)])
            obs, rew, te, tr, inf = env.step({aid: int(a[i]) for i, aid in enumerate(env.possible_agents)})
            o_o, m_o, r_o, f_o = orc.step(a)
            of, mf = canon.flatten_reference_obs(obs)
            assert np.array_equal(of, o_o) and np.array_equal(mf, m_o)
            assert np.allclose([rew[x] for x in env.possible_agents], r_o, rtol=1e-6, atol=0)
            assert te["agv"] == bool(f_o[0]) and tr["agv"] == bool(f_o[1])
            assert inf["agv"]["sim_time"] == 10.0 * (steps + 1)
            steps += 1
        sim = env.unwrapped.simulation
        assert sim.current_step == steps == 201
        s = orc.export()
        prog = sim.get_order_progress()
        assert prog["completed_orders"] == int(s["completed_orders"]) and prog["total_orders"] == no
        assert prog["products_packaged"] == int(s["total_products_packaged"])
        assert sim.agv.position == (int(s["agv_row"]), int(s["agv_col"]))
        assert (sim.agv.carrying_tray is not None) == (int(s["agv_carry"]) >= 0)
    assert env.state().shape == (71,)
    env.close()


def test_config1_10k_steps_through_the_dict_api_bit_exact(FJSPParallelEnv):
    """BASELINE.json configs[0], literally: the default layout, 1 env, uniform-random actions, 10,000 steps through the
    dict FJSPParallelEnv on the device, against the trajectory recorded from the unmodified reference
    (tests/golden/config1_uniform.npz): per step observation, masks, rewards, flags, infos and the canonical-state digest."""
    from tests.util import REL_TOL, digest, load_golden

    g, _ = load_golden("config1_uniform")
    T = g["actions"].shape[0]
    assert T == 10000
    env = FJSPParallelEnv()
    ids = env.possible_agents
    ep_start = g["ep_start"].tolist()
    ep = 0
    for t in range(T):
        if ep < len(ep_start) and ep_start[ep] == t:
            assert t == 0 or not env.agents  # the reference's loop resets when `agents` is empty
            no = int(g["ep_norders"][ep])
            table = [tuple(int(v) for v in row) for row in g["ep_orders"][ep][:no]]
            env._gen_orders = lambda n, table=table: setattr(env, "_orders", list(table))  # the golden's explicit order table
            obs, infos = env.reset(options={"num_orders": no})
            o, m = canon.flatten_reference_obs(obs)
            assert np.array_equal(o, g["ep_obs0"][ep]) and np.array_equal(m, g["ep_masks0"][ep]), ep
            ep += 1
        assert env.agents == ids
        a = g["actions"][t]
        obs, rew, te, tr, inf = env.step({aid: int(a[i]) for i, aid in enumerate(ids)})
        o, m = canon.flatten_reference_obs(obs)
        assert np.array_equal(o, g["obs"][t]), t
        assert np.array_equal(m, g["masks"][t]), t
        r = np.array([rew[x] for x in ids])
        assert np.all(np.abs(r - g["rewards"][t]) <= REL_TOL * np.abs(g["rewards"][t])), t
        assert all(te[x] == bool(g["flags"][t][0]) and tr[x] == bool(g["flags"][t][1]) for x in ids), t
        assert digest(env._canon()) == g["hashes"][t], t
    assert ep == len(ep_start)
    env.close()


def test_dict_api_on_the_scaled_shop_matches_restatement(FJSPParallelEnv):
    """config={'num_cells': 4}: 29 agents through the dict API on the device == the C restatement of the extension."""
    from oracle.fjsp_oracle import default_config

    env = FJSPParallelEnv(config={"num_cells": 4})
    ids = env.possible_agents
    assert len(ids) == 29
    ocfg = default_config()
    ocfg.num_cells = 4
    rs = np.random.RandomState(4)
    obs, _ = env.reset(seed=7, options={"num_orders": 20})
    orc = OracleEnv(ocfg)
    o_o, m_o = orc.reset(env._orders)
    steps = 0
    while env.agents:
        acts = {a: int(rs.randint(env.action_space(a).n)) for a in ids}
        row = np.zeros(32, np.uint8)
        row[:29] = [acts[a] for a in ids]
        obs, rew, te, tr, inf = env.step(acts)
        o_o, m_o, r_o, f_o = orc.step(row)
        assert np.allclose([rew[a] for a in ids], r_o[:29], rtol=1e-6, atol=0)
        for c in range(4):
            b = 7 + 31 * c
            agv = obs["agv_c%d" % c]
            assert agv["position"].tolist() == [int(o_o[b + 4]), int(o_o[b + 5])] and int(agv["tray_type"]) == int(o_o[b + 12])
            assert agv["action_mask"].tolist() == m_o[3 + 26 * c:11 + 26 * c].tolist()
            for j, name in enumerate(("small_machine", "big_machine", "packaging_blue_1", "packaging_blue_2", "packaging_red", "packaging_green")):
                d = obs["%s_c%d" % (name, c)]
                assert int(d["is_busy"]) == int(o_o[b + 13 + 3 * j]) and float(d["processing_progress"]) == float(o_o[b + 14 + 3 * j])
                assert d["action_mask"].tolist() == m_o[11 + 26 * c + 3 * j:14 + 26 * c + 3 * j].tolist()
        assert obs["pickup_station"]["action_mask"].tolist() == m_o[:3].tolist()
        assert te["agv_c2"] == bool(f_o[0]) and tr["agv_c2"] == bool(f_o[1])
        steps += 1
    assert steps == 201 or bool(f_o[0])
    prog = env.unwrapped.simulation.get_order_progress()
    assert prog["completed_orders"] == int(orc.export()["completed_orders"]) and prog["total_orders"] == 20
    env.close()
