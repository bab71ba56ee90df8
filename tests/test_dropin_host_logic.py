"""Host logic of the drop-in facade (dropin/FJSPParallelEnvWrapper.py) on CPU: the device is replaced by the
test-only host build of the step function (tests/fake_device_env.py).  With the reference tree present the
reference's OWN train.py and a2c.py are run unchanged on top of the facade."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import canon, refload
from oracle.fjsp_oracle import OracleEnv

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(REPO, "multi_agent_rl_for_fjsp_b200", "dropin")


@pytest.fixture
def facade_cls(monkeypatch):
    monkeypatch.syspath_prepend(DROPIN)
    sys.modules.pop("FJSPParallelEnvWrapper", None)
    mod = importlib.import_module("FJSPParallelEnvWrapper")
    from tests.fake_device_env import FakeBatchedFJSPEnv

    monkeypatch.setattr(mod, "BatchedFJSPEnv", FakeBatchedFJSPEnv)
    yield mod.FJSPParallelEnv
    sys.modules.pop("FJSPParallelEnvWrapper", None)


def test_order_stream_known_answer(facade_cls):
    """np.random.seed(0): first 8 orders of the reference (verified against the live reference)."""
    env = facade_cls()
    env.reset(seed=0)
    want = [(6, 1, 2), (4, 2, 3), (5, 3, 1), (9, 1, 3), (2, 3, 3), (9, 2, 2), (9, 2, 1), (4, 1, 2)]
    assert env._orders[:8] == want and len(env._orders) == 30


def test_surface_and_dtypes(facade_cls):
    env = facade_cls()
    assert env.possible_agents == canon.AGENT_IDS and env.unwrapped is env
    dims = []
    for a in env.possible_agents:  # a2c._get_obs_dim (a2c.py:118-135)
        total = 0
        for key, sp in env.observation_space(a).spaces.items():
            if key == "action_mask":
                continue
            total += 1 if hasattr(sp, "n") else int(np.prod(sp.shape))
        dims.append(total)
    assert dims == [7, 13, 3, 3, 3, 3, 3, 3] and sum(dims) == 38
    assert [env.action_space(a).n for a in env.possible_agents] == [3, 8, 3, 3, 3, 3, 3, 3]
    obs, infos = env.reset(seed=1, options={"num_orders": 25})
    assert infos == {a: {} for a in env.possible_agents}
    assert obs["agv"]["position"].dtype == np.int32 and obs["agv"]["position"].shape == (2,)
    assert obs["agv"]["action_mask"].dtype == np.int8 and obs["agv"]["action_mask"].tolist() == [1, 0, 1, 1, 1, 1, 0, 0]
    assert obs["small_machine"]["is_busy"].dtype == np.int8 and obs["small_machine"]["processing_progress"].dtype == np.float32
    assert obs["pickup_station"]["order_size"].dtype == np.int32 and obs["pickup_station"]["order_size"].shape == ()
    st = env.state()
    assert st.shape == (71,) and st.dtype == np.float64
    acts = {a: 0 for a in env.possible_agents}
    o, r, te, tr, inf = env.step(acts)
    assert r["pickup_station"] == pytest.approx(-1.125) and isinstance(r["agv"], float)
    assert te["agv"] is False and tr["agv"] is False
    assert inf["agv"]["sim_time"] == 10.0 and inf["agv"]["action_result"]["success"] is True
    assert env.unwrapped.simulation.current_step == 1
    prog = env.unwrapped.simulation.get_order_progress()
    assert prog["total_orders"] == 25 and prog["completed_orders"] == 0


def test_facade_matches_oracle_episode(facade_cls):
    env = facade_cls()
    rs = np.random.RandomState(3)
    for ep in range(3):
        obs, _ = env.reset(seed=100 + ep, options={"num_orders": 25})
        orc = OracleEnv()
        o_o, m_o = orc.reset(env._orders)
        steps = 0
        while env.agents:
            a = rs.randint(0, 3, size=8)
            a[1] = rs.randint(0, 8)
            obs, rew, te, tr, inf = env.step({aid: int(a[i]) for i, aid in enumerate(canon.AGENT_IDS)})
            o_o, m_o, r_o, f_o = orc.step(a)
            of, mf = canon.flatten_reference_obs(obs)
            assert np.array_equal(of, o_o) and np.array_equal(mf, m_o)
            assert np.allclose([rew[x] for x in canon.AGENT_IDS], r_o, rtol=1e-6, atol=0)
            assert te["agv"] == bool(f_o[0]) and tr["agv"] == bool(f_o[1])
            steps += 1
        assert steps == 201 and env.unwrapped.simulation.current_step == 201


@pytest.mark.reference
def test_missing_agents_and_out_of_range_match_reference(facade_cls):
    """Missing agents are not executed and get no idle penalty; out-of-range AGV action is invalid (-5)."""
    ns = refload.load_reference()
    ref = ns.FJSPParallelEnv()
    env = facade_cls()
    with refload.quiet():
        np.random.seed(5); ref.reset()
        np.random.seed(5); env.reset()
        for acts in ({"agv": 9, "pickup_station": 1}, {"pickup_station": 7}, {}, {"small_machine": 1, "agv": 6}):
            _, r1, _, _, i1 = ref.step(dict(acts))
            _, r2, _, _, i2 = env.step(dict(acts))
            assert r2 == pytest.approx(r1, rel=1e-6)
            assert {k: v["action_result"].get("success") for k, v in i1.items()} == {k: v["action_result"].get("success") for k, v in i2.items()}


@pytest.mark.reference
def test_reference_train_py_runs_unchanged_on_facade(tmp_path):
    """The reference's own train.py + a2c.py, unmodified, on top of dropin/FJSPParallelEnvWrapper.py (device replaced
    by the host build of the step function because this container has no GPU)."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import tests.fake_device_env as f\n"
        "import multi_agent_rl_for_fjsp_b200.env as e\n"
        "e.BatchedFJSPEnv = f.FakeBatchedFJSPEnv\n"
        "from multi_agent_rl_for_fjsp_b200 import run_reference_caller as r\n"
        "r.main([%r, 'train.py', '--timesteps', '260', '--batch_size', '64', '--save_path', %r])\n"
    ) % (REPO, refload.REFERENCE_ROOT, str(tmp_path / "ckpt"))
    proc = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-3000:]
    assert "Training complete" in proc.stdout


@pytest.mark.reference
def test_reference_heuristic_policy_drives_facade_like_reference(facade_cls):
    """a2c.MultiAgentA2C._get_heuristic_actions (a2c.py:390-537) reads the simulation OBJECT GRAPH; the facade's
    SimulationView must present the same graph: both envs follow the same trajectory under that policy."""
    import importlib

    ns = refload.load_reference()
    sys.modules.setdefault("visualization", type(sys)("visualization")).GridVisualizer = object
    a2c_ref = importlib.import_module("a2c")
    ref = ns.FJSPParallelEnv()
    env = facade_cls()
    pol = a2c_ref.MultiAgentA2C._get_heuristic_actions
    with refload.quiet():
        np.random.seed(21); ref.reset(options={"num_orders": 6})
        np.random.seed(21); env.reset(options={"num_orders": 6})
        total_ref = total = 0.0
        steps = 0
        while ref.agents and steps < 400:
            a1 = pol(None, ref.unwrapped.simulation)
            a2 = pol(None, env.unwrapped.simulation)
            assert a1 == a2, (steps, a1, a2)
            _, r1, te1, tr1, i1 = ref.step(a1)
            _, r2, te2, tr2, i2 = env.step(a2)
            assert r2 == pytest.approx(r1, rel=1e-6) and te1 == te2 and tr1 == tr2
            total_ref += sum(r1.values()); total += sum(r2.values()); steps += 1
        assert not env.agents and i1["agv"]["orders_completed"] == i2["agv"]["orders_completed"] >= 1


@pytest.mark.reference
def test_reference_cli_heuristic_test_run_prints_the_same_summary(tmp_path):
    """`train.py --test_only --heuristic --load_path checkpoints/model.pt` (the reference's CLI, its own load_model and
    heuristic test loop) prints the same TEST COMPLETE summary on the facade as on the reference's own env."""
    args = "['--test_only', '--heuristic', '--load_path', %r, '--test_orders', '5', '--max_steps', '400']" % (
        refload.REFERENCE_ROOT + "/checkpoints/model.pt")
    ours = (
        "import sys; sys.path.insert(0, %r)\n"
        "import tests.fake_device_env as f\n"
        "import multi_agent_rl_for_fjsp_b200.env as e\n"
        "e.BatchedFJSPEnv = f.FakeBatchedFJSPEnv\n"
        "from multi_agent_rl_for_fjsp_b200 import run_reference_caller as r\n"
        "r.main([%r, 'train.py'] + %s)\n") % (REPO, refload.REFERENCE_ROOT, args)
    theirs = (
        "import sys, runpy; sys.path.insert(0, %r)\n"
        "from oracle import refload\n"
        "refload.load_reference()\n"
        "sys.argv = [%r] + %s\n"
        "runpy.run_path(%r, run_name='__main__')\n") % (REPO, refload.REFERENCE_ROOT + "/train.py", args,
                                                        refload.REFERENCE_ROOT + "/train.py")
    outs = []
    for code in (ours, theirs):
        p = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        tail = p.stdout[p.stdout.index("TEST COMPLETE"):]
        outs.append([ln.strip() for ln in tail.splitlines() if ":" in ln])
    assert outs[0] == outs[1] and any("Orders completed" in ln for ln in outs[0]), outs


def test_scaled_shop_facade_surface_and_parity(facade_cls):
    """config={'num_cells': 4}: the same dict surface with 1 + 7K agents; every step equals the packed-state core driven
    with the same action row (observation fields, masks, rewards, flags)."""
    from tests.host_harness.hostharness import HostEnv
    from tests.test_scaled_shop import cfg_k

    env = facade_cls(config={"num_cells": 4})
    ids = env.possible_agents
    assert len(ids) == 29 and ids[0] == "pickup_station" and ids[1] == "agv_c0" and ids[8] == "agv_c1" and ids[-1] == "packaging_green_c3"
    assert [env.action_space(a).n for a in ids] == [3] + [8, 3, 3, 3, 3, 3, 3] * 4
    dims = []
    for a in ids:
        dims.append(sum(1 if hasattr(sp, "n") else int(np.prod(sp.shape))
                        for key, sp in env.observation_space(a).spaces.items() if key != "action_mask"))
    assert dims == [7] + [13, 3, 3, 3, 3, 3, 3] * 4 and sum(dims) == 7 + 31 * 4
    np.random.seed(5)
    obs, _ = env.reset(options={"num_orders": 12})
    ref = HostEnv(cfg_k(4))
    o, m = ref.reset(np.array(env._orders))
    assert obs["agv_c1"]["position"].tolist() == [3, 0] and obs["agv_c0"]["position"].tolist() == [0, 0]
    assert obs["agv_c2"]["action_mask"].tolist() == m[3 + 52:3 + 52 + 8].tolist()
    assert env.state().shape == (7 + 3 + 4 * (13 + 8 + 6 * (3 + 3)) + 4,)
    rs = np.random.RandomState(3)
    total = 0.0
    for t in range(120):
        acts = {a: int(rs.randint(0, env.action_space(a).n)) for a in ids}
        row = np.zeros(32, np.uint8)
        row[:29] = [acts[a] for a in ids]
        obs, rew, term, trunc, infos = env.step(acts)
        o, m, r, f = ref.step(row)
        assert [rew[a] for a in ids] == [float(x) for x in r[:29]]
        assert all(term[a] == bool(f[0]) and trunc[a] == bool(f[1]) for a in ids)
        for c in range(4):
            b = 7 + 31 * c
            agv = obs["agv_c%d" % c]
            assert agv["position"].tolist() == [int(o[b + 4]), int(o[b + 5])] and int(agv["carrying_tray"]) == int(o[b + 2])
            assert agv["action_mask"].tolist() == m[3 + 26 * c:11 + 26 * c].tolist()
            pk = obs["packaging_green_c%d" % c]
            assert int(pk["queue_length"]) == int(o[b + 30]) and pk["action_mask"].tolist() == m[26 + 26 * c:29 + 26 * c].tolist()
        assert infos["agv_c3"]["sim_time"] == 10.0 * (t + 1)
        total += sum(rew.values())
        if not env.agents:
            break
    assert env.unwrapped.simulation.current_step == t + 1


@pytest.mark.reference
def test_reference_train_py_runs_unchanged_on_the_scaled_shop(tmp_path):
    """The reference's own train.py + a2c.py (generic over possible_agents), unmodified, on the 4-cell façade selected
    with FJSP_B200_NUM_CELLS=4: 29 actor networks + the centralised critic over the 131-float observation."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import tests.fake_device_env as f\n"
        "import multi_agent_rl_for_fjsp_b200.env as e\n"
        "e.BatchedFJSPEnv = f.FakeBatchedFJSPEnv\n"
        "from multi_agent_rl_for_fjsp_b200 import run_reference_caller as r\n"
        "r.main([%r, 'train.py', '--timesteps', '230', '--batch_size', '64', '--save_path', %r])\n"
    ) % (REPO, refload.REFERENCE_ROOT, str(tmp_path / "ckpt"))
    proc = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=900,
                          env=dict(os.environ, FJSP_B200_NUM_CELLS="4"))
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-3000:]
    assert "Training complete" in proc.stdout and "Number of agents: 29" in proc.stdout
