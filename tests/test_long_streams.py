"""Long order streams (BASELINE configs[4]; include/fjsp_b200.h "long order streams"): the second packed layout on CPU.

* what the REFERENCE does with 150-300 orders and 500-700-step episodes is pinned by the `long_*` goldens (recorded from
  the unmodified reference, oracle/gen_golden.py) and replayed by the C restatement and the packed core in
  tests/test_oracle_golden.py / test_hostcore_golden.py like every other golden;
* here: the long layout with <= 32 orders and <= 240 steps reproduces every golden of the compact layout step for step;
  the restatement and the packed core agree on long random / heuristic episodes, on Philox order arrivals (an extension),
  on the scaled shop and through the cell-parallel phases; the capacity faults; the drop-in facade takes
  ``reset(options={"num_orders": 200})`` and ``max_episode_steps = 500`` like the reference."""
import importlib
import os
import sys

import numpy as np
import pytest

from oracle import canon, policies
from oracle.fjsp_oracle import OracleEnv, default_config, philox_orders
from tests.host_harness.hostharness import HostEnv
from tests.util import GOLDEN_FILES, REL_TOL, cfg_from_dict, load_golden

MOVE = {1: (0, 0), 2: (2, 3), 3: (0, 3), 4: (3, 0), 5: (3, 5)}


def cfg_long(cells=1, steps=600, arrival_q16=0, arrival_max=0, **kw):
    c = default_config()
    c.num_cells, c.long_streams, c.max_episode_steps = cells, 1, steps
    c.arrival_prob_q16, c.arrival_max_orders = arrival_q16, arrival_max
    for k, v in kw.items():
        setattr(c, k, v)
    return c


@pytest.mark.parametrize("name", [n for n in GOLDEN_FILES if not n.startswith("long_")])
def test_long_layout_reproduces_the_compact_goldens(name):
    """arrivals off, <= 32 orders, <= 240 steps: the long layout gives the reference's trajectory too (observations,
    masks, rewards, flags at every step; the canonical record differs only in its tray-entry format)."""
    g, cfgd = load_golden(name)
    cfg = cfg_from_dict(cfgd)
    cfg.long_streams = 1
    env = HostEnv(cfg)
    T = g["actions"].shape[0]
    starts = g["ep_start"].tolist()
    ep = 0
    for t in range(T):
        if ep < len(starts) and starts[ep] == t:
            no = int(g["ep_norders"][ep])
            o, m = env.reset(g["ep_orders"][ep][:no])
            assert np.array_equal(o, g["ep_obs0"][ep]) and np.array_equal(m, g["ep_masks0"][ep])
            ep += 1
        o, m, r, f = env.step(g["actions"][t])
        assert np.array_equal(o, g["obs"][t]) and np.array_equal(m, g["masks"][t]), (name, t)
        assert np.all(np.abs(r - g["rewards"][t]) <= REL_TOL * np.abs(g["rewards"][t])), (name, t)
        assert tuple(f[:3]) == (g["flags"][t][0], g["flags"][t][1], 0), (name, t)


def _policy(kind, rs, obs, masks, cells):
    if cells == 1:
        return policies.heuristic(rs, obs, masks, noise=0.1, move_cell=MOVE) if kind == "heuristic" else policies.masked_random(rs, obs, masks)
    from tests.test_scaled_shop import policy_actions

    return policy_actions(rs, obs, masks, cells, 2 if kind == "heuristic" else 1)


@pytest.mark.parametrize("cells,kind,orders,steps,arr", [
    (1, "heuristic", 250, 620, 0), (1, "masked", 300, 500, 0), (1, "heuristic", 20, 620, 9000), (1, "masked", 60, 400, 30000),
    (2, "heuristic", 200, 400, 0), (4, "heuristic", 300, 350, 12000), (3, "masked", 150, 300, 0)])
def test_packed_core_equals_restatement_on_long_episodes(cells, kind, orders, steps, arr):
    """Long episodes, with and without Philox order arrivals (`arr` = arrival_prob_q16; `orders` then counts the orders
    that can ever exist, 5 of them present at reset), K = 1..4; K >= 2 also through the cell-parallel phase functions."""
    seed, genv = 77 + cells, 1000 + orders
    total = orders
    cfg = cfg_long(cells, steps, arr, total if arr else 0)
    rs = np.random.RandomState(orders + cells)
    table = philox_orders(seed, genv, 0, total)
    initial = 5 if arr else total
    orc, core = OracleEnv(cfg), HostEnv(cfg)
    oo, om = orc.reset_stream(table, initial, seed, genv, 0)
    co, cm = core.reset(table, num_orders=initial, seed=seed, genv=genv, episode=0)
    assert np.array_equal(oo, co) and np.array_equal(om, cm)
    twin = HostEnv(cfg) if cells >= 2 else None   # the same core stepped through the cell-parallel phases
    if twin:
        twin.reset(table, num_orders=initial, seed=seed, genv=genv, episode=0)
    arrived_late = False
    for t in range(steps + 3):
        a = _policy(kind, rs, oo, om, cells)
        oo, om, orw, of = orc.step(a)
        co, cm, cr, cf = core.step(a)
        if cf[2] == 2:   # tray pool of the packed state exhausted: a capacity of the port the restatement does not model
            assert t > 200
            break
        assert np.array_equal(oo, co), (t, np.flatnonzero(oo != co))
        assert np.array_equal(om, cm), t
        assert np.allclose(orw, cr, rtol=1e-6, atol=0), (t, orw, cr)
        assert tuple(of[:3]) == tuple(cf[:3]), (t, of, cf)
        if twin:
            to, tm, tr, tf = twin.step_cells(a, reverse=bool(t & 1))
            assert np.array_equal(to, co) and np.array_equal(tm, cm) and np.array_equal(tr, cr) and tuple(tf[:3]) == tuple(cf[:3]), t
            assert np.array_equal(twin.words(), core.words()), t
        if t % 40 == 0 or t >= steps:
            for c in range(cells):
                d = canon.diff(orc.export(c), core.export(c))
                assert not d, (t, c, d[:4])
            n_now = int(orc.export()["num_orders"])
            assert np.array_equal(orc.export_orders(0, n_now), core.export_orders(0, n_now)), t
            arrived_late = arrived_late or n_now > initial
        if of[2]:
            break
    if arr:
        assert arrived_late, "no order ever arrived"
    assert t >= steps or of[2] or cf[2], "stopped early"


def test_step_after_truncation_is_inert_in_both_layouts():
    """with autoreset off an env stepped past its truncation step must not wander off silently: it stays put,
    reports truncated + FJSP_FAULT_PAST_END, rewards 0 (restatement and packed core alike, compact and long layouts)."""
    for long_streams in (0, 1):
        cfg = default_config()
        cfg.long_streams, cfg.max_episode_steps = long_streams, 30
        orc, core = OracleEnv(cfg), HostEnv(cfg)
        rs = np.random.RandomState(3)
        orders = policies.random_orders(rs, 12)
        oo, om = orc.reset(orders)
        core.reset(orders)
        for t in range(40):
            a = policies.masked_random(rs, oo, om)
            oo, om, orw, of = orc.step(a)
            co, cm, cr, cf = core.step(a)
            assert np.array_equal(oo, co) and np.array_equal(om, cm) and np.allclose(orw, cr, rtol=1e-6, atol=0)
            assert tuple(of[:3]) == tuple(cf[:3])
            if t == 30:
                assert tuple(cf[:3]) == (0, 1, 0)
                frozen = core.words().copy()
            if t > 30:
                assert tuple(cf[:3]) == (0, 1, 4) and not cr.any() and np.array_equal(core.words(), frozen)


def test_order_slot_capacity_fault():
    """More than 64 orders in process at once ends in FJSP_FAULT_ORDER_SLOTS on the pickup that would need the 65th slot —
    in the restatement and in the packed core.  Here every tray is dropped at a storage of capacity 2: from the third on
    the trays vanish (Storage.py:18-22), their orders can never complete and keep their slots, while the tray pool stays
    almost empty (a lost tray's pool slot is given back at once in the long layout)."""
    cfg = cfg_long(1, 3000, storage_capacity=2)
    orc, core = OracleEnv(cfg), HostEnv(cfg)
    orders = [(1, 1, 1)] * 200   # one product per order: one tray per order
    oo, om = orc.reset(orders)
    core.reset(orders)
    fault_at = None
    cycle = [(1, 0), (0, 6), (0, 4), (0, 7), (0, 1)]   # load, pick up at the station, move to storage, drop, move back
    for t in range(600):
        ps, agv = cycle[t % len(cycle)]
        a = np.array([ps, agv, 0, 0, 0, 0, 0, 0], np.uint8)
        oo, om, orw, of = orc.step(a)
        co, cm, cr, cf = core.step(a)
        assert np.array_equal(oo, co) and np.allclose(orw, cr, rtol=1e-6, atol=0) and tuple(of[:3]) == tuple(cf[:3]), t
        if cf[2]:
            fault_at = t
            break
    assert fault_at == 64 * 5 + 1 and int(cf[2]) == 5
    assert int(core.export()["storage_n"]) == 2


@pytest.fixture
def facade_cls(monkeypatch):
    dropin = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi_agent_rl_for_fjsp_b200", "dropin")
    monkeypatch.syspath_prepend(dropin)
    sys.modules.pop("FJSPParallelEnvWrapper", None)
    mod = importlib.import_module("FJSPParallelEnvWrapper")
    from tests.fake_device_env import FakeBatchedFJSPEnv

    monkeypatch.setattr(mod, "BatchedFJSPEnv", FakeBatchedFJSPEnv)
    yield mod.FJSPParallelEnv
    sys.modules.pop("FJSPParallelEnvWrapper", None)


def test_facade_takes_200_orders_and_500_steps_like_the_reference(facade_cls):
    """``reset(options={"num_orders": 200})`` with ``max_episode_steps = 500``
    (/root/reference/FJSPSimulation.py:223,315-318) through the dict API, against the golden recorded from the reference."""
    g, cfgd = load_golden("long_heuristic_200")
    env = facade_cls(config={"max_episode_steps": int(cfgd["max_episode_steps"])})
    ids = env.possible_agents
    table = [tuple(int(v) for v in row) for row in g["ep_orders"][0][:200]]
    env._gen_orders = lambda n: setattr(env, "_orders", list(table))
    obs, _ = env.reset(options={"num_orders": 200})
    o, m = canon.flatten_reference_obs(obs)
    assert np.array_equal(o, g["ep_obs0"][0]) and np.array_equal(m, g["ep_masks0"][0])
    t = 0
    while env.agents:
        a = g["actions"][t]
        obs, rew, te, tr, inf = env.step({aid: int(a[i]) for i, aid in enumerate(ids)})
        o, m = canon.flatten_reference_obs(obs)
        assert np.array_equal(o, g["obs"][t]) and np.array_equal(m, g["masks"][t]), t
        assert np.allclose([rew[x] for x in ids], g["rewards"][t], rtol=1e-6, atol=0), t
        assert (te["agv"], tr["agv"]) == (bool(g["flags"][t][0]), bool(g["flags"][t][1])), t
        t += 1
    assert t == 501 and env.unwrapped.simulation.current_step == 501
    prog = env.unwrapped.simulation.get_order_progress()
    assert prog["total_orders"] == 200 and prog["completed_orders"] == inf["agv"]["orders_completed"] > 5
    done = env.unwrapped.simulation.completed_orders
    assert len(done) == prog["completed_orders"] and all(o.completion_time is not None for o in done)
    assert sum(len(o.products) for o in done) <= inf["agv"]["total_products_packaged"]
    # the compact layout is kept for episodes that fit it, and the facade switches per reset
    env.reset(options={"num_orders": 25})
    assert env._env.long_streams
    small = facade_cls()
    assert not small._env.long_streams
    small.reset(options={"num_orders": 40})
    assert small._env.long_streams
