"""SHARED FLOOR (include/fjsp_b200.h "shared floor"): A = 2..4 AGVs on one set of stations, each station
position holding at most one AGV — a builder-defined extension (the reference has one AGV,
/root/reference/FJSPSimulation.py:62-82).  CPU side:

* A = 1 is the reference: every reference golden replays with ``shared_agvs = 1`` (restatement and packed core);
* the occupancy rule itself, on the restatement: a position holds one AGV, standing or under way; two AGVs asking for the
  same free position are served in agent order; a position left by an earlier AGV is free for a later one in the same
  step (and not the other way round); the masks say the same; a reservation holds while the AGV is under way;
* the packed-state core (the code the CUDA kernel runs) equals the restatement step by step under three policies,
  A = 2..4, default and far layouts — observations, masks, rewards, flags, action results, every AGV's canonical record;
* both replay the regression vectors of tests/golden/shared/ (recorded from the restatement, oracle/gen_golden_shared.py);
* invariant fuzz: no two AGVs ever hold the same position.
The GPU replay is tests/test_gpu_shared_floor.py."""
import ast
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import canon, policies
from oracle.fjsp_oracle import OracleEnv, default_config, dims, philox_actions_shared, philox_orders
from tests.host_harness.hostharness import HostEnv, lib as hh_lib
from tests.util import GOLDEN_FILES, REL_TOL, cfg_from_dict, load_golden

SHARED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shared")
SHARED_FILES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(SHARED_DIR, "*.npz")))
PICKUP, BIG, SMALL, STORAGE, PACKAGING = 0, 1, 2, 3, 4       # LocationType order (constants.py)
MOVE = {PICKUP: 1, SMALL: 2, BIG: 3, STORAGE: 4, PACKAGING: 5}   # location -> AGV move action (AGVAgent.py:218-224)


def digest(s):
    return np.frombuffer(hashlib.blake2b(s.tobytes(), digest_size=8).digest(), dtype="<u8")[0]


def cfg_shared(agvs, **kw):
    c = default_config()
    c.shared_agvs = agvs
    pos = kw.pop("pos", None)
    if pos:
        for i, (r, col) in enumerate(pos):
            c.pos[i][0], c.pos[i][1] = r, col
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def load_shared(name):
    with np.load(os.path.join(SHARED_DIR, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    d = ast.literal_eval(str(g["config"]))
    cfg = default_config()
    for i, (r, c) in enumerate(d.pop("pos")):
        cfg.pos[i][0], cfg.pos[i][1] = r, c
    for k, v in d.items():
        setattr(cfg, k, v)
    return g, cfg


def replay_shared(name, make_env, exact_rewards=True):
    g, cfg = load_shared(name)
    agvs = int(g["agvs"])
    env = make_env(cfg)
    starts = g["ep_start"].tolist()
    ep = 0
    for t in range(g["actions"].shape[0]):
        if ep < len(starts) and starts[ep] == t:
            no = int(g["ep_norders"][ep])
            o, m = env.reset(g["ep_orders"][ep][:no])
            assert np.array_equal(o, g["ep_obs0"][ep]) and np.array_equal(m, g["ep_masks0"][ep]), (name, ep)
            ep += 1
        o, m, r, f = env.step(g["actions"][t])
        assert np.array_equal(o, g["obs"][t]), (name, t, np.flatnonzero(o != g["obs"][t]))
        assert np.array_equal(m, g["masks"][t]), (name, t)
        if exact_rewards:
            assert np.array_equal(np.asarray(r, np.float64), g["rewards"][t]) or np.array_equal(
                np.asarray(r), g["rewards"][t].astype(np.float32)), (name, t)
        else:
            assert np.all(np.abs(r - g["rewards"][t]) <= REL_TOL * np.abs(g["rewards"][t])), (name, t)
        assert tuple(int(x) for x in f[:3]) == tuple(int(x) for x in g["flags"][t]), (name, t, f)
        assert np.array_equal(env.results, g["results"][t]), (name, t)
        assert [digest(env.export(j)) for j in range(agvs)] == g["hashes"][t].tolist(), (name, t)
    return g["actions"].shape[0]


# ------------------------------------------------------------------------------------------------ A = 1 is the reference
@pytest.mark.parametrize("name", [n for n in GOLDEN_FILES if not n.startswith("long_")][:4])
def test_one_agv_is_the_reference(name):
    g, cfgd = load_golden(name)
    for make in (OracleEnv, HostEnv):
        cfg = cfg_from_dict(cfgd)
        cfg.shared_agvs = 1
        env = make(cfg)
        starts, ep = g["ep_start"].tolist(), 0
        for t in range(min(g["actions"].shape[0], 600)):
            if ep < len(starts) and starts[ep] == t:
                no = int(g["ep_norders"][ep])
                o, m = env.reset(g["ep_orders"][ep][:no])
                assert np.array_equal(o, g["ep_obs0"][ep]) and np.array_equal(m, g["ep_masks0"][ep])
                ep += 1
            o, m, r, f = env.step(g["actions"][t])
            assert np.array_equal(o, g["obs"][t]) and np.array_equal(m, g["masks"][t]), (name, t)
            assert np.all(np.abs(r[:8] - g["rewards"][t]) <= REL_TOL * np.abs(g["rewards"][t])), (name, t)


def test_dims_and_layout():
    for a, (agents, act, obs, mask) in {2: (9, 16, 51, 48), 3: (10, 16, 64, 48), 4: (11, 16, 77, 64)}.items():
        d = dims(1, a)
        assert (d["agents"], d["act"], d["obs"], d["mask"]) == (agents, act, obs, mask)
        nact, moff, aoff, soff = policies.shared_layout(a)
        assert len(nact) == agents and moff[-1] + 3 == 21 + 8 * a and soff + 18 == obs


# ------------------------------------------------------------------------------------------------ the occupancy rule
def _agv_pos(env, j):
    s = env.export(j)
    return int(s["agv_row"]), int(s["agv_col"])


def _act(agvs, **kw):
    a = np.zeros(dims(1, agvs)["act"], np.uint8)
    for k, v in kw.items():
        a[1 + int(k[3:])] = v   # agv0=.., agv1=..
    return a


def test_start_positions_and_masks():
    env = OracleEnv(cfg_shared(4))
    obs, masks = env.reset(philox_orders(1, 0, 0, 10))
    pos = default_config().pos
    want = [PICKUP, STORAGE, SMALL, BIG]
    for j, loc in enumerate(want):
        assert _agv_pos(env, j) == (pos[loc][0], pos[loc][1])
        m = masks[3 + 8 * j:11 + 8 * j]
        # four positions are taken (one of them its own): only PACKAGING is a legal move for every AGV
        assert m[0] == 1 and [int(m[MOVE[l]]) for l in (PICKUP, SMALL, BIG, STORAGE, PACKAGING)] == [0, 0, 0, 0, 1], (j, m)


def test_same_target_is_served_in_agent_order():
    env = OracleEnv(cfg_shared(3))
    env.reset(philox_orders(1, 0, 0, 10))           # agv_0 PICKUP, agv_1 STORAGE, agv_2 SMALL; BIG and PACKAGING free
    pos = default_config().pos
    obs, masks, rew, flags = env.step(_act(3, agv0=MOVE[PACKAGING], agv1=MOVE[PACKAGING], agv2=MOVE[PACKAGING]))
    assert _agv_pos(env, 0) == tuple(pos[PACKAGING]) and _agv_pos(env, 1) == tuple(pos[STORAGE]) and _agv_pos(env, 2) == tuple(pos[SMALL])
    assert env.results[1] & 1 and env.results[2] & 2 and env.results[3] & 2       # success | invalid | invalid
    # local rewards: move -0.1, invalid -5 (RewardModel.py:62-77) on top of the shared global term
    g = rew[0]   # the pickup station idled with orders waiting: -1 local -> global part = rew[0] + 1
    assert np.isclose(rew[1] - (g + 1.0), -0.1) and np.isclose(rew[2] - (g + 1.0), -5.0) and np.isclose(rew[3] - (g + 1.0), -5.0)


def test_a_position_left_earlier_in_the_step_is_free_for_later_agents_only():
    env = OracleEnv(cfg_shared(2))
    env.reset(philox_orders(1, 0, 0, 10))           # agv_0 PICKUP, agv_1 STORAGE
    pos = default_config().pos
    # agv_0 leaves PICKUP for BIG, agv_1 (later in agent order) takes PICKUP in the same step
    env.step(_act(2, agv0=MOVE[BIG], agv1=MOVE[PICKUP]))
    assert _agv_pos(env, 0) == tuple(pos[BIG]) and _agv_pos(env, 1) == tuple(pos[PICKUP])
    # the other way round fails: agv_0 asks for PICKUP while agv_1 (who acts later) still stands there
    env.step(_act(2, agv0=MOVE[PICKUP], agv1=MOVE[STORAGE]))
    assert env.results[1] & 2 and env.results[2] & 1
    assert _agv_pos(env, 0) == tuple(pos[BIG]) and _agv_pos(env, 1) == tuple(pos[STORAGE])
    # a move to one's own position succeeds without moving (AGVAgent.py:226-233), occupied or not
    env.step(_act(2, agv0=MOVE[BIG], agv1=0))
    assert env.results[1] & 1 and not env.results[1] & 4


def test_reservation_holds_while_under_way():
    far = [(0, 0), (0, 15), (10, 15), (20, 0), (20, 15)]   # moves take 1..3 steps
    env = OracleEnv(cfg_shared(2, pos=far, grid_rows=21, grid_cols=16))
    env.reset(philox_orders(1, 0, 0, 10))           # agv_0 PICKUP (0,0), agv_1 STORAGE (20,0)
    env.step(_act(2, agv0=MOVE[PACKAGING]))         # 35 cells: 3 steps under way
    assert int(env.export(0)["agv_is_moving"]) == 1
    obs, masks, rew, flags = env.step(_act(2, agv1=MOVE[PACKAGING]))   # reserved by agv_0
    assert env.results[2] & 2 and masks[3 + 8 + MOVE[PACKAGING]] == 0
    assert masks[3 + 8 + MOVE[PICKUP]] == 1         # agv_0's origin is free since it left
    for _ in range(3):
        env.step(_act(2))
    assert int(env.export(0)["agv_is_moving"]) == 0 and _agv_pos(env, 0) == (20, 15)


@pytest.mark.parametrize("agvs,kind,seed", [(2, "uniform", 1), (3, "masked", 2), (4, "heuristic", 3), (4, "uniform", 4)])
def test_no_two_agvs_ever_hold_one_position(agvs, kind, seed):
    far = [(0, 0), (0, 12), (10, 12), (14, 0), (14, 12)]
    for cfg in (cfg_shared(agvs), cfg_shared(agvs, pos=far, grid_rows=15, grid_cols=13)):
        env = OracleEnv(cfg)
        rs = np.random.RandomState(seed)
        obs, masks = env.reset(policies.random_orders(rs, 25))
        move_cell = {1: tuple(cfg.pos[0]), 2: tuple(cfg.pos[2]), 3: tuple(cfg.pos[1]), 4: tuple(cfg.pos[3]), 5: tuple(cfg.pos[4])}
        for t in range(201):
            a = (policies.shared_heuristic(rs, obs, masks, agvs, move_cell=move_cell) if kind == "heuristic"
                 else policies.SHARED_POLICIES[kind](rs, obs, masks, agvs))
            obs, masks, rew, flags = env.step(a)
            standing = [_agv_pos(env, j) for j in range(agvs) if not int(env.export(j)["agv_is_moving"])]
            assert len(set(standing)) == len(standing), (t, standing)
            # a masked-in move is never refused for occupancy in the next step when it is the only AGV that moves
            if flags[1]:
                break


# ------------------------------------------------------------------------------------------------ packed core == restatement
@pytest.mark.parametrize("agvs,kind,seed,far", [(2, "heuristic", 11, False), (3, "heuristic", 12, True), (4, "heuristic", 13, False),
                                                (2, "masked", 14, True), (3, "uniform", 15, False), (4, "masked", 16, True),
                                                (4, "uniform", 17, False)])
def test_packed_core_equals_restatement(agvs, kind, seed, far):
    layout = [(0, 0), (0, 15), (10, 15), (20, 0), (20, 15)]
    cfg = cfg_shared(agvs, pos=layout, grid_rows=21, grid_cols=16) if far else cfg_shared(agvs)
    move_cell = {1: tuple(cfg.pos[0]), 2: tuple(cfg.pos[2]), 3: tuple(cfg.pos[1]), 4: tuple(cfg.pos[3]), 5: tuple(cfg.pos[4])}
    orc, core = OracleEnv(cfg), HostEnv(cfg)
    rs = np.random.RandomState(seed)
    n = 7 + agvs
    for ep in range(3):
        orders = policies.random_orders(rs, [30, 32, 9][ep])
        oo, om = orc.reset(orders)
        co, cm = core.reset(orders)
        assert np.array_equal(oo, co) and np.array_equal(om, cm)
        while True:
            a = (policies.shared_heuristic(rs, oo, om, agvs, move_cell=move_cell) if kind == "heuristic"
                 else policies.SHARED_POLICIES[kind](rs, oo, om, agvs))
            oo, om, orw, of = orc.step(a)
            co, cm, cr, cf = core.step(a)
            assert np.array_equal(oo, co), (ep, np.flatnonzero(oo != co))
            assert np.array_equal(om, cm), (ep, np.flatnonzero(om != cm))
            assert np.allclose(orw[:n], cr[:n], rtol=1e-6, atol=0) and not cr[n:].any()
            assert tuple(of[:3]) == tuple(cf[:3]) and np.array_equal(orc.results, core.results)
            for j in range(agvs):
                d = canon.diff(orc.export(j), core.export(j))
                assert not d, (ep, j, d[:3])
            if of[0] or of[1] or of[2]:
                break
    # stepping on after the truncation step: inert in both (FJSP_FAULT_PAST_END)
    a = policies.shared_masked(rs, oo, om, agvs)
    oo, om, orw, of = orc.step(a)
    co, cm, cr, cf = core.step(a)
    assert tuple(of[:3]) == tuple(cf[:3]) == (0, 1, 4) and not cr.any() and np.array_equal(oo, co)


def test_philox_action_stream():
    for agvs in (2, 3, 4):
        for t in (0, 1, 77, 1 << 33):
            a = philox_actions_shared(9, 1234, t, agvs)
            b = np.zeros_like(a)
            hh_lib().hh_philox_actions_shared(9, 1234, t, agvs, b.ctypes.data)
            assert np.array_equal(a, b)
            nact = policies.shared_layout(agvs)[0]
            assert all(a[i] < n for i, n in enumerate(nact)) and not a[len(nact):].any()


def test_config_validation():
    for kw in (dict(num_cells=2), dict(long_streams=1, max_episode_steps=300)):
        with pytest.raises(ValueError):
            HostEnv(cfg_shared(2, **kw))
    c = cfg_shared(2)
    c.shared_agvs = 5
    with pytest.raises(ValueError):
        HostEnv(c)


# ------------------------------------------------------------------------------------------------ regression vectors
def test_vectors_present():
    assert {"a2_heuristic", "a3_mixed", "a4_heuristic_pack_cap3", "a3_far_layout"} <= set(SHARED_FILES)


@pytest.mark.parametrize("name", SHARED_FILES)
def test_restatement_replays_its_vectors(name):
    assert replay_shared(name, lambda cfg: OracleEnv(cfg)) > 300


@pytest.mark.parametrize("name", SHARED_FILES)
def test_packed_core_replays_shared_vectors(name):
    assert replay_shared(name, lambda cfg: HostEnv(cfg)) > 300
