"""Scaled shop (BASELINE configs[4], builder-defined extension: K cells sharing one pickup station and its dock).
The reference has no such shop, so there are no goldens: the anchors are (a) K = 1 is the reference shop bit for bit
(all the golden tests run through the same K-templated code), (b) extra cells that idle do not disturb cell 0, which
then follows the K = 1 trajectory, (c) the packed-state step function == the C restatement for K = 2, 3, 4 under three
policies, (d) the dock rule's known answers."""
import numpy as np
import pytest

from oracle import canon, policies
from oracle.fjsp_oracle import OracleEnv, default_config, dims, philox_actions
from tests.host_harness.hostharness import HostEnv, lib as hh_lib


def cfg_k(k, **kw):
    c = default_config()
    c.num_cells = k
    for key, v in kw.items():
        setattr(c, key, v)
    return c


def cell_view(obs, masks, c):
    """The (38-float, 32-byte) view one cell's agents + the pickup station would see in a single shop."""
    o = np.concatenate([obs[:7], obs[7 + 31 * c:7 + 31 * (c + 1)]])
    m = np.zeros(32, np.int8)
    m[:3] = masks[:3]
    m[3:29] = masks[3 + 26 * c:3 + 26 * (c + 1)]
    return o, m


def policy_actions(rs, obs, masks, k, kind):
    d = dims(k)
    a = np.zeros(d["act"], np.uint8)
    for c in range(k):
        o, m = cell_view(obs, masks, c)
        ac = (policies.uniform_random(rs) if kind == 0 else policies.masked_random(rs, o, m) if kind == 1
              else policies.heuristic(rs, o, m, noise=0.1))
        if c == 0:
            a[0] = ac[0]
        a[1 + 7 * c:8 + 7 * c] = ac[1:]
    return a


@pytest.mark.parametrize("k", [2, 3, 4])
def test_packed_core_equals_oracle(k):
    orc, hh = OracleEnv(cfg_k(k)), HostEnv(cfg_k(k))
    rs = np.random.RandomState(100 + k)
    steps = completed = 0
    for ep in range(15):
        orders = policies.random_orders(rs, [30, 32, 12][ep % 3])
        o1, m1 = orc.reset(orders)
        o2, m2 = hh.reset(orders)
        assert o1.shape == (7 + 31 * k,) and np.array_equal(o1, o2) and np.array_equal(m1, m2)
        while True:
            a = policy_actions(rs, o1, m1, k, ep % 3)
            o1, m1, r1, f1 = orc.step(a)
            o2, m2, r2, f2 = hh.step(a)
            steps += 1
            assert np.array_equal(f1, f2), (ep, steps, f1, f2)
            if f1[2]:
                break
            assert np.array_equal(o1, o2), (ep, steps, np.flatnonzero(o1 != o2))
            assert np.array_equal(m1, m2), (ep, steps, np.flatnonzero(m1 != m2))
            assert np.array_equal(r1.astype(np.float32), r2), (ep, steps, r1, r2)
            assert np.array_equal(orc.results, hh.results), (ep, steps)
            if steps % 5 == 0 or f1[0] or f1[1]:
                for c in range(k):
                    d = canon.diff(orc.export(c), hh.export(c))
                    assert not d, (ep, steps, c, d[:4])
            if f1[0] or f1[1]:
                break
        completed += int(orc.export()["completed_orders"])
    assert steps > 1200 and completed > 30


def test_idle_extra_cells_leave_cell0_on_the_reference_trajectory():
    one, four = HostEnv(cfg_k(1)), HostEnv(cfg_k(4))
    rs = np.random.RandomState(3)
    orders = policies.random_orders(rs, 30)
    o1, m1 = one.reset(orders)
    o4, m4 = four.reset(orders)
    assert np.array_equal(o4[:38], o1) and np.array_equal(m4[:29], m1[:29])
    for t in range(201):
        a1 = policies.heuristic(rs, o1, m1, noise=0.1)
        a4 = np.zeros(32, np.uint8)
        a4[:8] = a1
        o1, m1, r1, f1 = one.step(a1)
        o4, m4, r4, f4 = four.step(a4)
        assert np.array_equal(o4[:38], o1) and np.array_equal(m4[:29], m1[:29]) and np.array_equal(f1, f4)
        assert not canon.diff(one.export(), four.export(0))
        # same events, but the global part is shared by 29 agents instead of 8: r = g/A + local
        g = (r1[5].astype(np.float64)) * 8        # packaging_blue_2 only ever idles here: its reward is g/8
        assert np.allclose(r4[:8], r1 - g / 8 + g / 29, rtol=1e-6, atol=1e-6)


def test_dock_rule_known_answers():
    """One dock: AGV 0 starts on it, AGV 1 (at STORAGE) cannot be sent there until AGV 0 has left."""
    for make in (lambda: OracleEnv(cfg_k(2)), lambda: HostEnv(cfg_k(2))):
        e = make()
        obs, m = e.reset([(5, 1, 1)] * 6)
        assert (obs[11], obs[12]) == (0, 0) and (obs[7 + 31 + 4], obs[7 + 31 + 5]) == (3, 0)   # AGV 1 starts at STORAGE
        assert m[3 + 1] == 0 and m[3 + 26 + 1] == 0          # AGV 0 is there already; AGV 1: dock taken
        a = np.zeros(16, np.uint8)
        a[8] = 1                                              # AGV 1 -> PICKUP: invalid (-5)
        obs, m, r, f = e.step(a)
        assert r[8] == pytest.approx(-1 / 15 - 5.0) and (obs[7 + 31 + 4], obs[7 + 31 + 5]) == (3, 0)
        a[:] = 0
        a[1], a[8] = 4, 1                                     # AGV 0 leaves for STORAGE; AGV 1 asks in the SAME step: still taken
        obs, m, r, f = e.step(a)
        assert r[8] == pytest.approx(-1 / 15 - 5.0) and (obs[11], obs[12]) == (3, 0)
        assert m[3 + 26 + 1] == 1                             # ... but now the dock is free
        a[:] = 0
        a[1], a[8] = 1, 1                                     # both ask: AGV 0 acts first and gets it, AGV 1 is refused
        obs, m, r, f = e.step(a)
        assert r[1] == pytest.approx(-1 / 15 - 0.1) and r[8] == pytest.approx(-1 / 15 - 5.0)
        assert (obs[11], obs[12]) == (0, 0) and (obs[7 + 31 + 4], obs[7 + 31 + 5]) == (3, 0)


def test_four_cells_raise_throughput():
    """Same orders, same per-cell heuristic: four cells complete far more orders per episode than one."""
    done = {}
    for k in (1, 4):
        e = HostEnv(cfg_k(k))
        rs = np.random.RandomState(9)
        obs, m = e.reset(policies.random_orders(np.random.RandomState(1), 32))
        for t in range(201):
            obs, m, r, f = e.step(policy_actions(rs, obs, m, k, 2))
            if f[0]:
                break
        done[k] = int(e.export()["completed_orders"])
    assert done[4] >= 2 * done[1] and done[1] >= 3, done


def test_philox_action_stream_per_cell():
    for k in (1, 2, 4):
        a = philox_actions(77, 123, 9, cells=k)
        b = np.zeros(dims(k)["act"], np.uint8)
        hh_lib().hh_philox_actions(77, 123, 9, k, b.ctypes.data)
        assert np.array_equal(a, b) and a[:8].tolist() == philox_actions(77, 123, 9, cells=1).tolist()
        assert all(a[1 + 7 * c] < 8 and (a[2 + 7 * c:8 + 7 * c] < 3).all() for c in range(k))


@pytest.mark.parametrize("k", [2, 3, 4])
def test_cell_parallel_phases_equal_sequential_step(k):
    """The cell-parallel decomposition (one lane per cell, exchange area, atomic ORs on order words) gives the packed
    state and every output of the sequential step, whatever order the lanes of a phase run in."""
    seq, par = HostEnv(cfg_k(k)), HostEnv(cfg_k(k))
    rs = np.random.RandomState(500 + k)
    steps = completed = contested = 0
    for ep in range(18):
        orders = policies.random_orders(rs, [30, 32, 8][ep % 3])
        o1, m1 = seq.reset(orders)
        o2, m2 = par.reset(orders)
        assert np.array_equal(o1, o2) and np.array_equal(m1, m2)
        while True:
            a = policy_actions(rs, o1, m1, k, ep % 3)
            if rs.rand() < 0.15:                      # several AGVs ask for the dock in the same step
                a[[1 + 7 * c for c in range(k)]] = 1
                contested += 1
            o1, m1, r1, f1 = seq.step(a)
            o2, m2, r2, f2 = par.step_cells(a, reverse=bool((steps + ep) & 1))
            steps += 1
            assert np.array_equal(f1, f2), (ep, steps, f1, f2)
            assert np.array_equal(o1, o2), (ep, steps, np.flatnonzero(o1 != o2))
            assert np.array_equal(m1, m2), (ep, steps, np.flatnonzero(m1 != m2))
            assert np.array_equal(r1, r2), (ep, steps, r1, r2)
            assert np.array_equal(seq.results, par.results) and np.array_equal(seq.infos, par.infos), (ep, steps)
            assert np.array_equal(seq.words(), par.words()), (ep, steps, np.flatnonzero(seq.words() != par.words()))
            if f1[0] or f1[1] or f1[2]:
                break
        completed += int(seq.export()["completed_orders"])
    assert steps > 1500 and completed > 40 and contested > 100
