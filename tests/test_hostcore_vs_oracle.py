"""Long random differential of the packed-state step function (device code, host build) against the C restatement:
three policies, several order counts, non-default configs, Philox-driven episodes with in-place reset."""
import numpy as np
import pytest

from oracle import canon, policies
from oracle.fjsp_oracle import OracleEnv, default_config, philox_actions, philox_orders
from tests.host_harness.hostharness import HostEnv


def _cfg(**kw):
    c = default_config()
    for k, v in kw.items():
        if k == "pos":
            for i, (r, col) in enumerate(v):
                c.pos[i][0], c.pos[i][1] = r, col
        else:
            setattr(c, k, v)
    return c


CONFIGS = {
    "default": {},
    "far": dict(pos=[(0, 0), (0, 17), (12, 9), (19, 0), (19, 23)]),
    "cap4_fast": dict(pack_capacity=4, proc_small=10, proc_big=20, proc_pack=20),
    "storage2": dict(storage_capacity=2),
    "trays9_short": dict(num_trays=9, max_episode_steps=90),
    "step20": dict(step_size=20, proc_small=60, proc_big=120, proc_pack=40),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_random_differential(name):
    kw = CONFIGS[name]
    orc, hh = OracleEnv(_cfg(**kw)), HostEnv(_cfg(**kw))
    rs = np.random.RandomState(hash(name) % 2**31)
    pos = kw.get("pos", [(0, 0), (0, 3), (2, 3), (3, 0), (3, 5)])
    move_cell = {1: tuple(pos[0]), 2: tuple(pos[2]), 3: tuple(pos[1]), 4: tuple(pos[3]), 5: tuple(pos[4])}
    steps = 0
    for ep in range(24):
        orders = policies.random_orders(rs, [30, 25, 5, 32, 1, 0][(ep // 3) % 6])
        o1, m1 = orc.reset(orders)
        o2, m2 = hh.reset(orders) if len(orders) else hh.reset(np.zeros((0,), np.uint32))
        assert np.array_equal(o1, o2) and np.array_equal(m1, m2)
        while True:
            kind = ep % 3
            a = (policies.uniform_random(rs) if kind == 0 else policies.masked_random(rs, o1, m1) if kind == 1
                 else policies.heuristic(rs, o1, m1, noise=0.15, move_cell=move_cell))
            o1, m1, r1, f1 = orc.step(a)
            o2, m2, r2, f2 = hh.step(a)
            steps += 1
            assert np.array_equal(f1, f2), (ep, steps, f1, f2)
            if f1[2]:
                break  # R-PKG-cap-b fault: undefined from here on both sides
            assert np.array_equal(o1, o2) and np.array_equal(m1, m2), (ep, steps)
            assert np.array_equal(r1.astype(np.float32), r2), (ep, steps, r1, r2)
            assert np.array_equal(orc.results, hh.results)
            if steps % 7 == 0 or f1[0] or f1[1]:
                d = canon.diff(orc.export(), hh.export())
                assert not d, (ep, steps, d[:5])
            if f1[0] or f1[1]:
                break
    assert steps > 1500


def test_shard_map_invariance():
    """Env g behaves the same whatever (first_env, local index) pair addresses it: Philox counters use the GLOBAL index."""
    seed = 77
    for g in (0, 5, 1 << 20, (1 << 32) - 1):
        a, b = HostEnv(), OracleEnv()
        a.reset(orders=None, num_orders=30, seed=seed, genv=g, episode=3)
        b.reset(philox_orders(seed, g, 3, 30))
        for t in range(60):
            act = philox_actions(seed, g, t)
            oa, ma, ra, fa = a.step(act)
            ob, mb, rb, fb = b.step(act)
            assert np.array_equal(oa, ob) and np.array_equal(ma, mb)
        assert not canon.diff(a.export(), b.export())


def test_config_validation_messages():
    for bad in (dict(proc_small=55), dict(tray_capacity=4), dict(max_episode_steps=400), dict(pack_capacity=40),
                dict(pos=[(0, 0), (0, 0), (2, 3), (3, 0), (3, 5)])):
        with pytest.raises(ValueError):
            HostEnv(_cfg(**bad))


def test_random_config_fuzz():
    """Random (valid) configurations: layouts, timings, capacities, episode lengths.  The packed-state step function
    and the C restatement must agree on every one of them (both are generic in FjspConfig)."""
    rs = np.random.RandomState(2026)
    total = 0
    for trial in range(14):
        step = int(rs.choice([5, 10, 20]))
        cells = set()
        while len(cells) < 5:
            cells.add((int(rs.randint(0, 12)), int(rs.randint(0, 30))))
        kw = dict(pos=list(cells), step_size=step, proc_small=step * int(rs.randint(1, 8)), proc_big=step * int(rs.randint(1, 13)),
                  proc_pack=step * int(rs.randint(1, 5)), agv_speed=int(rs.randint(1, 3)), max_episode_steps=int(rs.randint(40, 241)),
                  storage_capacity=int(rs.randint(0, 6)), pack_capacity=int(rs.randint(1, 32)), num_trays=int(rs.choice([3, 40, 1000])))
        orc, hh = OracleEnv(_cfg(**kw)), HostEnv(_cfg(**kw))
        pos = kw["pos"]
        move_cell = {1: tuple(pos[0]), 2: tuple(pos[2]), 3: tuple(pos[1]), 4: tuple(pos[3]), 5: tuple(pos[4])}
        for ep in range(3):
            orders = policies.random_orders(rs, int(rs.randint(1, 33)))
            o1, m1 = orc.reset(orders)
            o2, m2 = hh.reset(orders)
            assert np.array_equal(o1, o2) and np.array_equal(m1, m2), kw
            while True:
                a = (policies.heuristic(rs, o1, m1, noise=0.2, move_cell=move_cell) if ep else policies.masked_random(rs, o1, m1))
                o1, m1, r1, f1 = orc.step(a)
                o2, m2, r2, f2 = hh.step(a)
                total += 1
                assert np.array_equal(f1, f2), (kw, f1, f2)
                if f1[2]:
                    break
                assert np.array_equal(o1, o2) and np.array_equal(m1, m2), kw
                assert np.array_equal(r1.astype(np.float32), r2), (kw, r1, r2)
                if f1[0] or f1[1]:
                    d = canon.diff(orc.export(), hh.export())
                    assert not d, (kw, d[:4])
                    break
    assert total > 2000


def test_harness_runs_the_step_function_under_bounds_checks():
    """The test-only host build range-checks every word index the step function computes (FJSP_BOUNDS_CHECK in
    fjsp_host.h: dynamically indexed words stay inside the tile's sub-tile, hot words outside it; a violation aborts).
    compute-sanitizer is closed on the GPU pool, and the kernels run the same index arithmetic, so the whole CPU suite
    running clean under the check is the addressing evidence for the step function."""
    from tests.host_harness.hostharness import lib

    assert lib().hh_bounds_checked() == 1
