"""Known-answer facts of the reference (each verified against the live reference in
tests/test_oracle_vs_reference.py) checked on the C restatement and on the packed-state step function."""
import numpy as np
import pytest

from oracle.fjsp_oracle import OracleEnv, default_config
from tests.host_harness.hostharness import HostEnv

IDLE = [0] * 8
M = dict(ps=0, agv=3, sm=11, bm=14, b1=17, b2=20, red=23, green=26)


def act(**kw):
    a = [0] * 8
    idx = dict(ps=0, agv=1, sm=2, bm=3, b1=4, b2=5, red=6, green=7)
    for k, v in kw.items():
        a[idx[k]] = v
    return a


def drive_to_pack_overflow():
    """Action script (pack_capacity=3, orders of 5 SMALL RED products): load a tray, process it, deliver it to the
    red station, START (3 start, 2 wait), START again while they wait -> R-PKG-cap-b."""
    s = [act(ps=1)] * 5            # steps 0-4: tray of 5 ready after step 4
    s += [act(agv=6), act(agv=2), act(agv=7), act(sm=1)]   # 5 pickup, 6 move, 7 drop, 8 START (busy through step 38)
    s += [IDLE] * 30               # steps 9-38
    s += [act(sm=2), act(agv=6), act(agv=5), act(agv=7)]   # 39 signal, 40 pickup, 41 move, 42 drop at packaging
    s += [act(red=1), act(red=1)]  # 43 START (3 of 5 granted), 44 START with waiters
    return [np.array(x, dtype=np.uint8) for x in s]


@pytest.fixture(params=["oracle", "packed_core"])
def make(request):
    return (lambda cfg=None: OracleEnv(cfg)) if request.param == "oracle" else (lambda cfg=None: HostEnv(cfg))


ORDERS = [(6, 1, 2), (4, 2, 3), (5, 3, 1)] + [(3, 1, 1)] * 27  # (n, type, colour): SMALL/BLUE, MEDIUM/GREEN, BIG/RED


def test_initial_observation(make):
    e = make()
    obs, m = e.reset(ORDERS)
    assert m[0:3].tolist() == [1, 1, 0]
    assert obs[11] == 0 and obs[12] == 0
    assert m[3:11].tolist() == [1, 0, 1, 1, 1, 1, 0, 0]
    for off in (11, 14, 17, 20, 23, 26):
        assert m[off:off + 3].tolist() == [1, 0, 0]
    assert np.count_nonzero(obs) == 0


def test_idle_rewards(make):
    e = make()
    e.reset(ORDERS)
    _, _, r, f = e.step(IDLE)
    assert r.tolist() == [-1.125] + [-0.125] * 7   # pickup idles with orders waiting
    assert f[0] == 0 and f[1] == 0


def test_load_and_tray_release(make):
    e = make()
    e.reset(ORDERS)
    for k in range(4):
        obs, m, r, _ = e.step(act(ps=1))
        assert r[0] == pytest.approx(0.875) and obs[1] == k + 1 and obs[10] == 0
    obs, m, r, _ = e.step(act(ps=1))             # 5th product fills the tray: +1 +5
    assert r[0] == pytest.approx(5.875) and obs[10] == 1 and obs[1] == 0
    obs, m, r, _ = e.step(act(ps=1))             # 6th product exhausts the order: +1 +5, second tray with 1 product
    assert r[0] == pytest.approx(5.875) and obs[10] == 2 and obs[5] == 0


def test_agv_pickup_move_and_invalid(make):
    e = make()
    e.reset(ORDERS)
    for _ in range(5):
        e.step(act(ps=1))
    obs, m, r, _ = e.step(act(agv=6))
    assert r[1] == pytest.approx(1.875) and obs[9] == 1 and obs[18] == 5 and obs[19] == 1 and obs[17] == 1
    obs, m, r, _ = e.step(act(agv=6))            # pickup with full hands
    assert r[1] == pytest.approx(-5.125)
    obs, m, r, _ = e.step(act(agv=3))            # SMALL tray to the big machine: move is fine ...
    assert r[1] == pytest.approx(-0.225) and (obs[11], obs[12]) == (0, 3)
    obs, m, r, _ = e.step(act(agv=7))            # ... the drop is not
    assert r[1] == pytest.approx(-5.125) and obs[9] == 1
    obs, m, r, _ = e.step(act(agv=3))            # move to the current location: success, no 'moved'
    assert r[1] == pytest.approx(-0.125)
    obs, m, r, _ = e.step(act(agv=2))
    assert (obs[11], obs[12]) == (2, 3)          # d = 2 < 10: there within the step
    obs, m, r, _ = e.step(act(agv=7, sm=0))      # drop, the machine idles in the SAME step -> -2 (acts after the AGV)
    assert r[1] == pytest.approx(1.875) and r[2] == pytest.approx(-2.125) and obs[22] == 1
    obs, m, r, _ = e.step(act(agv=7))            # drop with empty hands
    assert r[1] == pytest.approx(-5.125)


def test_small_machine_timing(make):
    e = make()
    e.reset(ORDERS)
    for _ in range(5):
        e.step(act(ps=1))
    for a in (act(agv=6), act(agv=2), act(agv=7)):
        e.step(a)
    k = 8
    obs, m, r, _ = e.step(act(sm=1))             # step k: START with n = 5
    assert r[2] == pytest.approx(0.875) and obs[20] == 1 and obs[21] == 0.0
    for s in range(k + 1, k + 30):
        obs, m, r, _ = e.step(IDLE)
        assert obs[20] == 1, s
    obs, m, r, _ = e.step(IDLE)                  # step k + 6n
    assert obs[20] == 0 and obs[21] == 1.0 and m[11:14].tolist() == [1, 0, 1]
    obs, m, r, _ = e.step(act(sm=2))
    assert r[2] == pytest.approx(4.875) and obs[14] == 1


def test_packaging_timing_and_signal_farm(make):
    e = make()
    e.reset([(5, 1, 1)] * 3)                     # SMALL / RED
    script = [act(ps=1)] * 5 + [act(agv=6), act(agv=2), act(agv=7), act(sm=1)] + [IDLE] * 30
    script += [act(sm=2), act(agv=6), act(agv=5)]
    for a in script:
        e.step(a)
    obs, m, r, _ = e.step(act(agv=7))            # deliver to packaging: +2 +10
    assert r[1] == pytest.approx(11.875) and obs[34] == 5
    obs, m, r, _ = e.step(act(red=1))            # step j: START, L = 5
    assert r[6] == pytest.approx(1.875) and obs[32] == 1 and obs[33] == np.float32(20.0) and obs[34] == 0
    e.step(IDLE), e.step(IDLE)
    obs, m, r, f = e.step(IDLE)                  # step j + 3: 5 products packaged, order complete
    assert r.tolist()[1] == pytest.approx((100 + 50 - 1) / 8) and obs[32] == 0
    assert m[23:26].tolist() == [1, 0, 1]
    for _ in range(3):                           # SIGNAL_COMPLETE is a repeatable +20
        obs, m, r, _ = e.step(act(red=2))
        assert r[6] == pytest.approx(19.875)


def test_truncation_on_step_201(make):
    e = make()
    e.reset(ORDERS)
    for k in range(200):
        _, _, _, f = e.step(IDLE)
        assert f[1] == 0, k
    _, _, _, f = e.step(IDLE)
    assert f[1] == 1


def test_termination_needs_orders(make):
    e = make()
    e.reset(np.zeros((0, 3), dtype=np.int64))
    _, _, r, f = e.step(IDLE)
    assert f[0] == 0 and r[0] == pytest.approx(-0.125)


def test_lost_tray_on_restart(make):
    """Machine START while a finished tray was never signalled: the finished tray is overwritten and lost."""
    e = make()
    e.reset([(1, 1, 1), (1, 1, 1)] + [(2, 1, 1)] * 3)
    s = [act(ps=1), act(ps=1, agv=6), act(agv=2), act(agv=7), act(sm=1, agv=1), act(agv=6), act(agv=2), act(agv=7)]
    for a in s:
        e.step(a)
    for _ in range(4):
        e.step(IDLE)                             # first tray (1 product) finished at step 4 + 6
    st = e.export()
    assert st["machine"][0]["current_tray"] != -1 and st["machine"][0]["is_busy"] == 0 and st["machine"][0]["queue_n"] == 1
    e.step(act(sm=1))                            # START again without SIGNAL
    st2 = e.export()
    assert st2["machine"][0]["current_tray"] != st["machine"][0]["current_tray"] and st2["machine"][0]["ready_n"] == 0
    assert st2["processed_mask"][0] == 1         # the lost tray's product stays processed, but is nowhere


def test_multi_step_move_far_layout(make):
    cfg = default_config()
    for i, (r, c) in enumerate([(0, 0), (0, 17), (12, 9), (19, 0), (19, 23)]):
        cfg.pos[i][0], cfg.pos[i][1] = r, c
    e = make(cfg)
    e.reset(ORDERS)
    obs, m, r, _ = e.step(act(agv=2))            # to SMALL_MACHINE (12, 9): d = 21 -> arrives in step k + 2
    assert (obs[11], obs[12]) == (0, 0) and m[3:11].tolist() == [1, 0, 0, 0, 0, 0, 0, 0]
    obs, m, r, _ = e.step(IDLE)                  # even IDLE is invalid while moving
    assert r[1] == pytest.approx(-5.125) and (obs[11], obs[12]) == (0, 0)
    obs, m, r, _ = e.step(IDLE)
    assert (obs[11], obs[12]) == (12, 9) and r[1] == pytest.approx(-5.125)
    obs, m, r, _ = e.step(IDLE)
    assert r[1] == pytest.approx(-0.125)


def test_restart_with_waiters_faults(make):
    cfg = default_config()
    cfg.pack_capacity = 3
    e = make(cfg)
    e.reset([(5, 1, 1)] * 4)
    flags = None
    script = drive_to_pack_overflow()
    for i, a in enumerate(script):
        obs, m, r, flags = e.step(a)
        if i == len(script) - 2:                 # first START: 3 of 5 in flight, 2 keep waiting in the queue
            assert obs[32] == 1 and obs[34] == 2 and flags[2] == 0 and m[23:26].tolist() == [1, 0, 0]
    assert flags[2] == 1
