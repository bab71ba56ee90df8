"""Live differential test of the C restatement against the UNMODIFIED reference (build container only: needs
/root/reference; skipped on the GPU box).  A 900-episode / 177,770-step run of the same loop found 0 mismatches
(DESIGN.md §oracle); this keeps a shorter version in the suite."""
import numpy as np
import pytest

from oracle import canon, policies, refload
from oracle.fjsp_oracle import OracleEnv, default_config

pytestmark = pytest.mark.reference


def test_live_differential_three_policies():
    ns = refload.load_reference()
    env = ns.FJSPParallelEnv()
    orc = OracleEnv()
    rs = np.random.RandomState(11)
    names = ["uniform", "masked", "heuristic"]
    steps = 0
    for ep in range(18):
        pol = policies.POLICIES[names[ep % 3]]
        orders = policies.random_orders(rs, [30, 25, 5][(ep // 3) % 3])
        with refload.quiet():
            robs, _ = refload.reset_with_orders(env, orders)
        o_ref, m_ref = canon.flatten_reference_obs(robs)
        o, m = orc.reset(orders)
        assert np.array_equal(o, o_ref) and np.array_equal(m, m_ref)
        sim = env.unwrapped.simulation
        while env.agents:
            a = pol(rs, o_ref, m_ref)
            with refload.quiet():
                robs, rrew, rterm, rtrunc, _ = env.step({aid: int(a[i]) for i, aid in enumerate(canon.AGENT_IDS)})
            o_ref, m_ref = canon.flatten_reference_obs(robs)
            o, m, r, f = orc.step(a)
            assert np.array_equal(o, o_ref) and np.array_equal(m, m_ref)
            assert np.array_equal(r, np.array([rrew[x] for x in canon.AGENT_IDS]))  # exact float64
            assert (bool(f[0]), bool(f[1]), int(f[2])) == (rterm["agv"], rtrunc["agv"], 0)
            assert not canon.diff(canon.export_reference(sim), orc.export())
            steps += 1
    assert steps > 3000


def test_reference_raises_where_we_flag_a_fault():
    """R-PKG-cap-b: packaging START while requests are waiting.  We flag the fault in that step; the reference
    raises ValueError out of env.run a few steps later (duplicate processes for the same products)."""
    from oracle.gen_golden import DEFAULT_CFG, PatchedReference
    from tests.test_known_answers import drive_to_pack_overflow

    cfg = dict(DEFAULT_CFG, pack_capacity=3)
    ocfg = default_config()
    ocfg.pack_capacity = 3
    orc = OracleEnv(ocfg)
    orders = [(5, 1, 1)] * 4
    orc.reset(orders)
    with PatchedReference(cfg) as ref:
        ref.reset(orders)
        fault_step = None
        script = drive_to_pack_overflow() + [np.zeros(8, np.uint8)] * 6
        with pytest.raises(ValueError, match="not in list"):
            for t, a in enumerate(script):
                _, _, _, f = orc.step(a)
                if f[2] and fault_step is None:
                    fault_step = t
                ref.step(a)
        assert fault_step == len(drive_to_pack_overflow()) - 1


def test_survey_known_answers_on_live_reference():
    """A few §4 facts straight on the reference, so the known-answer tests are anchored to it."""
    ns = refload.load_reference()
    env = ns.FJSPParallelEnv()
    with refload.quiet():
        np.random.seed(0)
        obs, _ = env.reset(seed=0)
        sim = env.unwrapped.simulation
        got = [(len(o.products), o.products[0].product_type.value, o.products[0].packaging_color.value) for o in sim.orders[:8]]
        assert got == [(6, 1, 2), (4, 2, 3), (5, 3, 1), (9, 1, 3), (2, 3, 3), (9, 2, 2), (9, 2, 1), (4, 1, 2)]
        assert obs["agv"]["action_mask"].tolist() == [1, 0, 1, 1, 1, 1, 0, 0]
        _, r, _, _, _ = env.step({a: 0 for a in env.possible_agents})
        assert r["pickup_station"] == -1.125 and r["agv"] == -0.125
        assert env.state().shape == (71,) and env.state().dtype == np.float64
