"""Replay of the scaled-shop regression vectors (tests/golden/scaled/, recorded from the C restatement of the builder's
spec by oracle/gen_golden_scaled.py — the reference has no K-cell shop) through the restatement itself, the sequential
packed-state core and its cell-parallel decomposition.  The GPU replay is in tests/test_gpu_scaled_shop.py."""
import ast
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle.fjsp_oracle import OracleEnv, default_config
from tests.host_harness.hostharness import HostEnv

SCALED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scaled")
SCALED_FILES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(SCALED_DIR, "*.npz")))


def digest(s):
    return np.frombuffer(hashlib.blake2b(s.tobytes(), digest_size=8).digest(), dtype="<u8")[0]


def load_scaled(name):
    with np.load(os.path.join(SCALED_DIR, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    d = ast.literal_eval(str(g["config"]))
    cfg = default_config()
    for i, (r, c) in enumerate(d.pop("pos")):
        cfg.pos[i][0], cfg.pos[i][1] = r, c
    for k, v in d.items():
        setattr(cfg, k, v)
    return g, cfg


def replay_scaled(name, make_env, step=None, exact_rewards=True, export=True):
    g, cfg = load_scaled(name)
    k = int(g["cells"])
    env = make_env(cfg)
    step = step or (lambda e, a, t: e.step(a))
    starts = g["ep_start"].tolist()
    ep = 0
    for t in range(g["actions"].shape[0]):
        if ep < len(starts) and starts[ep] == t:
            no = int(g["ep_norders"][ep])
            o, m = env.reset(g["ep_orders"][ep][:no])
            assert np.array_equal(o, g["ep_obs0"][ep]) and np.array_equal(m, g["ep_masks0"][ep]), (name, ep)
            ep += 1
        o, m, r, f = step(env, g["actions"][t], t)
        assert np.array_equal(o, g["obs"][t]), (name, t, np.flatnonzero(o != g["obs"][t]))
        assert np.array_equal(m, g["masks"][t]), (name, t)
        if exact_rewards:
            assert np.array_equal(np.asarray(r, np.float64), g["rewards"][t]) or np.array_equal(
                np.asarray(r), g["rewards"][t].astype(np.float32)), (name, t)
        else:
            assert np.all(np.abs(r - g["rewards"][t]) <= 1e-6 * np.abs(g["rewards"][t])), (name, t)
        assert tuple(int(x) for x in f[:3]) == tuple(int(x) for x in g["flags"][t]), (name, t, f)
        if export:
            assert [digest(env.export(c)) for c in range(k)] == g["hashes"][t].tolist(), (name, t)
    return g["actions"].shape[0]


def test_vectors_present():
    assert {"k2_mixed", "k3_pack_cap3", "k4_heuristic"} <= set(SCALED_FILES)


@pytest.mark.parametrize("name", SCALED_FILES)
def test_restatement_replays_its_vectors(name):
    assert replay_scaled(name, lambda cfg: OracleEnv(cfg)) > 300


@pytest.mark.parametrize("name", SCALED_FILES)
def test_packed_core_replays_scaled_vectors(name):
    assert replay_scaled(name, lambda cfg: HostEnv(cfg)) > 300


@pytest.mark.parametrize("name", SCALED_FILES)
def test_cell_parallel_phases_replay_scaled_vectors(name):
    assert replay_scaled(name, lambda cfg: HostEnv(cfg), step=lambda e, a, t: e.step_cells(a, reverse=bool(t & 1))) > 300
