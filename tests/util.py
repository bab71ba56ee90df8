"""Shared helpers for the parity tests: golden replay and config conversion."""
from __future__ import annotations

import glob
import hashlib
import json
import os

import numpy as np

from oracle.fjsp_oracle import FjspConfig, default_config

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_FILES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
REL_TOL = 1e-6  # BASELINE.json north_star: rewards/observations within fp32 1e-6 relative


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}  # NpzFile re-reads a member on every access; materialise once
    cfg = json.loads(str(g["config"]))
    return g, cfg


def cfg_from_dict(d: dict) -> FjspConfig:
    cfg = default_config()
    for i, (r, c) in enumerate(d["pos"]):
        cfg.pos[i][0], cfg.pos[i][1] = int(r), int(c)
    for k in ("grid_rows", "grid_cols", "proc_small", "proc_big", "proc_pack", "step_size", "agv_speed",
              "max_episode_steps", "storage_capacity", "pack_capacity", "tray_capacity", "num_trays"):
        setattr(cfg, k, int(d[k]))
    cfg.long_streams = int(d.get("long_streams", 0))   # > 32 orders / > 240 steps: the long layout (include/fjsp_b200.h)
    return cfg


def digest(s: np.ndarray) -> np.uint64:
    return np.frombuffer(hashlib.blake2b(s.tobytes(), digest_size=8).digest(), dtype="<u8")[0]


def assert_rewards_close(got, want, ctx=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
    assert np.all(err <= REL_TOL), "%s reward mismatch: got %s want %s" % (ctx, got, want)


def replay_golden(name, make_env, exact_rewards=False, max_steps=None):
    """Replay a golden trajectory through `make_env(cfg)` (an object with reset(orders)->(obs,masks),
    step(actions)->(obs,masks,rewards,flags), export()->canonical record) and compare every step."""
    g, cfgd = load_golden(name)
    env = make_env(cfg_from_dict(cfgd))
    T = g["actions"].shape[0] if max_steps is None else min(max_steps, g["actions"].shape[0])
    ep_start = g["ep_start"].tolist()
    checks = {int(t): i for i, t in enumerate(g["check_steps"].tolist())}
    ep = 0
    for t in range(T):
        if ep < len(ep_start) and ep_start[ep] == t:
            no = int(g["ep_norders"][ep])
            o, m = env.reset(g["ep_orders"][ep][:no])
            assert np.array_equal(o, g["ep_obs0"][ep]), "%s ep %d reset obs" % (name, ep)
            assert np.array_equal(m, g["ep_masks0"][ep]), "%s ep %d reset masks" % (name, ep)
            ep += 1
        o, m, r, f = env.step(g["actions"][t])
        ctx = "%s step %d" % (name, t)
        assert np.array_equal(o, g["obs"][t]), "%s obs: got %s want %s" % (ctx, o, g["obs"][t])
        assert np.array_equal(m, g["masks"][t]), "%s masks: got %s want %s" % (ctx, m, g["masks"][t])
        if exact_rewards:
            assert np.array_equal(np.asarray(r, np.float64), g["rewards"][t]), "%s rewards" % ctx
        else:
            assert_rewards_close(r, g["rewards"][t], ctx)
        assert int(f[0]) == int(g["flags"][t][0]) and int(f[1]) == int(g["flags"][t][1]), "%s flags %s" % (ctx, f)
        assert int(f[2]) == 0, "%s fault flag set" % ctx
        s = env.export()
        if t in checks:
            from oracle import canon

            d = canon.diff(g["checks"][checks[t]], s)
            assert not d, "%s canonical state: %s" % (ctx, d[:6])
        assert digest(s) == g["hashes"][t], "%s canonical-state digest" % ctx
    return T
