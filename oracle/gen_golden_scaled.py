"""Regression vectors for the scaled shop (K cells) under tests/golden/scaled/.

TEST INFRASTRUCTURE.  The reference has no K-cell shop, so these trajectories are recorded from the C restatement of
the builder's spec (``oracle/fjsp_oracle.c``), NOT from the reference: they pin the extension against regressions (the
packed-state core, its cell-parallel decomposition and the CUDA kernels must all replay them bit for bit); the anchors to
the reference are the ones listed in tests/test_scaled_shop.py (K = 1 is the reference shop; idle extra cells leave
cell 0 on the reference trajectory).

    python oracle/gen_golden_scaled.py            # writes tests/golden/scaled/*.npz

Per step: the action row, the observation row (7 + 31K float32), the mask row, the float64 rewards, the
terminated/truncated/fault flags, the action_result bytes and a 64-bit digest of every cell's canonical record.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import policies  # noqa: E402
from oracle.fjsp_oracle import OracleEnv, default_config, dims  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden", "scaled")


def digest(s: np.ndarray) -> np.uint64:
    return np.frombuffer(hashlib.blake2b(s.tobytes(), digest_size=8).digest(), dtype="<u8")[0]


def cell_policy_actions(rs, obs, masks, k, kind, contested=0.1):
    """The single-shop policies of oracle/policies.py applied to each cell's view (pickup station + the cell's seven
    agents); now and then every AGV asks for the dock in the same step."""
    d = dims(k)
    a = np.zeros(d["act"], np.uint8)
    for c in range(k):
        o = np.concatenate([obs[:7], obs[7 + 31 * c:38 + 31 * c]])
        m = np.zeros(32, np.int8)
        m[:3], m[3:29] = masks[:3], masks[3 + 26 * c:29 + 26 * c]
        ac = (policies.uniform_random(rs) if kind == 0 else policies.masked_random(rs, o, m) if kind == 1
              else policies.heuristic(rs, o, m, noise=0.1))
        if c == 0:
            a[0] = ac[0]
        a[1 + 7 * c:8 + 7 * c] = ac[1:]
    if rs.rand() < contested:
        a[[1 + 7 * c for c in range(k)]] = 1
    return a


def record(name, k, episodes, seed, **cfg_kw):
    cfg = default_config()
    cfg.num_cells = k
    for key, v in cfg_kw.items():
        setattr(cfg, key, v)
    env = OracleEnv(cfg)
    rs = np.random.RandomState(seed)
    rec = {n: [] for n in ("actions", "obs", "masks", "rewards", "flags", "results", "hashes")}
    ep_start, ep_orders, ep_norders, ep_obs0, ep_masks0 = [], [], [], [], []
    t = completed = 0
    for ep, (kind, norders) in enumerate(episodes):
        orders = policies.random_orders(rs, norders)
        tab = np.zeros((32, 3), np.int64)
        tab[:norders] = orders
        obs, masks = env.reset(orders)
        ep_start.append(t), ep_orders.append(tab), ep_norders.append(norders), ep_obs0.append(obs), ep_masks0.append(masks)
        while True:
            a = cell_policy_actions(rs, obs, masks, k, kind)
            obs, masks, rew, flags = env.step(a)
            rec["actions"].append(a), rec["obs"].append(obs), rec["masks"].append(masks), rec["rewards"].append(rew)
            rec["flags"].append(flags[:3].copy()), rec["results"].append(env.results.copy())
            rec["hashes"].append(np.array([digest(env.export(c)) for c in range(k)], dtype=np.uint64))
            t += 1
            if flags[0] or flags[1] or flags[2]:
                break
        completed += int(env.export()["completed_orders"])
    os.makedirs(OUT, exist_ok=True)
    cfgd = {f: int(getattr(cfg, f)) for f, _ in cfg._fields_ if f not in ("pos", "struct_size")}
    cfgd["pos"] = [[int(cfg.pos[i][0]), int(cfg.pos[i][1])] for i in range(5)]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), config=np.array(repr(cfgd)), cells=np.int64(k),
                        ep_start=np.array(ep_start), ep_orders=np.stack(ep_orders), ep_norders=np.array(ep_norders),
                        ep_obs0=np.stack(ep_obs0), ep_masks0=np.stack(ep_masks0),
                        **{n: np.stack(v) for n, v in rec.items()})
    print("%s: K=%d, %d episodes, %d steps, %d orders completed" % (name, k, len(episodes), t, completed))


if __name__ == "__main__":
    record("k4_heuristic", 4, [(2, 30), (2, 32), (2, 12), (1, 30)], seed=41)
    record("k2_mixed", 2, [(0, 30), (1, 25), (2, 30), (2, 8)], seed=42)
    record("k3_pack_cap3", 3, [(2, 30), (2, 20)], seed=43, pack_capacity=3)
