"""Regression vectors for the SHARED FLOOR (A AGVs on one set of stations with station occupancy; include/fjsp_b200.h
"shared floor") under tests/golden/shared/.

TEST INFRASTRUCTURE.  The reference has ONE AGV (/root/reference/FJSPSimulation.py:62-82), so these trajectories are
recorded from the C restatement of the builder's spec (``oracle/fjsp_oracle.c``), NOT from the reference: they pin the
extension against regressions — the packed-state core and the CUDA kernels must replay them bit for bit.  The anchor to
the reference: A = 1 IS the reference shop (every reference golden replays with shared_agvs = 1, tests/test_shared_floor.py)
and the AGV logic is the reference's own (the restated act_agv) plus one occupancy test on moves.

    python oracle/gen_golden_shared.py            # writes tests/golden/shared/*.npz

Per step: the action row, the observation row (25 + 13A float32), the mask row, the float64 rewards, the
terminated/truncated/fault flags, the action_result bytes and a 64-bit digest of every AGV's canonical record.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import policies  # noqa: E402
from oracle.fjsp_oracle import OracleEnv, default_config  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden", "shared")
KINDS = ("uniform", "masked", "heuristic")


def digest(s: np.ndarray) -> np.uint64:
    return np.frombuffer(hashlib.blake2b(s.tobytes(), digest_size=8).digest(), dtype="<u8")[0]


def record(name, agvs, episodes, seed, contested=0.1, **cfg_kw):
    cfg = default_config()
    cfg.shared_agvs = agvs
    pos = cfg_kw.pop("pos", None)
    if pos:
        for i, (r, c) in enumerate(pos):
            cfg.pos[i][0], cfg.pos[i][1] = r, c
    for key, v in cfg_kw.items():
        setattr(cfg, key, v)
    move_cell = {1: tuple(cfg.pos[0]), 2: tuple(cfg.pos[2]), 3: tuple(cfg.pos[1]), 4: tuple(cfg.pos[3]), 5: tuple(cfg.pos[4])}
    env = OracleEnv(cfg)
    rs = np.random.RandomState(seed)
    rec = {n: [] for n in ("actions", "obs", "masks", "rewards", "flags", "results", "hashes")}
    ep_start, ep_orders, ep_norders, ep_obs0, ep_masks0 = [], [], [], [], []
    t = completed = refused = 0
    for ep, (kind, norders) in enumerate(episodes):
        orders = policies.random_orders(rs, norders)
        tab = np.zeros((32, 3), np.int64)
        tab[:norders] = orders
        obs, masks = env.reset(orders)
        ep_start.append(t), ep_orders.append(tab), ep_norders.append(norders), ep_obs0.append(obs), ep_masks0.append(masks)
        while True:
            if kind == 2:
                a = policies.shared_heuristic(rs, obs, masks, agvs, noise=0.1, move_cell=move_cell)
            else:
                a = policies.SHARED_POLICIES[KINDS[kind]](rs, obs, masks, agvs)
            if rs.rand() < contested:   # every AGV asks for the same position in one step
                a[1:1 + agvs] = 1 + rs.randint(5)
            obs, masks, rew, flags = env.step(a)
            refused += int(np.sum((env.results[1:1 + agvs] & 2) != 0))
            rec["actions"].append(a), rec["obs"].append(obs), rec["masks"].append(masks), rec["rewards"].append(rew)
            rec["flags"].append(flags[:3].copy()), rec["results"].append(env.results.copy())
            rec["hashes"].append(np.array([digest(env.export(j)) for j in range(agvs)], dtype=np.uint64))
            t += 1
            if flags[0] or flags[1] or flags[2]:
                break
        completed += int(env.export()["completed_orders"])
    os.makedirs(OUT, exist_ok=True)
    cfgd = {f: int(getattr(cfg, f)) for f, _ in cfg._fields_ if f not in ("pos", "struct_size")}
    cfgd["pos"] = [[int(cfg.pos[i][0]), int(cfg.pos[i][1])] for i in range(5)]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), config=np.array(repr(cfgd)), agvs=np.int64(agvs),
                        ep_start=np.array(ep_start), ep_orders=np.stack(ep_orders), ep_norders=np.array(ep_norders),
                        ep_obs0=np.stack(ep_obs0), ep_masks0=np.stack(ep_masks0),
                        **{n: np.stack(v) for n, v in rec.items()})
    print("%s: A=%d, %d episodes, %d steps, %d orders completed, %d invalid AGV actions" % (name, agvs, len(episodes), t, completed, refused))


if __name__ == "__main__":
    record("a2_heuristic", 2, [(2, 30), (2, 32), (1, 20)], seed=51)
    record("a3_mixed", 3, [(0, 30), (1, 25), (2, 30), (2, 10)], seed=52)
    record("a4_heuristic_pack_cap3", 4, [(2, 30), (2, 24)], seed=53, pack_capacity=3)
    # a floor where moves take 1..3 steps (distances of 10..35 at speed 1, step 10): positions stay reserved while under way
    record("a3_far_layout", 3, [(2, 28), (1, 20)], seed=54, pos=[(0, 0), (0, 15), (10, 15), (20, 0), (20, 15)], grid_rows=21, grid_cols=16)
