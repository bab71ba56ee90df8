"""Generate the golden vectors under tests/golden/ by executing the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the build container (``/root/reference`` present):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference ships no tests or golden files, so these trajectories — produced by
the reference's own code under the import stand-ins described in ``oracle/refload.py`` — are what
pins the C restatement (and, through it, the CUDA path).  Each file stores, per step: the 8
actions, the 38-float observation in layout O, the 32 mask bytes, the 8 float64 rewards, the
terminated/truncated flags, a 64-bit digest of the canonical integer state S, and the full S
record every ``CHECK_EVERY`` steps; plus the explicit order table of every episode and the config.

Non-default scenarios patch the reference's module-level dicts in place (``constants.CONFIG``,
``LOCATION_POSITIONS``, ``PROCESSING_TIMES``) and, for the packaging capacity, the SimPy resource
of each station after reset; the same numbers go into ``FjspConfig`` on our side.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import canon, policies, refload  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
CHECK_EVERY = 250

DEFAULT_CFG = dict(
    pos=[[0, 0], [0, 3], [2, 3], [3, 0], [3, 5]],  # PICKUP, BIG, SMALL, STORAGE, PACKAGING
    grid_rows=4, grid_cols=6, proc_small=60, proc_big=120, proc_pack=30, step_size=10, agv_speed=1,
    max_episode_steps=200, storage_capacity=100, pack_capacity=20, tray_capacity=5, num_trays=1000,
)


def digest(s: np.ndarray) -> np.uint64:
    return np.frombuffer(hashlib.blake2b(s.tobytes(), digest_size=8).digest(), dtype="<u8")[0]


class PatchedReference:
    """Context manager applying a scenario config to the live reference's globals."""

    def __init__(self, cfg: dict):
        self.cfg = cfg
        self.ns = refload.load_reference()

    def __enter__(self):
        ns, cfg = self.ns, self.cfg
        c = ns.constants
        self._saved = (dict(c.CONFIG), dict(c.LOCATION_POSITIONS), dict(c.PROCESSING_TIMES))
        L = ns.LocationType
        for loc, (r, col) in zip((L.PICKUP, L.BIG_MACHINE, L.SMALL_MACHINE, L.STORAGE, L.PACKAGING), cfg["pos"]):
            c.LOCATION_POSITIONS[loc] = (int(r), int(col))
        c.PROCESSING_TIMES.update(small_machine=cfg["proc_small"], big_machine=cfg["proc_big"], packaging=cfg["proc_pack"])
        c.CONFIG.update(num_trays=cfg["num_trays"], tray_capacity=cfg["tray_capacity"], grid_rows=cfg["grid_rows"],
                        grid_cols=cfg["grid_cols"], agv_speed=cfg["agv_speed"], step_size=cfg["step_size"],
                        max_episode_steps=cfg["max_episode_steps"], storage_capacity=cfg["storage_capacity"])
        self.env = ns.FJSPParallelEnv()
        return self

    def __exit__(self, *a):
        c = self.ns.constants
        cfg0, pos0, pt0 = self._saved
        c.CONFIG.clear(), c.CONFIG.update(cfg0)
        c.LOCATION_POSITIONS.clear(), c.LOCATION_POSITIONS.update(pos0)
        c.PROCESSING_TIMES.clear(), c.PROCESSING_TIMES.update(pt0)
        return False

    def reset(self, orders):
        with refload.quiet():
            obs, _ = refload.reset_with_orders(self.env, orders)
        sim = self.env.unwrapped.simulation
        for st in sim.packaging_stations.values():
            st.resource._capacity = self.cfg["pack_capacity"]
        return obs

    def step(self, a):
        acts = {aid: int(a[i]) for i, aid in enumerate(canon.AGENT_IDS)}
        with refload.quiet():
            return self.env.step(acts)


def lazy_pack_heuristic(rs, obs, masks, move_cell):
    """Heuristic whose packaging stations start only now and then, so queues build up (capacity paths)."""
    a = policies.heuristic(rs, obs, masks, noise=0.05, move_cell=move_cell)
    for i in (4, 5, 6, 7):
        if a[i] == 1 and rs.random_sample() < 0.8:
            a[i] = 0
    return a


def storage_shuffler(rs, obs, masks, move_cell):
    """Heuristic that parks trays in storage half of the time (exercises Storage FIFO / overflow)."""
    a = policies.heuristic(rs, obs, masks, noise=0.1, move_cell=move_cell)
    if obs[9] > 0 and rs.random_sample() < 0.5:
        pos = (int(obs[11]), int(obs[12]))
        a[1] = 7 if pos == move_cell[4] else 4
    return a


def run(name, cfg_over, policy, steps, num_orders_cycle, seed):
    cfg = dict(DEFAULT_CFG)
    cfg.update(cfg_over)
    long_streams = bool(cfg.get("long_streams", 0))  # > 32 orders / > 240 steps: the port's long layout and its export conventions
    export = canon.export_reference_long if long_streams else canon.export_reference
    max_no = max(32, max(num_orders_cycle))
    move_cell = {1: tuple(cfg["pos"][0]), 2: tuple(cfg["pos"][2]), 3: tuple(cfg["pos"][1]), 4: tuple(cfg["pos"][3]),
                 5: tuple(cfg["pos"][4])}
    rs = np.random.RandomState(seed)
    T = steps
    actions = np.zeros((T, 8), np.uint8)
    obs = np.zeros((T, 38), np.float32)
    masks = np.zeros((T, 32), np.int8)
    rewards = np.zeros((T, 8), np.float64)
    flags = np.zeros((T, 2), np.uint8)
    hashes = np.zeros(T, np.uint64)
    checks = []
    check_steps = []
    ep_start, ep_orders, ep_norders, ep_obs0, ep_masks0 = [], [], [], [], []
    completed_total = 0
    with PatchedReference(cfg) as ref:
        t = 0
        ep = 0
        while t < T:
            no = num_orders_cycle[ep % len(num_orders_cycle)]
            orders = policies.random_orders(rs, no)
            robs = ref.reset(orders)
            o, m = canon.flatten_reference_obs(robs)
            tab = np.zeros((max_no, 3), np.int64)
            tab[:no] = orders
            ep_start.append(t), ep_orders.append(tab), ep_norders.append(no), ep_obs0.append(o), ep_masks0.append(m)
            sim = ref.env.unwrapped.simulation
            while ref.env.agents and t < T:
                if policy == "uniform":
                    a = policies.uniform_random(rs)
                elif policy == "masked":
                    a = policies.masked_random(rs, o, m)
                elif policy == "heuristic":
                    a = policies.heuristic(rs, o, m, noise=0.15, move_cell=move_cell)
                elif policy == "lazy_pack":
                    a = lazy_pack_heuristic(rs, o, m, move_cell)
                elif policy == "storage":
                    a = storage_shuffler(rs, o, m, move_cell)
                else:
                    raise ValueError(policy)
                # R-PKG-cap-b: a START while requests are still waiting makes the reference raise ValueError
                # a few steps later (duplicate processes); keep the goldens inside defined behaviour.
                for i, pid in enumerate(canon.PACK_IDS):
                    if a[4 + i] == 1 and len(sim.packaging_stations[pid].resource.put_queue) > 0:
                        a[4 + i] = 0
                robs, rrew, rterm, rtrunc, _ = ref.step(a)
                o, m = canon.flatten_reference_obs(robs)
                actions[t], obs[t], masks[t] = a, o, m
                rewards[t] = [rrew[aid] for aid in canon.AGENT_IDS]
                flags[t] = (int(rterm["agv"]), int(rtrunc["agv"]))
                s = export(sim)
                hashes[t] = digest(s)
                if t % CHECK_EVERY == 0 or not ref.env.agents:
                    checks.append(s), check_steps.append(t)
                t += 1
            completed_total += len(sim.completed_orders)
            ep += 1
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(
        path, config=np.array(json.dumps(cfg)), policy=np.array(policy), actions=actions, obs=obs, masks=masks,
        rewards=rewards, flags=flags, hashes=hashes, checks=np.array(checks, dtype=canon.CANON_DT),
        check_steps=np.array(check_steps, np.int64), ep_start=np.array(ep_start, np.int64),
        ep_orders=np.array(ep_orders, np.int64), ep_norders=np.array(ep_norders, np.int64),
        ep_obs0=np.array(ep_obs0, np.float32), ep_masks0=np.array(ep_masks0, np.int8))
    print("%-28s steps=%d episodes=%d orders_completed=%d terminated=%d max_reward=%.3f -> %s (%.0f KB)" % (
        name, T, len(ep_start), completed_total, int(flags[:, 0].sum()), rewards.max(), os.path.relpath(path, REPO),
        os.path.getsize(path) / 1024))


SCENARIOS = [
    # name, config overrides, policy, steps, num_orders cycle, seed
    ("config1_uniform", {}, "uniform", 10000, [30], 1001),
    ("default_masked", {}, "masked", 3000, [30, 25, 5], 1002),
    ("default_heuristic", {}, "heuristic", 6000, [30, 25, 5, 32, 1], 1003),
    ("far_layout", dict(pos=[[0, 0], [0, 17], [12, 9], [19, 0], [19, 23]], grid_rows=20, grid_cols=24), "heuristic", 2500,
     [12, 6], 1004),
    ("pack_cap3", dict(pack_capacity=3), "lazy_pack", 3000, [30, 12], 1005),
    ("pack_cap7_fast", dict(pack_capacity=7, proc_small=10, proc_big=20, proc_pack=40), "lazy_pack", 3000, [30], 1006),
    ("tiny_storage", dict(storage_capacity=1, proc_small=20, proc_big=30), "storage", 2500, [30], 1007),
    ("few_trays", dict(num_trays=6), "heuristic", 1200, [30, 8], 1008),
    ("step5_short", dict(step_size=5, proc_small=15, proc_big=25, proc_pack=10, max_episode_steps=120), "heuristic", 2500,
     [30, 20], 1009),
    ("speed2_far_step20", dict(pos=[[0, 0], [0, 29], [14, 11], [33, 2], [25, 40]], grid_rows=34, grid_cols=41, agv_speed=2,
                               step_size=20, proc_small=40, proc_big=80, proc_pack=20, max_episode_steps=150), "heuristic",
     2000, [16, 9], 1010),
    # long order streams (BASELINE configs[4]): what the reference does with num_orders = 200 / 120 and 500-step episodes
    # (FJSPSimulation.py:223,315-318), replayed by the port's long layout
    ("long_heuristic_200", dict(long_streams=1, max_episode_steps=500), "heuristic", 1600, [200, 120], 1011),
    ("long_masked_300", dict(long_streams=1, max_episode_steps=700), "masked", 1500, [300], 1012),
    ("long_fast_machines_150", dict(long_streams=1, max_episode_steps=600, proc_small=20, proc_big=30, proc_pack=10), "heuristic",
     1300, [150, 60], 1013),
]

if __name__ == "__main__":
    only = set(sys.argv[1:])
    for sc in SCENARIOS:
        if only and sc[0] not in only:
            continue
        run(*sc)
