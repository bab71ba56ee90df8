"""Drop-in ``FJSPParallelEnvWrapper`` module: ``from FJSPParallelEnvWrapper import FJSPParallelEnv`` (train.py:20).

Same PettingZoo-style surface as the reference class (/root/reference/FJSPParallelEnvWrapper.py:9-136) — dict
``reset()`` / ``step(actions)``, ``possible_agents`` / ``agents``, ``observation_space`` / ``action_space``, ``state()``,
``render()``, ``unwrapped.simulation`` — but the simulation is env 0 of a ``BatchedFJSPEnv`` of size 1: every
``step`` is one launch of the sm_100a step kernel through the C ABI.  There is no Python or CPU simulation here.

Order generation in ``reset`` performs the reference's three draws per order from the global legacy NumPy RNG
(FJSPSimulation.py:107-112), so ``np.random.seed(s)`` / ``reset(seed=s)`` produce the reference's order stream.

Scaled shop (builder-defined extension): ``config={"num_cells": K}`` or the environment variable
``FJSP_B200_NUM_CELLS=K`` (for callers that construct the env without a config, like the reference's train.py) gives the
same surface with 1 + 7K agents — ``pickup_station`` and per cell ``agv_cN, small_machine_cN, ...`` — so the reference's
own ``a2c.MultiAgentA2C`` (generic over ``possible_agents``) trains on it unchanged.  ``unwrapped.simulation`` then shows
the shared pickup station / orders and cell 0.
"""
from __future__ import annotations

import os
import re

import numpy as np
import torch

from multi_agent_rl_for_fjsp_b200 import abi, spaces as _spaces
from multi_agent_rl_for_fjsp_b200.env import AGENT_IDS, MASK_OFFSETS, N_ACTIONS, BatchedFJSPEnv, agent_layout

try:  # the reference subclasses pettingzoo.ParallelEnv; keep isinstance() working when it is installed
    from pettingzoo import ParallelEnv as _Base  # type: ignore
except Exception:  # noqa: BLE001
    class _Base:  # minimal stand-in (see compat/pettingzoo)
        @property
        def unwrapped(self):
            return self

# reference CONFIG defaults (constants.py:20-32) that the facade itself needs
_DEFAULTS = dict(num_trays=1000, tray_capacity=5, grid_rows=4, grid_cols=6, agv_speed=1, step_size=10, max_episode_steps=200)

_PS_KEYS = ["current_tray_color", "current_tray_count", "current_tray_type", "next_product_color", "next_product_type",
            "order_size", "products_remaining"]
_AGV_KEYS = ["big_machine_busy", "big_machine_ready", "carrying_tray", "pickup_ready_trays", None, None,
             "small_machine_busy", "small_machine_ready", "storage_tray_count", "tray_needs_packaging",
             "tray_needs_processing", "tray_product_count", "tray_type"]
_RESULT_KEYS = {
    "pickup_station": [("success", 0x01), ("product_loaded", 0x02), ("tray_completed", 0x04), ("idle_with_orders", 0x08)],
    "agv": [("success", 0x01), ("invalid_action", 0x02), ("moved", 0x04), ("pickup_success", 0x08), ("drop_success", 0x10),
            ("delivered_to_packaging", 0x20)],
    "machine": [("success", 0x01), ("started_processing", 0x02), ("completed_processing", 0x04), ("idle_with_queue", 0x08)],
    "packaging": [("success", 0x01), ("started_packaging", 0x02), ("completed_packaging", 0x04), ("idle_with_queue", 0x08)],
}


def draw_reference_orders(n):
    """The reference's three draws per order from the GLOBAL legacy NumPy RNG, in its order
    (FJSPSimulation.py:107-112): randint(1, 10), choice over the 3 product types, choice over the 3 colours.
    ``np.random.choice`` over a 3-element list consumes the stream identically for enum members and ints."""
    out = []
    for _ in range(n):
        k = int(np.random.randint(1, 10))
        t = int(np.random.choice([1, 2, 3]))
        c = int(np.random.choice([1, 2, 3]))
        out.append((k, t, c))
    return out


def _kind(agent_id: str) -> str:
    """'agv_c2' -> 'agv', 'packaging_red_c1' -> 'packaging', 'small_machine' -> 'machine'."""
    base = re.sub(r"_c\d+$", "", agent_id)
    if base in ("pickup_station", "agv"):
        return base
    return "machine" if "machine" in base else "packaging"


class _Obj:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class SimulationView:
    """Read-only stand-in for ``env.unwrapped.simulation`` built lazily from the exported canonical state.

    Provides what a2c.py reads: ``current_step``, ``get_order_progress()`` (a2c.py:353-354), ``agv.position`` /
    ``agv.carrying_tray`` (a2c.py:298-306), ``total_products_packaged``, ``completed_orders``, ``orders``."""

    def __init__(self, facade):
        self._f = facade

    def _s(self):
        return self._f._canon()

    @property
    def current_step(self):
        return int(self._s()["current_step"])

    @property
    def total_products_packaged(self):
        return int(self._s()["total_products_packaged"])

    @property
    def config(self):
        return self._f.config

    @property
    def orders(self):
        s, out = self._s(), []
        recs = self._f._order_records()
        for o in range(int(s["num_orders"])):
            n, t, c = self._f._orders[o]
            km, pm, done, cs = (int(v) for v in recs[o])
            prods = [_Obj(id=o * 100 + i, order_id=o, product_type=_Obj(name=("", "SMALL", "MEDIUM", "BIG")[t], value=t),
                          packaging_color=_Obj(name=("", "RED", "BLUE", "GREEN")[c], value=c),
                          is_processed=bool((pm >> i) & 1), is_packaged=bool((km >> i) & 1)) for i in range(n)]
            if cs < 0 and done:  # long order streams: the packed state has let go of the order; the facade remembers when it completed
                cs = self._f._completion_steps.get(o, -1)
            out.append(_Obj(id=o, products=prods, is_complete=bool(done),
                            completion_time=None if cs < 0 else float((cs + 1) * self._f.config["step_size"])))
        return out

    @property
    def completed_orders(self):
        done = [o for o in self.orders if o.is_complete]
        return sorted(done, key=lambda o: (o.completion_time if o.completion_time is not None else -1.0, o.id))

    # ---- object-graph views read by the reference's heuristic policy and visualiser (a2c.py:407-535)
    def _tray(self, entry):
        """Tray view from a canonical tray entry (tray_id | order<<16 | first<<22 | count<<26)."""
        entry = int(entry)
        if entry < 0:
            return None
        if self._f._env.long_streams:   # FJSP_TRAY_ENTRY_LONG
            tid, o, first, cnt = entry & 0xfff, (entry >> 12) & 0xfff, (entry >> 24) & 15, (entry >> 28) & 7
        else:
            tid, o, first, cnt = entry & 0xffff, (entry >> 16) & 63, (entry >> 22) & 15, (entry >> 26) & 7
        n, t, c = self._f._orders[o]
        km, pm = int(self._f._order_records()[o][0]), int(self._f._order_records()[o][1])
        bits = ((1 << cnt) - 1) << first
        prods = [_Obj(id=o * 100 + i, order_id=o, is_processed=bool((pm >> i) & 1), is_packaged=bool((km >> i) & 1))
                 for i in range(first, first + cnt)]
        return _Obj(id=tid, order_id=o, products=prods, capacity=5,
                    tray_type=("", "SMALL", "MEDIUM", "BIG")[t], tray_color=("", "RED", "BLUE", "GREEN")[c],
                    needs_processing=(pm & bits) != bits, needs_packaging=(km & bits) != bits)

    def _trays(self, arr, n):
        return [self._tray(e) for e in arr[:int(n)]]

    @property
    def agv(self):
        s = self._s()
        return _Obj(position=(int(s["agv_row"]), int(s["agv_col"])), carrying_tray=self._tray(s["agv_carry"]),
                    is_moving=bool(s["agv_is_moving"]))

    @property
    def pickup_station(self):
        s = self._s()
        orders = self.orders
        n = int(s["num_orders"])
        cur = int(s["ps_current_order"])
        return _Obj(order_queue=orders[n - int(s["ps_order_queue_len"]):], current_order=orders[cur] if cur >= 0 else None,
                    current_order_product_idx=int(s["ps_product_idx"]), current_tray=self._tray(s["ps_current_tray"]),
                    ready_trays=self._trays(s["ps_ready"], s["ps_ready_n"]),
                    trays_at_station=[None] * int(s["ps_trays_at_station"]))

    def _machine(self, i):
        m = self._s()["machine"][i]
        return _Obj(is_busy=bool(m["is_busy"]), current_tray=self._tray(m["current_tray"]),
                    processing_progress=float(m["progress_done"]), tray_queue=self._trays(m["queue"], m["queue_n"]),
                    ready_trays=self._trays(m["ready"], m["ready_n"]))

    @property
    def small_machine(self):
        return self._machine(0)

    @property
    def big_machine(self):
        return self._machine(1)

    @property
    def storage(self):
        s = self._s()
        return _Obj(trays=self._trays(s["storage"], s["storage_n"]), capacity=self._f.config.get("storage_capacity", 100))

    @property
    def packaging_stations(self):
        s, out = self._s(), {}
        for i, name in enumerate(AGENT_IDS[4:]):
            p = s["pack"][i]
            cp = int(p["current_product"])
            out[name] = _Obj(is_busy=bool(p["is_busy"]), products_completed=int(p["products_completed"]),
                             product_queue=[_Obj(id=int(x), order_id=int(x) // 100) for x in p["queue"][:int(p["queue_n"])]],
                             current_product=_Obj(id=cp, order_id=cp // 100) if cp >= 0 else None,
                             processing_progress=(100.0 / int(p["progress_L"])) if int(p["progress_L"]) else 0.0)
        return out

    def get_order_progress(self):  # FJSPSimulation.py:260-284
        orders = self.orders
        return {
            "total_orders": len(orders),
            "completed_orders": sum(1 for o in orders if o.is_complete),
            "total_products": sum(len(o.products) for o in orders),
            "products_processed": sum(sum(1 for p in o.products if p.is_processed) for o in orders),
            "products_packaged": self.total_products_packaged,
            "orders_detail": [{"order_id": o.id, "total_products": len(o.products),
                               "processed": sum(1 for p in o.products if p.is_processed),
                               "packaged": sum(1 for p in o.products if p.is_packaged),
                               "is_complete": o.is_complete} for o in orders],
        }


class FJSPParallelEnv(_Base):
    metadata = {"name": "fjsp_v1", "render_modes": ["human", "rgb_array"], "is_parallelizable": True}

    def __init__(self, config=None, render_mode=None, device="cuda:0"):
        self.config = dict(_DEFAULTS)
        if config:
            self.config.update(config)
        self.render_mode = render_mode
        if "num_cells" not in self.config and os.environ.get("FJSP_B200_NUM_CELLS"):
            self.config["num_cells"] = int(os.environ["FJSP_B200_NUM_CELLS"])
        self.num_cells = int(self.config.get("num_cells", 1))
        self._device = device
        # the reference accepts any num_orders (FJSPSimulation.py:315-318) and max_episode_steps (:223): episodes beyond the
        # compact packed state (32 orders, 240 steps) run on the long order-stream layout (include/fjsp_b200.h)
        if int(self.config.get("max_episode_steps", 200)) > 240:
            self.config["long_streams"] = 1
        self._env = BatchedFJSPEnv(1, config=abi.config_from_dict(self.config), device=device, autoreset=False,
                                   with_infos=True)
        self._completion_steps, self._orders_done_seen, self._recs = {}, 0, None
        self._ids, self._nact, self._obs_slices, self._mask_off = agent_layout(self.num_cells)
        self._dims = abi.dims(self.num_cells)
        self.possible_agents = list(self._ids)
        self.agents = self.possible_agents.copy()
        self._orders = []
        self._cache = None
        self.simulation = SimulationView(self)
        self._obs_spaces = {}
        for a in self._ids:
            kind = _kind(a)
            self._obs_spaces[a] = (_spaces.pickup_station() if kind == "pickup_station"
                                   else _spaces.agv(self.config["grid_rows"], self.config["grid_cols"], self.config["tray_capacity"])
                                   if kind == "agv" else _spaces.machine() if kind == "machine" else _spaces.packaging())
        self._env.reset(orders=self._order_table())  # no orders (and no RNG draws) until reset(), as in the reference

    # ---- spaces
    def observation_space(self, agent):
        return self._obs_spaces[agent]

    def action_space(self, agent):
        return _spaces.action_space(_kind(agent))

    # ---- reset / step
    def _gen_orders(self, n):
        self._orders = draw_reference_orders(n)

    def _order_table(self):
        if len(self._orders) > abi.LONG_MAX_ORDERS:
            raise ValueError("num_orders > %d is not supported by the packed state" % abi.LONG_MAX_ORDERS)
        if len(self._orders) > abi.MAX_ORDERS and not self._env.long_streams:
            # more orders than the compact layout holds: continue on the long order-stream layout
            self._env.close()
            self.config["long_streams"] = 1
            self._env = BatchedFJSPEnv(1, config=abi.config_from_dict(self.config), device=self._device, autoreset=False,
                                       with_infos=True)
        tab = np.zeros((1, max(len(self._orders), 1) if self._env.long_streams else abi.MAX_ORDERS), dtype=np.uint32)
        for i, (k, t, c) in enumerate(self._orders):
            tab[0, i] = abi.order_rec(k, t, c)
        self._env.num_orders = len(self._orders)
        return tab

    def reset(self, seed=None, options=None):
        self.agents = self.possible_agents.copy()
        num_orders = options.get("num_orders") if options else None
        if seed is not None:
            np.random.seed(seed)
        self._gen_orders(num_orders if num_orders is not None else 30)
        obs, masks = self._env.reset(orders=self._order_table())
        self._cache, self._recs = None, None
        self._completion_steps, self._orders_done_seen = {}, 0
        observations = self._obs_dicts(obs[0].cpu().numpy(), masks[0].cpu().numpy())
        return observations, {a: {} for a in self.possible_agents}

    def step(self, actions):
        ids, d = self._ids, self._dims
        n_act, n_obs, n_mask = d["act"], d["obs"], d["mask"]
        a = np.zeros((1, n_act), dtype=np.uint8)
        present = []
        for i, aid in enumerate(ids):
            if aid in actions:
                v = int(actions[aid])
                # out-of-range actions: AGV -> invalid_action (AGVAgent.py:249-250), others -> silently nothing
                a[0, i] = v if 0 <= v < 255 else 255
                present.append(aid)
            # a missing agent is not executed and is rewarded as action 0 (FJSPSimulation.py:172-174,201):
            # for every agent action 0 has exactly that effect, except that its idle penalty must not apply.
        e = self._env
        obs, rew, term, trunc, masks = e.step(torch.from_numpy(a))
        self._cache, self._recs = None, None
        out = torch.cat([obs[0], rew[0], e.flags[0].float(), e.infos[0].float(), e.results[0].float(),
                         masks[0].float()]).cpu().numpy()
        p = n_obs + n_act
        o, r = out[:n_obs], out[n_obs:p]
        flags, infos = out[p:p + 4].astype(np.int64), out[p + 4:p + 8].astype(np.int64)
        results, m = out[p + 8:p + 8 + n_act].astype(np.int64), out[p + 8 + n_act:p + 8 + n_act + n_mask]
        observations = self._obs_dicts(o, m.astype(np.int8))
        rewards = {aid: float(r[i]) for i, aid in enumerate(ids)}
        for i, aid in enumerate(ids):
            if aid not in actions:  # undo the idle penalty of the implicit action 0
                res, kind = int(results[i]), _kind(aid)
                if kind == "pickup_station" and res & 0x08:
                    rewards[aid] += 1.0
                elif kind == "machine" and res & 0x08:
                    rewards[aid] += 2.0
                elif kind == "packaging" and res & 0x08:
                    rewards[aid] += 1.0
                elif kind == "agv" and res & 0x02:  # implicit IDLE while moving is not an invalid action
                    rewards[aid] += 5.0
        if flags[2] == 1:  # FJSP_FAULT_PKG_RESTART_WITH_WAITERS: the one exception the reference itself raises
            raise ValueError("list.remove(x): x not in list  [packaging START while requests were still waiting; "
                             "the reference raises here too (SURVEY R-PKG-cap-b)]")
        if flags[2]:      # limits of the packed state: not reference behaviour, reported as such
            what = {2: "more than %d trays in transit (tray pool of the packed state)" % 64,
                    3: "an order record outside n 1..9 / type 1..3 / colour 1..3",
                    4: "stepped past the end of the episode (call reset() after truncation)",
                    5: "more open orders than the order ring of the packed state holds"}.get(int(flags[2]), "fault %d" % int(flags[2]))
            raise RuntimeError("fjsp_b200 capacity limit, not a reference error: " + what)
        if e.long_streams and int(infos[1]) != self._orders_done_seen:
            # an order's record leaves the packed state one step after it completes: note its completion step now
            self._orders_done_seen = int(infos[1])
            for o, r in enumerate(self._order_records()):
                if r[2] and r[3] >= 0:
                    self._completion_steps.setdefault(o, int(r[3]))
        terminated, truncated = bool(flags[0]), bool(flags[1])
        terminations = {aid: terminated for aid in ids}
        truncations = {aid: truncated for aid in ids}
        sim_time = float(int(infos[0]) * self.config["step_size"])
        infos_d = {}
        for i, aid in enumerate(ids):
            kind = _kind(aid)
            if aid in actions:
                ar = {"action": int(actions[aid])}
                ar.update({k: bool(int(results[i]) & bit) for k, bit in _RESULT_KEYS[kind]})
            else:
                ar = {}
            infos_d[aid] = {"action_result": ar, "sim_time": sim_time, "orders_completed": int(infos[1]),
                            "total_products_packaged": int(infos[2])}
        if terminated or truncated:
            self.agents = []
        return observations, rewards, terminations, truncations, infos_d

    # ---- observation dicts with the reference's dtypes (AGVAgent.py:60-75, MachineAgent.py:64-69, ...)
    def _obs_dicts(self, o, m):
        i32 = lambda v: np.array(int(v), dtype=np.int32)  # noqa: E731
        i8 = lambda v: np.array(int(v), dtype=np.int8)  # noqa: E731
        obs = {}
        for j, aid in enumerate(self._ids):
            b, kind = self._obs_slices[j][0], _kind(aid)
            mask = m[self._mask_off[j]:self._mask_off[j] + self._nact[j]].astype(np.int8)
            if kind == "pickup_station":
                d = {k: i32(o[b + i]) for i, k in enumerate(_PS_KEYS)}
            elif kind == "agv":
                d = {k: i32(o[b + i]) for i, k in enumerate(_AGV_KEYS) if k is not None}
                d["position"] = np.array([int(o[b + 4]), int(o[b + 5])], dtype=np.int32)
            else:
                d = {"is_busy": i8(o[b]), "processing_progress": np.array(o[b + 1], dtype=np.float32), "queue_length": i8(o[b + 2])}
            d["action_mask"] = mask
            obs[aid] = d
        return obs

    def _canon(self):
        if self._cache is None:
            self._cache = self._env.export_state(0)
        return self._cache

    def _order_records(self):
        """[num_orders, 4]: packaged_mask, processed_mask, is_complete, completion_step of every order of the episode."""
        if self._recs is None:
            self._recs = self._env.export_orders(0, 0, max(len(self._orders), 1))
        return self._recs

    # ---- misc surface
    def render(self):
        if self.render_mode == "human":
            s = self._canon()
            print("Step %d | orders %d/%d | packaged %d | AGV (%d,%d) carrying=%s" % (
                s["current_step"], s["completed_orders"], s["num_orders"], s["total_products_packaged"], s["agv_row"],
                s["agv_col"], "yes" if s["agv_carry"] >= 0 else "no"))
        elif self.render_mode == "rgb_array":
            grid = np.ones((self.config["grid_rows"], self.config["grid_cols"], 3), dtype=np.uint8) * 255
            s = self._canon()
            grid[int(s["agv_row"]), int(s["agv_col"])] = [0, 0, 0]
            return grid

    def close(self):
        self._env.close()

    def state(self):
        """71-float64 global state (FJSPParallelEnvWrapper.py:119-136): per agent sorted keys incl. action_mask,
        then [len(orders), len(completed_orders), total_products_packaged, env.now]."""
        e = self._env
        o, m = e.obs[0].cpu().numpy(), e.masks[0].cpu().numpy()
        d = self._obs_dicts(o, m)
        parts = []
        for aid in self.possible_agents:
            for key in sorted(d[aid].keys()):
                parts.append(np.asarray(d[aid][key]).flatten())
        s = self._canon()
        parts.append(np.array([s["num_orders"], s["completed_orders"], s["total_products_packaged"],
                               float(int(s["current_step"]) * self.config["step_size"])], dtype=np.float32))
        return np.concatenate(parts)
