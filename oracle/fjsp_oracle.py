"""ctypes binding of the CPU restatement ``oracle/fjsp_oracle.c`` (TEST INFRASTRUCTURE ONLY).

May be imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.  The product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

try:  # package-style import (tests put the repo root on sys.path)
    from oracle.canon import CANON_DT
except ImportError:  # pragma: no cover - direct script use
    from canon import CANON_DT

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfjsp_oracle.so")
_SO_LONG = os.path.join(_HERE, "_build", "libfjsp_oracle_long.so")  # the same source built with room for long order streams

N_ACTIONS = (3, 8, 3, 3, 3, 3, 3, 3)


class FjspConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("pos", (C.c_int32 * 2) * 5),
        ("grid_rows", C.c_int32), ("grid_cols", C.c_int32),
        ("proc_small", C.c_int32), ("proc_big", C.c_int32), ("proc_pack", C.c_int32),
        ("step_size", C.c_int32), ("agv_speed", C.c_int32), ("max_episode_steps", C.c_int32),
        ("storage_capacity", C.c_int32), ("pack_capacity", C.c_int32),
        ("tray_capacity", C.c_int32), ("num_trays", C.c_int32), ("num_cells", C.c_int32),
        ("long_streams", C.c_int32), ("arrival_prob_q16", C.c_int32), ("arrival_max_orders", C.c_int32),
        ("shared_agvs", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile the restatement with the system gcc (``make -C oracle``)."""
    src = os.path.join(_HERE, "fjsp_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "fjsp_b200.h")
    stale = any((not os.path.exists(so)) or any(os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr))
                for so in (_SO, _SO_LONG))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"] if force else ["make", "-C", _HERE], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None
_lib_long = None


def lib(long_streams: bool = False):
    """The restatement; long_streams=True: the build with room for 4095 orders / 1000 trays per env (~2 MB per env)."""
    global _lib, _lib_long
    if long_streams:
        if _lib_long is None:
            _lib_long = _bind(_SO_LONG)
        return _lib_long
    if _lib is None:
        _lib = _bind(_SO)
    return _lib


def _bind(path):
    if True:
        build()
        L = C.CDLL(path)
        L.fjsp_oracle_create.restype = C.c_void_p
        L.fjsp_oracle_create.argtypes = [C.POINTER(FjspConfig)]
        L.fjsp_oracle_destroy.argtypes = [C.c_void_p]
        L.fjsp_oracle_default_config.argtypes = [C.POINTER(FjspConfig)]
        L.fjsp_oracle_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.fjsp_oracle_step.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.fjsp_oracle_observe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fjsp_oracle_export.argtypes = [C.c_void_p, C.c_void_p]
        L.fjsp_oracle_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fjsp_oracle_philox_orders.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
        L.fjsp_oracle_philox_actions.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.fjsp_oracle_philox_actions_k.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        L.fjsp_oracle_export_cell.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.fjsp_oracle_env_size.restype = C.c_int64
        L.fjsp_oracle_rollout_random.argtypes = [
            C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
            C.c_void_p]
        L.fjsp_oracle_reset_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint32]
        L.fjsp_oracle_export_orders.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.fjsp_oracle_philox_actions_shared.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    return L


def default_config() -> FjspConfig:
    cfg = FjspConfig()
    lib().fjsp_oracle_default_config(C.byref(cfg))
    return cfg


def order_rec(n, t, c) -> int:
    return int(n) | (int(t) << 8) | (int(c) << 16)


def philox(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().fjsp_oracle_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def philox_orders(seed: int, genv: int, episode: int, num_orders: int) -> np.ndarray:
    out = np.zeros(max(32, num_orders), dtype=np.uint32)
    lib().fjsp_oracle_philox_orders(seed, genv, episode, num_orders, out.ctypes.data)
    return out[:num_orders]


def philox_actions(seed: int, genv: int, t: int, cells: int = 1) -> np.ndarray:
    out = np.zeros(dims(cells)["act"], dtype=np.uint8)
    lib().fjsp_oracle_philox_actions_k(seed, genv, t, cells, out.ctypes.data)
    return out


def philox_actions_shared(seed: int, genv: int, t: int, agvs: int) -> np.ndarray:
    out = np.zeros(dims(1, agvs)["act"], dtype=np.uint8)
    lib().fjsp_oracle_philox_actions_shared(seed, genv, t, agvs, out.ctypes.data)
    return out


def dims(cells: int = 1, shared_agvs: int = 0) -> dict:
    """Row widths of a K-cell shop (include/fjsp_b200.h FJSP_*_K); K = 1 is the reference shop.  shared_agvs >= 2: the
    shared floor (FJSP_SHARED_*)."""
    if shared_agvs >= 2:
        a = shared_agvs
        return {"agents": 7 + a, "act": (7 + a + 7) // 8 * 8, "obs": 25 + 13 * a, "mask": (21 + 8 * a + 15) // 16 * 16}
    agents = 1 + 7 * cells
    return {"agents": agents, "act": (agents + 7) // 8 * 8, "obs": 7 + 31 * cells, "mask": (3 + 26 * cells + 31) // 32 * 32}


class OracleEnv:
    """One shop floor, stepped on the CPU by the C restatement."""

    def __init__(self, cfg: FjspConfig | None = None):
        self.cfg = cfg if cfg is not None else default_config()
        self._L = lib(bool(self.cfg.long_streams))
        self._h = self._L.fjsp_oracle_create(C.byref(self.cfg))
        if not self._h:
            raise MemoryError("fjsp_oracle_create failed")
        self.cells = max(1, int(self.cfg.num_cells))
        self.shared_agvs = int(self.cfg.shared_agvs) if int(self.cfg.shared_agvs) >= 2 else 0
        d = dims(self.cells, self.shared_agvs)
        self.obs = np.zeros(d["obs"], dtype=np.float32)
        self.masks = np.zeros(d["mask"], dtype=np.int8)
        self.rewards = np.zeros(d["act"], dtype=np.float64)
        self.flags = np.zeros(4, dtype=np.uint8)
        self.results = np.zeros(d["act"], dtype=np.uint8)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.fjsp_oracle_destroy(h)

    def reset(self, orders):
        """orders: iterable of (n, type, colour) or packed uint32 array."""
        arr = np.asarray(orders)
        if arr.ndim == 2:
            arr = np.array([order_rec(*o) for o in arr], dtype=np.uint32)
        arr = np.ascontiguousarray(arr, dtype=np.uint32)
        self._L.fjsp_oracle_reset(self._h, arr.ctypes.data, int(arr.shape[0]))
        self._L.fjsp_oracle_observe(self._h, self.obs.ctypes.data, self.masks.ctypes.data)
        return self.obs.copy(), self.masks.copy()

    def reset_stream(self, orders, num_initial, seed, genv, episode=0):
        """Long order streams with arrivals: `orders` = the attributes of every order that can ever exist (packed uint32),
        `num_initial` of them exist at reset, the others arrive from the Philox arrival stream (seed, genv, episode)."""
        arr = np.ascontiguousarray(orders, dtype=np.uint32)
        self._L.fjsp_oracle_reset_stream(self._h, arr.ctypes.data, int(arr.shape[0]), int(num_initial), int(seed), int(genv), int(episode))
        self._L.fjsp_oracle_observe(self._h, self.obs.ctypes.data, self.masks.ctypes.data)
        return self.obs.copy(), self.masks.copy()

    def export_orders(self, first, count):
        out = np.zeros((count, 4), dtype=np.int32)
        self._L.fjsp_oracle_export_orders(self._h, int(first), int(count), out.ctypes.data)
        return out

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        self._L.fjsp_oracle_step(self._h, a.ctypes.data, self.obs.ctypes.data, self.masks.ctypes.data,
                                 self.rewards.ctypes.data, self.flags.ctypes.data, self.results.ctypes.data)
        return self.obs.copy(), self.masks.copy(), self.rewards.copy(), self.flags.copy()

    def export(self, cell: int = 0) -> np.ndarray:
        s = np.zeros((), dtype=CANON_DT)
        self._L.fjsp_oracle_export_cell(self._h, int(cell), s.ctypes.data)
        return s


class OracleBatch:
    """N envs with Philox orders/actions and auto-reset (CPU baseline + statistics parity)."""

    def __init__(self, n_envs: int, seed: int, num_orders: int = 30, first_env: int = 0,
                 cfg: FjspConfig | None = None):
        self._L = lib()
        self.cfg = cfg if cfg is not None else default_config()
        self.n, self.seed, self.num_orders, self.first_env = n_envs, seed, num_orders, first_env
        self._ptrs = (C.c_void_p * n_envs)()
        self.episodes = np.zeros(n_envs, dtype=np.uint32)
        orders = np.zeros(32, dtype=np.uint32)
        for i in range(n_envs):
            h = self._L.fjsp_oracle_create(C.byref(self.cfg))
            self._ptrs[i] = h
            self._L.fjsp_oracle_philox_orders(seed, first_env + i, 0, num_orders, orders.ctypes.data)
            self._L.fjsp_oracle_reset(h, orders.ctypes.data, num_orders)
        self.stats = np.zeros(8, dtype=np.uint64)
        self.t = 0

    def __del__(self):
        for i in range(getattr(self, "n", 0)):
            if self._ptrs[i]:
                self._L.fjsp_oracle_destroy(self._ptrs[i])
                self._ptrs[i] = None

    def rollout(self, steps: int, nthreads: int = 1):
        self._L.fjsp_oracle_rollout_random(self._ptrs, self.episodes.ctypes.data, self.n, self.first_env, steps,
                                           self.seed, self.t, self.num_orders, nthreads, self.stats.ctypes.data)
        self.t += steps
        return self.stats.copy()

    def export(self, i: int) -> np.ndarray:
        s = np.zeros((), dtype=CANON_DT)
        self._L.fjsp_oracle_export(self._ptrs[i], s.ctypes.data)
        return s
