"""ctypes binding of the TEST-ONLY host build of the device step function (see harness.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle.canon import CANON_DT
from oracle.fjsp_oracle import FjspConfig, default_config, dims, order_rec

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfjsp_hostharness.so")
_CSRC = os.path.join(_HERE, "..", "..", "multi_agent_rl_for_fjsp_b200", "csrc")


def build():
    srcs = [os.path.join(_HERE, "harness.cpp"), os.path.join(_CSRC, "fjsp_core.h"), os.path.join(_CSRC, "fjsp_host.h"),
            os.path.join(_CSRC, "fjsp_shared.h"),
            os.path.join(_HERE, "..", "..", "include", "fjsp_b200.h")]
    if (not os.path.exists(_SO)) or any(os.path.getmtime(p) > os.path.getmtime(_SO) for p in srcs):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas", "-DFJSP_BOUNDS_CHECK",
                        "-o", _SO, srcs[0]], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.hh_create.restype = C.c_void_p
        L.hh_create.argtypes = [C.c_void_p]
        L.hh_check_config.restype = C.c_char_p
        L.hh_check_config.argtypes = [C.c_void_p]
        L.hh_destroy.argtypes = [C.c_void_p]
        L.hh_observe.argtypes = [C.c_void_p] * 3
        L.hh_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint32]
        L.hh_step.argtypes = [C.c_void_p] * 8
        L.hh_step_wire.argtypes = [C.c_void_p] * 7
        L.hh_step_cells.argtypes = [C.c_void_p] * 8 + [C.c_int]
        L.hh_export.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hh_words.argtypes = [C.c_void_p, C.c_void_p]
        L.hh_export_orders.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.hh_philox_actions.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        L.hh_philox.argtypes = [C.c_void_p] * 3
        L.hh_philox_actions_shared.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        _lib = L
    return _lib


class HostEnv:
    def __init__(self, cfg: FjspConfig | None = None):
        self._L = lib()
        self.cfg = cfg if cfg is not None else default_config()
        self._h = self._L.hh_create(C.addressof(self.cfg))
        if not self._h:
            raise ValueError(self._L.hh_check_config(C.addressof(self.cfg)).decode())
        self.cells = max(1, int(self.cfg.num_cells))
        self.shared_agvs = int(self.cfg.shared_agvs) if int(self.cfg.shared_agvs) >= 2 else 0
        d = dims(self.cells, self.shared_agvs)
        self.obs = np.zeros(d["obs"], np.float32)
        self.masks = np.zeros(d["mask"], np.int8)
        self.rewards = np.zeros(d["act"], np.float32)
        self.flags = np.zeros(4, np.uint8)
        self.results = np.zeros(d["act"], np.uint8)
        self.infos = np.zeros(4, np.int32)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.hh_destroy(h)

    def reset(self, orders=None, num_orders=None, seed=0, genv=0, episode=0):
        if orders is not None:
            arr = np.asarray(orders)
            if arr.ndim == 2:
                arr = np.array([order_rec(*o) for o in arr], dtype=np.uint32)
            arr = np.ascontiguousarray(arr, dtype=np.uint32)
            buf = np.zeros(max(32, arr.shape[0]), np.uint32)
            buf[:arr.shape[0]] = arr
            # (seed / genv / episode still select the ARRIVAL stream of a long-streams env)
            self._L.hh_reset(self._h, buf.ctypes.data, int(arr.shape[0] if num_orders is None else num_orders), seed, genv, episode)
        else:
            self._L.hh_reset(self._h, None, int(num_orders), seed, genv, episode)
        self._L.hh_observe(self._h, self.obs.ctypes.data, self.masks.ctypes.data)
        return self.obs.copy(), self.masks.copy()

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        self._L.hh_step(self._h, a.ctypes.data, self.obs.ctypes.data, self.masks.ctypes.data, self.rewards.ctypes.data,
                        self.flags.ctypes.data, self.results.ctypes.data, self.infos.ctypes.data)
        return self.obs.copy(), self.masks.copy(), self.rewards.copy(), self.flags.copy()

    def step_cells(self, actions, reverse=False):
        """The same step through the cell-parallel phase functions (K >= 2), lanes emulated one after the other."""
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        rc = self._L.hh_step_cells(self._h, a.ctypes.data, self.obs.ctypes.data, self.masks.ctypes.data, self.rewards.ctypes.data,
                                   self.flags.ctypes.data, self.results.ctypes.data, self.infos.ctypes.data, int(reverse))
        assert rc == 1, rc
        return self.obs.copy(), self.masks.copy(), self.rewards.copy(), self.flags.copy()

    def step_wire(self, actions):
        """step() that also returns the wire row (include/fjsp_b200.h) of the same step."""
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        words = (2 + 5 * self.cells + 1) // 2 * 2  # FJSP_WIRE_WORDS_K
        wire = np.zeros(words, np.uint32)
        self._L.hh_step_wire(self._h, a.ctypes.data, self.obs.ctypes.data, self.masks.ctypes.data, self.rewards.ctypes.data,
                             self.flags.ctypes.data, wire.ctypes.data)
        return self.obs.copy(), self.masks.copy(), self.rewards.copy(), self.flags.copy(), wire

    def export(self, cell=0):
        s = np.zeros((), dtype=CANON_DT)
        self._L.hh_export(self._h, int(cell), s.ctypes.data)
        return s

    def export_orders(self, first, count):
        out = np.zeros((count, 4), np.int32)
        self._L.hh_export_orders(self._h, int(first), int(count), out.ctypes.data, None)
        return out

    def words(self):
        k = self.cells
        if self.shared_agvs:
            w = np.zeros(132, np.uint32)
            self._L.hh_words(self._h, w.ctypes.data)
            return w
        w = np.zeros((228 + 64 * k + 24 * (k - 1)) if self.cfg.long_streams else (64 + 64 * k + 20 * (k - 1)), np.uint32)
        self._L.hh_words(self._h, w.ctypes.data)
        return w
