#!/usr/bin/env python
"""Host-only: time fjsp_wire_decode (wire rows -> float32/int8 tensors) on this box's cores, with and without
streaming stores.  Tells how much of the e2e step the host decode can sustain (no GPU used).

    python tools/wire_decode_bench.py [envs]
"""
import ctypes as C
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(n):
    import numpy as np

    from multi_agent_rl_for_fjsp_b200 import abi

    L, cfg = abi.lib(), abi.default_config()
    rows = np.random.randint(0, 5, size=(n, 32), dtype=np.uint8).view(np.uint32).reshape(n, 8)   # 32-byte wire rows (K = 1)

    def al(shape, dt):
        nb = int(np.prod(shape)) * np.dtype(dt).itemsize
        raw = np.zeros(nb + 64, np.uint8)
        s = (-raw.ctypes.data) % 64
        return raw[s:s + nb].view(dt).reshape(shape)

    obs, masks, rew, flags = al((n, 38), np.float32), al((n, 32), np.int8), al((n, 8), np.float32), al((n, 4), np.uint8)
    out = {}
    for th in (1, 2, 4, 8, 16, 32):
        if th > 2 * (os.cpu_count() or 1):
            break
        best = 1e9
        for _ in range(5):
            t = time.perf_counter()
            L.fjsp_wire_decode(C.byref(cfg), rows.ctypes.data, n, obs.ctypes.data, masks.ctypes.data, rew.ctypes.data, flags.ctypes.data, th)
            best = min(best, time.perf_counter() - t)
        out[th] = round(best * 1e3, 3)
    print("nt_stores=%s envs=%d ms per decode by threads: %s" % ("off" if os.environ.get("FJSP_DECODE_NO_NT") else "on", n, out))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1] != "--child" else 1 << 20
    if "--child" in sys.argv:
        run(int(sys.argv[-1]))
    else:
        for env in ({}, {"FJSP_DECODE_NO_NT": "1"}):
            subprocess.run([sys.executable, __file__, "--child", str(n)], env=dict(os.environ, **env), check=True)
