"""Wire rows (include/fjsp_b200.h): the compact form in which a step's results cross PCIe on the host-buffer path.
CPU-only: rows produced by the packed-state core (test-only host build of the device step function) are decoded by the
product library's HOST function fjsp_wire_decode and must give bit for bit the float32 / int8 tensors of the same step —
on the reference's own golden trajectories (so the decoded tensors are also checked against the reference) and on
scaled shops under a policy that packages products (progress table, int8 queue lengths, negative rewards)."""
import ctypes as C
import os

import numpy as np
import pytest

from multi_agent_rl_for_fjsp_b200 import abi
from oracle import policies
from oracle.fjsp_oracle import default_config, dims
from tests.host_harness.hostharness import HostEnv
from tests.util import cfg_from_dict, load_golden


def _abi_cfg(ocfg):
    cfg = abi.FjspConfig()
    C.memmove(C.addressof(cfg), C.addressof(ocfg), C.sizeof(cfg))
    return cfg


def aligned(shape, dtype, fill, align=64, offset=0):
    """Array whose data pointer is `offset` bytes past an `align`-byte boundary (32-byte alignment of all four outputs
    selects the decoder's streaming-store path for blocks of 8 rows; anything else the plain-store path)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + align + offset, np.uint8)
    start = (-raw.ctypes.data) % align + offset
    out = raw[start:start + n].view(dtype).reshape(shape)
    out[...] = fill
    return out


def decode(cfg, rows, threads=1, offset=0):
    k = int(cfg.num_cells)
    d = dims(k)
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    n = rows.shape[0]
    assert rows.shape[1] * 4 == abi.lib().fjsp_wire_row_bytes(k) == 4 * abi.dims(k)["wire_words"]
    obs = aligned((n, d["obs"]), np.float32, -7, offset=offset)
    masks = aligned((n, d["mask"]), np.int8, -7, offset=offset)
    rew = aligned((n, d["act"]), np.float32, -7, offset=offset)
    flags = aligned((n, 4), np.uint8, 77, offset=offset)
    abi.check(abi.lib().fjsp_wire_decode(C.byref(cfg), rows.ctypes.data, n, obs.ctypes.data, masks.ctypes.data, rew.ctypes.data,
                                         flags.ctypes.data, threads))
    return obs, masks, rew, flags


@pytest.mark.parametrize("name", ["default_heuristic", "pack_cap3", "speed2_far_step20"])
def test_wire_rows_decode_to_the_golden_tensors(name):
    g, cfgd = load_golden(name)
    ocfg = cfg_from_dict(cfgd)
    starts = g["ep_start"].tolist() + [g["actions"].shape[0]]
    e = HostEnv(ocfg)
    rows, want = [], []
    for ep in range(min(3, len(starts) - 1)):
        no = int(g["ep_norders"][ep])
        e.reset(g["ep_orders"][ep][:no])
        for t in range(starts[ep], starts[ep + 1]):
            o, m, r, f, w = e.step_wire(g["actions"][t])
            assert np.array_equal(o, g["obs"][t]) and np.array_equal(m, g["masks"][t])
            rows.append(w), want.append((o, m, r, f))
    obs, masks, rew, flags = decode(_abi_cfg(ocfg), np.stack(rows), threads=3)
    for i, (o, m, r, f) in enumerate(want):
        assert np.array_equal(obs[i], o), (i, np.flatnonzero(obs[i] != o))
        assert np.array_equal(masks[i], m) and np.array_equal(rew[i], r) and np.array_equal(flags[i][:3], f[:3]), i
    assert np.abs(rew - g["rewards"][:len(want)].astype(np.float32)).max() <= 1e-6 * np.abs(g["rewards"]).max()


@pytest.mark.parametrize("k", [1, 2, 4])
def test_wire_rows_scaled_shop(k):
    from tests.test_scaled_shop import cfg_k, policy_actions

    ocfg = cfg_k(k)
    e = HostEnv(ocfg)
    rs = np.random.RandomState(k)
    rows, want = [], []
    for ep in range(4):
        o, m = e.reset(policies.random_orders(rs, 30))
        for t in range(201):
            o, m, r, f, w = e.step_wire(policy_actions(rs, o, m, k, 2 if ep else 0))
            rows.append(w), want.append((o, m, r, f))
            if f[0] or f[1]:
                break
    obs, masks, rew, flags = decode(_abi_cfg(ocfg), np.stack(rows), threads=1 + k)
    o2, m2, r2, f2 = decode(_abi_cfg(ocfg), np.stack(rows), threads=1, offset=8)   # unaligned outputs: plain stores
    assert np.array_equal(obs, o2) and np.array_equal(masks, m2) and np.array_equal(rew, r2) and np.array_equal(flags, f2)
    assert any(o[7 + 20] > 0 or o[7 + 23] > 0 or o[7 + 26] > 0 or o[7 + 29] > 0 for o, _, _, _ in want), "no packaging progress seen"
    for i, (o, m, r, f) in enumerate(want):
        assert np.array_equal(obs[i], o) and np.array_equal(masks[i], m) and np.array_equal(rew[i], r), i
        assert np.array_equal(flags[i][:3], f[:3])


def test_wire_decode_partial_outputs_and_errors():
    cfg = abi.default_config()
    rows = np.zeros((5, abi.dims(1)["wire_words"]), np.uint32)
    L = abi.lib()
    obs = np.ones((5, 38), np.float32)
    assert L.fjsp_wire_decode(C.byref(cfg), rows.ctypes.data, 5, obs.ctypes.data, None, None, None, 1) == 0
    assert (obs[:, :11] == 0).all()
    assert L.fjsp_wire_decode(C.byref(cfg), None, 5, obs.ctypes.data, None, None, None, 1) != 0
    assert L.fjsp_wire_row_bytes(0) == 0 and L.fjsp_wire_row_bytes(5) == 0 and L.fjsp_wire_row_bytes(1) == 32 and L.fjsp_wire_row_bytes(4) == 88


def numpy_decode(cfg, rows):
    """Independent NumPy statement of the wire format, written from the description in include/fjsp_b200.h — not the
    library's code."""
    k = int(cfg.num_cells)
    d = dims(k)
    A, obs_n, mask_n, act_n = d["agents"], d["obs"], d["mask"], d["act"]
    n = rows.shape[0]
    r = rows.astype(np.int64)

    def f(word, shift, width):
        return (r[:, word] >> shift) & ((1 << width) - 1)

    pos = np.array([[cfg.pos[i][0], cfg.pos[i][1]] for i in range(5)] + [[cfg.pos[0][0], cfg.pos[0][1]]] * 3, np.float32)
    prog = np.zeros(256, np.float32)
    prog[1:] = ((1.0 / np.arange(1, 256, dtype=np.float64)) * 100.0).astype(np.float32)
    i8 = lambda x: x.astype(np.uint8).view(np.int8).astype(np.float32)  # noqa: E731
    obs = np.zeros((n, obs_n), np.float32)
    masks = np.zeros((n, mask_n), np.int8)
    local = np.zeros((n, act_n), np.int64)
    lut = {"ps": np.array([0, -10, 10, 60]), "agv": np.array([0, -1, 20, 120, -50, 0, 0, 0]),
           "m": np.array([0, -20, 10, 50]), "p": np.array([0, -10, 20, 200])}
    # word 0: pickup station, flags, its mask bits and reward code; word 1: counters
    for j, (sh, wd) in enumerate(((0, 2), (2, 3), (5, 2), (7, 2), (9, 2), (11, 4), (15, 4))):
        obs[:, j] = f(0, sh, wd)
    fb = f(0, 19, 6)
    flags = np.stack([fb & 1, (fb >> 1) & 1, (fb >> 2) & 7, (fb >> 5) & 1], axis=1).astype(np.uint8)
    masks[:, 0], masks[:, 1], masks[:, 2] = 1, f(0, 25, 1), f(0, 26, 1)
    local[:, 0] = lut["ps"][f(0, 27, 2)]
    products, orders, ready = f(1, 0, 10), f(1, 10, 9), f(1, 19, 13)
    g = 10 * (100 * orders + 10 * products) - int(cfg.step_size)
    for c in range(k):
        w, b, mo, ac = 2 + 5 * c, 7 + 31 * c, 3 + 26 * c, 1 + 7 * c
        loc = f(w, 0, 3)
        carrying = f(w, 3, 1)
        small = (f(w + 1, 8, 1), f(w + 1, 9, 1), f(w + 1, 10, 6))
        big = (f(w + 1, 16, 1), f(w + 1, 17, 1), f(w + 1, 18, 6))
        agv = [big[0], f(w, 10, 6), carrying, ready, pos[loc, 0], pos[loc, 1], small[0], f(w, 16, 6), f(w + 1, 0, 8), carrying,
               f(w, 4, 1), f(w, 5, 3), f(w, 8, 2)]
        for j, v in enumerate(agv + list(small) + list(big)):
            obs[:, b + j] = v
        masks[:, mo] = 1
        for j in range(7):
            masks[:, mo + 1 + j] = f(w, 22 + j, 1)
        local[:, ac] = lut["agv"][f(w, 29, 3)]
        for i, (sh, csh) in enumerate(((24, 28), (26, 30))):
            masks[:, mo + 8 + 3 * i] = 1
            masks[:, mo + 9 + 3 * i], masks[:, mo + 10 + 3 * i] = f(w + 1, sh, 1), f(w + 1, sh + 1, 1)
            local[:, ac + 1 + i] = lut["m"][f(w + 1, csh, 2)]
        # packaging_blue_1, _blue_2, _red: one word each; packaging_green: shared out over the spare bits of the first two
        stations = [(f(w + 2 + i, 0, 1), f(w + 2 + i, 1, 8), f(w + 2 + i, 9, 8), f(w + 2 + i, 17, 1), f(w + 2 + i, 18, 1), f(w + 2 + i, 19, 2))
                    for i in range(3)]
        stations.append((f(w + 2, 29, 1), f(w + 2, 21, 8), f(w + 3, 21, 8), f(w + 2, 30, 1), f(w + 2, 31, 1), f(w + 3, 29, 2)))
        for i, (busy, L, q, m1, m2, code) in enumerate(stations):
            obs[:, b + 19 + 3 * i], obs[:, b + 20 + 3 * i], obs[:, b + 21 + 3 * i] = busy, prog[L], i8(q)
            masks[:, mo + 14 + 3 * i], masks[:, mo + 15 + 3 * i], masks[:, mo + 16 + 3 * i] = 1, m1, m2
            local[:, ac + 3 + i] = lut["p"][code]
    num = (g[:, None] + A * local).astype(np.float32)
    rew = (num / np.float32(10 * A)).astype(np.float32)
    rew[:, A:] = 0.0
    return obs, masks, rew, flags


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_wire_decode_on_arbitrary_rows_matches_numpy_statement(k):
    """Every byte value in every field (int8 wrap of queue lengths, all 256 progress indices, out-of-range station
    indices, negative numerators, all flag bits): library decoder (AVX2 + streaming stores on aligned outputs, plain
    stores on unaligned ones, several thread splits) == an independent NumPy statement of the format."""
    cfg = abi.default_config()
    cfg.num_cells = k
    rs = np.random.RandomState(10 + k)
    n = 4099
    words = abi.dims(k)["wire_words"]
    rows = rs.randint(0, 1 << 32, size=(n, words), dtype=np.uint64).astype(np.uint32)
    rows[:64] = 0
    rows[64:128] = 0xFFFFFFFF
    want = numpy_decode(cfg, rows)
    for threads, offset in ((1, 0), (3, 0), (7, 0), (1, 4), (2, 16)):
        got = decode(cfg, rows, threads=threads, offset=offset)
        for name, a, b in zip(("obs", "masks", "rewards", "flags"), got, want):
            assert np.array_equal(a, b), (k, threads, offset, name, np.argwhere(a != b)[:4])


def test_wire_decode_generic_body_equals_avx2_body():
    """The portable decoder body (FJSP_DECODE_GENERIC=1, read once per process) in a subprocess == this process's."""
    import subprocess
    import sys

    if abi.lib() is None:
        pytest.skip("library not built")
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from tests.test_wire_format import decode, abi\n"
        "cfg = abi.default_config(); cfg.num_cells = 4\n"
        "rows = np.random.RandomState(5).randint(0, 1 << 32, size=(1000, abi.dims(4)['wire_words']), dtype=np.uint64).astype(np.uint32)\n"
        "import hashlib; print(hashlib.sha256(b''.join(np.ascontiguousarray(x).tobytes() for x in decode(cfg, rows, threads=2))).hexdigest())\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for env in ({}, {"FJSP_DECODE_GENERIC": "1"}):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, check=True)
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1] and len(outs[0]) == 64
