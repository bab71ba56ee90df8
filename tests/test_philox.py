"""Philox4x32-10 streams: known-answer vectors (Random123 kat_vectors) and agreement of the three
implementations (numpy reference written here, oracle C, device code built for the host)."""
import ctypes as C

import numpy as np

from oracle import fjsp_oracle
from tests.host_harness import hostharness

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox_np(ctr, key):
    c = [int(x) for x in ctr]
    k = [int(x) for x in key]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xffffffff, p1 & 0xffffffff, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xffffffff, p0 & 0xffffffff]
        k = [(k[0] + W0) & 0xffffffff, (k[1] + W1) & 0xffffffff]
    return np.array(c, dtype=np.uint32)


KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def hh_philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, np.uint32)
    hostharness.lib().hh_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def test_known_answers():
    for ctr, key, want in KAT:
        assert philox_np(ctr, key).tolist() == list(want)
        assert fjsp_oracle.philox(ctr, key).tolist() == list(want)
        assert hh_philox(ctr, key).tolist() == list(want)


def test_streams_agree_and_are_in_range():
    rs = np.random.RandomState(0)
    nact = (3, 8, 3, 3, 3, 3, 3, 3)
    counts = np.zeros((8, 8), dtype=np.int64)
    for _ in range(400):
        seed = int(rs.randint(0, 2**62))
        genv = int(rs.randint(0, 2**32))
        t = int(rs.randint(0, 2**40))
        r = philox_np((genv, t & 0xffffffff, t >> 32, 1), (seed & 0xffffffff, seed >> 32))
        want = [((int(r[j >> 1]) >> 16 if j & 1 else int(r[j >> 1]) & 0xffff) * nact[j]) >> 16 for j in range(8)]
        a = fjsp_oracle.philox_actions(seed, genv, t)
        b = np.zeros(8, np.uint8)
        hostharness.lib().hh_philox_actions(seed, genv, t, 1, b.ctypes.data)
        assert a.tolist() == want and b.tolist() == want
        for j in range(8):
            counts[j, want[j]] += 1
        ep = int(rs.randint(0, 1000))
        orders = fjsp_oracle.philox_orders(seed, genv, ep, 32)
        for o, rec in enumerate(orders):
            r = philox_np((genv, ep, o, 0), (seed & 0xffffffff, seed >> 32))
            n, ty, co = 1 + ((int(r[0]) * 9) >> 32), 1 + ((int(r[1]) * 3) >> 32), 1 + ((int(r[2]) * 3) >> 32)
            assert int(rec) == n | (ty << 8) | (co << 16) and 1 <= n <= 9
    assert (counts[1] > 20).all() and (counts[0, :3] > 90).all() and counts[0, 3:].sum() == 0


def test_device_reset_stream_equals_host_replay():
    """reset_env's Philox path (device code, host build) draws exactly the orders the host replays."""
    seed, genv, ep = 0x1234567890, 4242, 7
    e = hostharness.HostEnv()
    e.reset(orders=None, num_orders=30, seed=seed, genv=genv, episode=ep)
    o = fjsp_oracle.OracleEnv()
    o.reset(fjsp_oracle.philox_orders(seed, genv, ep, 30))
    a, b = e.export(), o.export()
    assert a["num_orders"] == 30
    # same orders => same behaviour: load every product of order 0 and compare sizes
    w = e.words()
    for i, rec in enumerate(fjsp_oracle.philox_orders(seed, genv, ep, 30)):
        n, ty, co = int(rec) & 0xff, (int(rec) >> 8) & 0xff, (int(rec) >> 16) & 0xff
        assert int(w[32 + i]) == n | (ty << 4) | (co << 6)
    assert int(w[6]) == ep
