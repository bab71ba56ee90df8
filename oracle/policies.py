"""Action generators used by the parity tests and the golden-vector generator (TEST INFRASTRUCTURE).

All policies read only the flattened observation (layout O, 38 floats) and the 32-byte mask
block, so the same policy can drive the live reference, the C restatement and the CUDA path.
Three regimes: uniform random (config 1), mask-respecting random, and a competent
heuristic with noise (reaches packaging / termination states random play almost never does).
"""
from __future__ import annotations

import numpy as np

N_ACTIONS = (3, 8, 3, 3, 3, 3, 3, 3)
MASK_OFF = (0, 3, 11, 14, 17, 20, 23, 26)

# default layout (constants.py:5-11); AGV move action -> cell
_MOVE_CELL = {1: (0, 0), 2: (2, 3), 3: (0, 3), 4: (3, 0), 5: (3, 5)}


def uniform_random(rs: np.random.RandomState, obs=None, masks=None) -> np.ndarray:
    """a_i ~ U{0..n_i-1} (BASELINE.json config 1)."""
    return np.array([rs.randint(n) for n in N_ACTIONS], dtype=np.uint8)


def masked_random(rs: np.random.RandomState, obs, masks) -> np.ndarray:
    out = np.zeros(8, dtype=np.uint8)
    for i, (n, off) in enumerate(zip(N_ACTIONS, MASK_OFF)):
        valid = np.flatnonzero(masks[off:off + n])
        out[i] = valid[rs.randint(len(valid))]
    return out


def heuristic(rs: np.random.RandomState, obs, masks, noise: float = 0.15, move_cell=_MOVE_CELL) -> np.ndarray:
    out = np.zeros(8, dtype=np.uint8)
    m = masks
    # pickup station: keep loading
    out[0] = 1 if m[1] else 0
    # machines: signal finished trays, else start
    for i, off in ((2, 11), (3, 14)):
        out[i] = 2 if m[off + 2] else (1 if m[off + 1] else 0)
    # packaging: start whenever allowed
    for i, off in ((4, 17), (5, 20), (6, 23), (7, 26)):
        out[i] = 1 if m[off + 1] else 0
    # AGV
    am = m[3:11]
    pos = (int(obs[11]), int(obs[12]))
    carrying = obs[9] > 0

    def goto(move_action, manip):
        if pos == move_cell[move_action]:
            return manip if am[manip] else 0
        return move_action

    if carrying:
        if obs[17] > 0:  # needs processing
            ttype = int(obs[19])
            if ttype == 1:
                tgt = 2
            elif ttype == 3:
                tgt = 3
            else:
                tgt = 2 if obs[22] <= obs[25] else 3  # MEDIUM: shorter queue
            a = goto(tgt, 7)
        elif obs[16] > 0:
            a = goto(5, 7)
        else:
            a = goto(4, 7)
    else:
        if obs[14] > 0:
            a = goto(2, 6)
        elif obs[8] > 0:
            a = goto(3, 6)
        elif obs[10] > 0:
            a = goto(1, 6)
        elif obs[15] > 0:
            a = goto(4, 6)
        else:
            a = goto(1, 6)
    out[1] = a
    if noise > 0:
        for i, n in enumerate(N_ACTIONS):
            if rs.random_sample() < noise:
                out[i] = rs.randint(n)
    return out


POLICIES = {"uniform": uniform_random, "masked": masked_random, "heuristic": heuristic}


def random_orders(rs: np.random.RandomState, num_orders: int) -> np.ndarray:
    """(n, type, colour) rows with the reference's ranges (FJSPSimulation.py:107-112)."""
    return np.stack([rs.randint(1, 10, size=num_orders), rs.randint(1, 4, size=num_orders),
                     rs.randint(1, 4, size=num_orders)], axis=1).astype(np.int64)


# ---- shared floor (include/fjsp_b200.h "shared floor"): pickup station | agv_0 .. agv_{A-1} | machines | stations
def shared_layout(agvs: int):
    """(n_actions, mask offsets, obs offsets of the AGV blocks, obs offset of the machine block)"""
    nact = (3,) + (8,) * agvs + (3,) * 6
    moff = [0, 3]
    for n in nact[1:-1]:
        moff.append(moff[-1] + n)
    return nact, tuple(moff), tuple(7 + 13 * j for j in range(agvs)), 7 + 13 * agvs


def shared_uniform(rs, obs, masks, agvs):
    nact = shared_layout(agvs)[0]
    out = np.zeros((len(nact) + 7) // 8 * 8, dtype=np.uint8)
    out[:len(nact)] = [rs.randint(n) for n in nact]
    return out


def shared_masked(rs, obs, masks, agvs):
    nact, moff, _, _ = shared_layout(agvs)
    out = np.zeros((len(nact) + 7) // 8 * 8, dtype=np.uint8)
    for i, (n, off) in enumerate(zip(nact, moff)):
        valid = np.flatnonzero(masks[off:off + n])
        out[i] = valid[rs.randint(len(valid))]
    return out


def shared_heuristic(rs, obs, masks, agvs, noise=0.1, move_cell=_MOVE_CELL):
    """Every AGV follows the single-AGV heuristic on its own block; a move the occupancy rule forbids is replaced by a
    random allowed move (or by waiting), so AGVs get out of each other's way."""
    nact, moff, aoff, soff = shared_layout(agvs)
    out = np.zeros((len(nact) + 7) // 8 * 8, dtype=np.uint8)
    m = masks
    out[0] = 1 if m[1] else 0
    for i in range(2):   # machines
        off = moff[1 + agvs + i]
        out[1 + agvs + i] = 2 if m[off + 2] else (1 if m[off + 1] else 0)
    for i in range(4):   # packaging stations
        off = moff[3 + agvs + i]
        out[3 + agvs + i] = 1 if m[off + 1] else 0
    small_q, big_q = obs[soff + 2], obs[soff + 5]
    for j in range(agvs):
        o = obs[aoff[j]:aoff[j] + 13]
        am = m[moff[1 + j]:moff[1 + j] + 8]
        pos = (int(o[4]), int(o[5]))
        carrying = o[2] > 0

        def goto(move_action, manip):
            if pos == move_cell[move_action]:
                return manip if am[manip] else 0
            return move_action

        if carrying:
            if o[10] > 0:
                ttype = int(o[12])
                tgt = 2 if ttype == 1 else 3 if ttype == 3 else (2 if small_q <= big_q else 3)
                a = goto(tgt, 7)
            elif o[9] > 0:
                a = goto(5, 7)
            else:
                a = goto(4, 7)
        else:
            want = [2] if o[7] > 0 else []
            want += [3] if o[1] > 0 else []
            want += [1] if o[3] > 0 else []
            want += [4] if o[8] > 0 else []
            want = want[j % len(want):] + want[:j % len(want)] if want else [1]   # AGVs prefer different sources
            a = goto(want[0], 6)
        if 1 <= a <= 5 and not am[a]:   # position taken: step aside or wait
            free = [x for x in range(1, 6) if am[x]]
            a = free[rs.randint(len(free))] if free and rs.random_sample() < 0.5 else 0
        out[1 + j] = a
    if noise > 0:
        for i, n in enumerate(nact):
            if rs.random_sample() < noise:
                out[i] = rs.randint(n)
    return out


SHARED_POLICIES = {"uniform": shared_uniform, "masked": shared_masked, "heuristic": shared_heuristic}
