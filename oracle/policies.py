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
