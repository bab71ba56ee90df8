"""fjsp-b200: B200-native batched FJSP environment (drop-in for FJSPParallelEnv's step path).

Public surface:
    BatchedFJSPEnv          tensor API over N lockstep envs (multi_agent_rl_for_fjsp_b200/env.py)
    CellViewEnv             a K-cell (scaled shop) env seen as N*K rows of the reference's 8-agent layout
    FJSPParallelEnv         the reference's PettingZoo-style dict API, env 0 of a batch of 1
                            (multi_agent_rl_for_fjsp_b200/dropin/FJSPParallelEnvWrapper.py)
    abi                     ctypes binding of libfjsp_b200.so (include/fjsp_b200.h)

There is no CPU implementation: importing works anywhere, but creating an env requires the compiled
CUDA library and a CUDA device, and fails loudly otherwise.
"""
from . import abi  # noqa: F401
from .env import BatchedFJSPEnv, CellViewEnv, AGENT_IDS, N_ACTIONS, OBS_DIM, MASK_DIM, MASK_OFFSETS  # noqa: F401

__all__ = ["abi", "BatchedFJSPEnv", "CellViewEnv", "AGENT_IDS", "N_ACTIONS", "OBS_DIM", "MASK_DIM", "MASK_OFFSETS"]
