"""Observation / action space declarations of the 8 agents, with the reference's shapes
(/root/reference/utils/ObservationSpaces.py:14-105, utils/ActionSpaces.py:10-56).

Uses ``gymnasium.spaces`` when it is importable and the package's shape-only stand-ins otherwise (this image has no
gymnasium); callers only read ``.spaces``, ``.n``, ``.shape`` and ``.sample()`` (a2c.py:118-135,80; train.py:268)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the image
    from gymnasium import spaces  # type: ignore
except Exception:  # noqa: BLE001
    from .compat.gymnasium import spaces  # type: ignore


def pickup_station():
    return spaces.Dict({
        "order_size": spaces.Discrete(21), "products_remaining": spaces.Discrete(21),
        "next_product_type": spaces.Discrete(4), "next_product_color": spaces.Discrete(4),
        "current_tray_type": spaces.Discrete(4), "current_tray_color": spaces.Discrete(4),
        "current_tray_count": spaces.Discrete(6),
        "action_mask": spaces.Box(low=0, high=1, shape=(3,), dtype=np.int8),
    })


def agv(grid_rows=4, grid_cols=6, tray_capacity=5):
    return spaces.Dict({
        "position": spaces.MultiDiscrete([grid_rows, grid_cols]),
        "carrying_tray": spaces.Discrete(2), "tray_product_count": spaces.Discrete(tray_capacity + 1),
        "tray_type": spaces.Discrete(4), "tray_needs_processing": spaces.Discrete(2),
        "tray_needs_packaging": spaces.Discrete(2), "pickup_ready_trays": spaces.Discrete(10),
        "small_machine_busy": spaces.Discrete(2), "big_machine_busy": spaces.Discrete(2),
        "small_machine_ready": spaces.Discrete(10), "big_machine_ready": spaces.Discrete(10),
        "storage_tray_count": spaces.Discrete(100),
        "action_mask": spaces.Box(low=0, high=1, shape=(8,), dtype=np.int8),
    })


def machine():
    return spaces.Dict({
        "is_busy": spaces.Discrete(2),
        "processing_progress": spaces.Box(low=0, high=1, shape=(1,), dtype=np.float32),
        "queue_length": spaces.Discrete(10),
        "action_mask": spaces.Box(low=0, high=1, shape=(3,), dtype=np.int8),
    })


def packaging():
    return spaces.Dict({
        "is_busy": spaces.Discrete(2),
        "processing_progress": spaces.Box(low=0, high=1, shape=(1,), dtype=np.float32),
        "queue_length": spaces.Discrete(20),
        "action_mask": spaces.Box(low=0, high=1, shape=(3,), dtype=np.int8),
    })


def action_space(agent_id: str):
    return spaces.Discrete(8 if agent_id == "agv" else 3)
