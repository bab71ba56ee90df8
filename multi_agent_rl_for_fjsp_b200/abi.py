"""ctypes binding of the C ABI declared in include/fjsp_b200.h (libfjsp_b200.so).

The library is built in-tree by ``build()`` (``nvcc -gencode arch=compute_100a,code=sm_100a``) and is the
only implementation of the step: if it is missing, ``lib()`` raises — there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_PKG)
SO_PATH = os.path.join(_PKG, "lib", "libfjsp_b200.so")
if os.environ.get("FJSP_B200_LIB"):   # another build of the same sources (A/B measurements of a kernel change)
    SO_PATH = os.path.abspath(os.environ["FJSP_B200_LIB"])
SOURCES = [os.path.join(_PKG, "csrc", f) for f in ("fjsp_api.cu", "fjsp_wire.cpp", "fjsp_kernels.cuh", "fjsp_a2c.cuh", "fjsp_umma.cuh", "fjsp_core.h", "fjsp_host.h", "fjsp_wire.h", "fjsp_shared.h", "fjsp_shared.cuh")]
HEADER = os.path.join(_REPO, "include", "fjsp_b200.h")

NUM_AGENTS, OBS_DIM, MASK_DIM, FLAG_DIM, INFO_DIM, MAX_ORDERS = 8, 38, 32, 4, 4, 32
STATE_WORDS, TILE_ENVS = 128, 64
MAX_CELLS = 4
LONG_RING, LONG_MAX_ORDERS, LONG_MAX_STEPS = 64, 4095, 65000
ABI_VERSION = 4


def dims(cells: int = 1, long_streams: bool = False, shared_agvs: int = 0) -> dict:
    """Row widths of a K-cell shop (include/fjsp_b200.h FJSP_*_K); K = 1 is the reference shop (8 / 38 / 32 / 128).
    shared_agvs >= 2: the shared floor (FJSP_SHARED_*: 7 + A agents, no wire rows)."""
    if shared_agvs >= 2:
        a = shared_agvs
        return {"agents": 7 + a, "act": (7 + a + 7) // 8 * 8, "obs": 25 + 13 * a, "mask": (21 + 8 * a + 15) // 16 * 16,
                "mask_used": 21 + 8 * a, "state_words": 132, "wire_words": 0}
    agents = 1 + 7 * cells
    return {"agents": agents, "act": (agents + 7) // 8 * 8, "obs": 7 + 31 * cells, "mask": (3 + 26 * cells + 31) // 32 * 32,
            "mask_used": 3 + 26 * cells,
            "state_words": (228 + 64 * cells + 24 * (cells - 1)) if long_streams else (64 + 64 * cells + 20 * (cells - 1)),
            "wire_words": (2 + 5 * cells + 1) // 2 * 2}

CANON_MAXQ, CANON_PS_READY, CANON_MAXPQ = 64, 256, 256

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


class FjspConfig(C.Structure):
    """Mirror of ``FjspConfig`` (constants.py:5-32 of the reference)."""
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


_MACHINE_DT = np.dtype([
    ("is_busy", "<i4"), ("current_tray", "<i4"), ("progress_done", "<i4"),
    ("queue_n", "<i4"), ("queue", "<i4", (CANON_MAXQ,)), ("ready_n", "<i4"), ("ready", "<i4", (CANON_MAXQ,))])
_PACK_DT = np.dtype([
    ("is_busy", "<i4"), ("current_product", "<i4"), ("progress_L", "<i4"), ("products_completed", "<i4"),
    ("users", "<i4"), ("queue_n", "<i4"), ("queue", "<i4", (CANON_MAXPQ,))])
CANON_DT = np.dtype([
    ("current_step", "<i4"), ("num_orders", "<i4"), ("fault", "<i4"),
    ("agv_row", "<i4"), ("agv_col", "<i4"), ("agv_carry", "<i4"), ("agv_is_moving", "<i4"),
    ("ps_order_queue_len", "<i4"), ("ps_current_order", "<i4"), ("ps_product_idx", "<i4"),
    ("ps_current_tray", "<i4"), ("ps_trays_at_station", "<i4"),
    ("ps_ready_n", "<i4"), ("ps_ready", "<i4", (CANON_PS_READY,)),
    ("machine", _MACHINE_DT, (2,)), ("storage_n", "<i4"), ("storage", "<i4", (CANON_MAXQ,)),
    ("pack", _PACK_DT, (4,)),
    ("processed_mask", "<i4", (MAX_ORDERS,)), ("packaged_mask", "<i4", (MAX_ORDERS,)),
    ("order_complete", "<i4", (MAX_ORDERS,)), ("order_completion_step", "<i4", (MAX_ORDERS,)),
    ("total_products_packaged", "<i4"), ("completed_orders", "<i4")])

EXPORTS = [
    "fjsp_last_error", "fjsp_abi_version", "fjsp_default_config", "fjsp_create", "fjsp_destroy", "fjsp_num_envs",
    "fjsp_state_bytes", "fjsp_state_ptr", "fjsp_reset", "fjsp_step", "fjsp_step_host", "fjsp_random_actions",
    "fjsp_rollout_random", "fjsp_export_state", "fjsp_export_state_cell", "fjsp_export_packed", "fjsp_launch_count",
    "fjsp_num_cells", "fjsp_step_wire", "fjsp_step_host_wire", "fjsp_wire_decode", "fjsp_wire_row_bytes", "fjsp_set_decode_threads",
    "fjsp_state_total_bytes", "fjsp_state_save", "fjsp_state_load",
    "fjsp_a2c_sample", "fjsp_a2c_counter_add", "fjsp_a2c_gae", "fjsp_cells_pack_actions", "fjsp_cells_unpack_views",
    "fjsp_a2c_gemm", "fjsp_a2c_loss_grad", "fjsp_export_orders", "fjsp_a2c_gemm_pack", "fjsp_a2c_wgrad_small",
    "fjsp_host_stream_write_probe", "fjsp_a2c_clip_adam", "fjsp_a2c_layer1", "fjsp_a2c_head_backward",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libfjsp_b200.so")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a.  Cross-compiles without a GPU."""
    deps = SOURCES + [HEADER]
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(p) > os.path.getmtime(SO_PATH) for p in deps)
    if force or stale:
        os.makedirs(os.path.dirname(SO_PATH), exist_ok=True)
        cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH, SOURCES[0], SOURCES[1]]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            print(proc.stderr)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    return SO_PATH


_lib = None


def lib() -> C.CDLL:
    """Load libfjsp_b200.so (raises if it has not been built: there is no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        try:  # fresh checkout: compile in-tree once (needs nvcc); there is no other implementation to fall back to
            build()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(
                "libfjsp_b200.so is not built (%s) and could not be compiled here: %s. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'`. This package has no CPU fallback." % (SO_PATH, e))
    L = C.CDLL(SO_PATH)
    vp, i64, u64 = C.c_void_p, C.c_int64, C.c_uint64
    L.fjsp_last_error.restype = C.c_char_p
    L.fjsp_abi_version.restype = C.c_int
    L.fjsp_default_config.argtypes = [C.POINTER(FjspConfig)]
    L.fjsp_create.argtypes = [C.POINTER(FjspConfig), i64, i64, C.c_int, C.POINTER(vp)]
    L.fjsp_destroy.argtypes = [vp]
    L.fjsp_num_envs.restype, L.fjsp_num_envs.argtypes = i64, [vp]
    L.fjsp_state_bytes.restype, L.fjsp_state_bytes.argtypes = C.c_size_t, [vp]
    L.fjsp_state_ptr.restype, L.fjsp_state_ptr.argtypes = vp, [vp]
    L.fjsp_launch_count.restype, L.fjsp_launch_count.argtypes = i64, [vp]
    L.fjsp_reset.argtypes = [vp, vp, u64, vp, C.c_int, vp, vp, vp]
    L.fjsp_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp]
    L.fjsp_step_host.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, vp]
    L.fjsp_random_actions.argtypes = [vp, u64, u64, vp, vp]
    L.fjsp_rollout_random.argtypes = [vp, C.c_int, u64, u64, vp, vp]
    L.fjsp_state_total_bytes.restype, L.fjsp_state_total_bytes.argtypes = C.c_size_t, [vp]
    L.fjsp_state_save.argtypes = [vp, vp, C.c_size_t, vp]
    L.fjsp_state_load.argtypes = [vp, vp, C.c_size_t, vp]
    L.fjsp_a2c_sample.argtypes = [vp, vp, vp, vp, i64, i64, u64, vp, u64, vp]
    L.fjsp_a2c_counter_add.argtypes = [vp, u64, vp]
    L.fjsp_a2c_gae.argtypes = [vp, vp, vp, vp, vp, C.c_int, i64, C.c_float, C.c_float, vp]
    L.fjsp_a2c_loss_grad.argtypes = [vp] * 8 + [C.c_float, i64, vp, vp, vp, vp]
    L.fjsp_a2c_gemm.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.fjsp_a2c_gemm_pack.argtypes = [vp, C.c_int, vp]
    L.fjsp_a2c_wgrad_small.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.fjsp_a2c_head_backward.argtypes = [vp, C.c_int, C.c_int, vp]
    L.fjsp_a2c_layer1.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.fjsp_a2c_clip_adam.argtypes = [vp, C.c_int, C.c_int, vp, C.c_float, C.c_double, C.c_double, C.c_double, vp]
    L.fjsp_host_stream_write_probe.argtypes = [vp, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
    L.fjsp_cells_pack_actions.argtypes = [vp, vp, i64, C.c_int, vp]
    L.fjsp_cells_unpack_views.argtypes = [vp] * 8 + [i64, C.c_int, vp]
    L.fjsp_export_state.argtypes = [vp, i64, vp]
    L.fjsp_export_state_cell.argtypes = [vp, i64, C.c_int, vp]
    L.fjsp_export_packed.argtypes = [vp, i64, vp]
    L.fjsp_export_orders.argtypes = [vp, i64, C.c_int, C.c_int, vp, vp]
    L.fjsp_num_cells.restype, L.fjsp_num_cells.argtypes = C.c_int, [vp]
    L.fjsp_set_decode_threads.argtypes = [vp, C.c_int]
    L.fjsp_step_host_wire.argtypes = [vp, vp, vp, C.c_int, vp]
    L.fjsp_step_wire.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp]
    L.fjsp_wire_decode.argtypes = [C.POINTER(FjspConfig), vp, i64, vp, vp, vp, vp, C.c_int]
    L.fjsp_wire_row_bytes.restype, L.fjsp_wire_row_bytes.argtypes = C.c_size_t, [C.c_int]
    if L.fjsp_abi_version() != ABI_VERSION:
        raise RuntimeError("libfjsp_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise RuntimeError("libfjsp_b200: " + lib().fjsp_last_error().decode())


def default_config() -> FjspConfig:
    cfg = FjspConfig()
    check(lib().fjsp_default_config(C.byref(cfg)))
    return cfg


def config_from_dict(d: dict | None) -> FjspConfig:
    """Build an FjspConfig from a reference-style CONFIG dict (constants.py:20-32); unknown keys are ignored,
    ``positions`` / ``processing_times`` may carry LOCATION_POSITIONS / PROCESSING_TIMES overrides."""
    cfg = default_config()
    if not d:
        return cfg
    for k in ("grid_rows", "grid_cols", "step_size", "agv_speed", "max_episode_steps", "storage_capacity",
              "pack_capacity", "tray_capacity", "num_trays", "proc_small", "proc_big", "proc_pack", "num_cells", "long_streams",
              "arrival_prob_q16", "arrival_max_orders"):
        if k in d:
            setattr(cfg, k, int(d[k]))
    pt = d.get("processing_times") or {}
    for k, f in (("small_machine", "proc_small"), ("big_machine", "proc_big"), ("packaging", "proc_pack")):
        if k in pt:
            setattr(cfg, f, int(pt[k]))
    pos = d.get("positions") or d.get("pos")
    if pos is not None:
        for i, (r, c) in enumerate(pos):
            cfg.pos[i][0], cfg.pos[i][1] = int(r), int(c)
    return cfg


OP_KC, OP_KCS, OP_MC, OP_PK = 0, 1, 2, 3
PACK_JOB_DT = np.dtype([("src", "<u8"), ("dst", "<u8"), ("op", "<i4"), ("ld", "<i4"), ("N", "<i4"), ("K", "<i4")])  # FjspPackJob


WGRAD_JOB_DT = np.dtype([("X", "<u8"), ("Y", "<u8"), ("G", "<u8"), ("B", "<i4"), ("nx", "<i4"), ("ny", "<i4"), ("ldx", "<i4"),
                         ("ldy", "<i4"), ("gsi", "<i4"), ("gsj", "<i4"), ("reserved", "<i4", (3,))])  # FjspWgradJob (64 B)


HEAD_BWD_JOB_DT = np.dtype([("dl", "<u8"), ("W", "<u8"), ("H", "<u8"), ("dH", "<u8"), ("gW", "<u8"), ("gb", "<u8"), ("rows", "<i4"),
                            ("n", "<i4"), ("na", "<i4"), ("ld_dl", "<i4")])  # FjspHeadBwdJob (64 B)
assert HEAD_BWD_JOB_DT.itemsize == 64
LAYER1_JOB_DT = np.dtype([("X", "<u8"), ("W", "<u8"), ("bias", "<u8"), ("Y", "<u8"), ("rows", "<i4"), ("k", "<i4"), ("n", "<i4"),
                          ("ldx", "<i4"), ("ldy", "<i4"), ("relu", "<i4"), ("reserved", "<i4", (2,))])  # FjspLayer1Job (64 B)
assert LAYER1_JOB_DT.itemsize == 64
OPT_SEG_DT = np.dtype([("param", "<u8"), ("grad", "<u8"), ("m", "<u8"), ("v", "<u8"), ("step", "<u8"), ("n", "<i4"), ("net", "<i4"),
                       ("lr", "<f4"), ("bump", "<i4"), ("reserved", "<i8")])  # FjspOptSeg (64 B)
assert OPT_SEG_DT.itemsize == 64


def pack_image_floats(n: int, k: int) -> int:
    return (k + 15) // 16 * 32 * ((n + 15) // 16 * 16)
GEMM_RELU, GEMM_ATOMIC = 1, 2
GEMM_PROB_DT = np.dtype([
    ("A", "<u8"), ("B", "<u8"), ("C", "<u8"), ("bias", "<u8"), ("mask", "<u8"), ("colsum", "<u8"),
    ("M", "<i4"), ("N", "<i4"), ("K", "<i4"), ("lda", "<i4"), ("ldb", "<i4"), ("csm", "<i4"), ("csn", "<i4"),
    ("flags", "<i4"), ("splitk", "<i4"), ("head_n", "<i4"), ("head_ld", "<i4"), ("reserved", "<i4"),
    ("rowdot_w", "<u8"), ("rowdot_out", "<u8"), ("rowdot_bias", "<u8"), ("reserved2", "<i8")])
assert GEMM_PROB_DT.itemsize == 128


def order_rec(n: int, ptype: int, colour: int) -> int:
    return int(n) | (int(ptype) << 8) | (int(colour) << 16)


if __name__ == "__main__":
    print(build(force=True, verbose=True))
