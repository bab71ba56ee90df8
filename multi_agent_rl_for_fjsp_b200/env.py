"""BatchedFJSPEnv — tensor API over N independent shop floors stepped in lockstep on one B200.

Host-side mirror of the reference's env surface for the batched case: the same reset/step meaning as
``FJSPParallelEnv.reset/step`` (/root/reference/FJSPParallelEnvWrapper.py:43-69), with the 8 agents'
dicts replaced by dense tensors in the canonical agent order (FJSPSimulation.py:76-82):

    actions  uint8 [N, 8]      obs  float32 [N, 38] (a2c._flatten_obs order)      masks int8 [N, 32] (29 used)
    rewards  float32 [N, 8]    flags uint8 [N, 4] = terminated, truncated, fault, was_reset

A scaled shop (``num_cells`` = K in 2..4, include/fjsp_b200.h) has 1 + 7K agents: the row widths become
``abi.dims(K)`` (K = 4: actions/rewards 32, obs 131, masks 128) and ``agent_ids`` / ``n_actions`` / ``obs_slices`` /
``mask_offsets`` describe the columns.

PyTorch is plumbing here (device memory + the current stream); every simulation step is one launch of
the sm_100a kernel behind ``fjsp_step`` in libfjsp_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import abi

AGENT_IDS = [
    "pickup_station", "agv", "small_machine", "big_machine",
    "packaging_blue_1", "packaging_blue_2", "packaging_red", "packaging_green",
]
N_ACTIONS = (3, 8, 3, 3, 3, 3, 3, 3)
OBS_DIM, MASK_DIM = abi.OBS_DIM, abi.MASK_DIM
OBS_SLICES = [(0, 7), (7, 20), (20, 23), (23, 26), (26, 29), (29, 32), (32, 35), (35, 38)]
MASK_OFFSETS = [0, 3, 11, 14, 17, 20, 23, 26, 29]


_CELL_AGENTS = AGENT_IDS[1:]


def agent_layout(cells: int = 1):
    """(agent_ids, n_actions, obs_slices, mask_offsets) of a K-cell shop; K = 1 gives the module constants."""
    if cells == 1:
        return list(AGENT_IDS), tuple(N_ACTIONS), list(OBS_SLICES), list(MASK_OFFSETS)
    ids, nact, obs_sl, mask_off = ["pickup_station"], [3], [(0, 7)], [0, 3]
    for c in range(cells):
        for j, name in enumerate(_CELL_AGENTS):
            ids.append("%s_c%d" % (name, c))
            nact.append(N_ACTIONS[1 + j])
            lo, hi = OBS_SLICES[1 + j]
            obs_sl.append((lo + 31 * c, hi + 31 * c))
            mask_off.append(mask_off[-1] + N_ACTIONS[1 + j])
    return ids, tuple(nact), obs_sl, mask_off


def shared_agent_layout(agvs: int):
    """(agent_ids, n_actions, obs_slices, mask_offsets) of the shared floor (include/fjsp_b200.h): pickup station, agv_0 ..
    agv_{A-1}, machines, packaging stations."""
    ids, nact, obs_sl, mask_off = ["pickup_station"], [3], [(0, 7)], [0, 3]
    for j in range(agvs):
        ids.append("agv_%d" % j)
        nact.append(8)
        obs_sl.append((7 + 13 * j, 20 + 13 * j))
        mask_off.append(mask_off[-1] + 8)
    base = 7 + 13 * agvs
    for j, name in enumerate(AGENT_IDS[2:]):
        ids.append(name)
        nact.append(3)
        obs_sl.append((base + 3 * j, base + 3 * j + 3))
        mask_off.append(mask_off[-1] + 3)
    return ids, tuple(nact), obs_sl, mask_off


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedFJSPEnv:
    def __init__(self, num_envs: int, config=None, device="cuda:0", first_env: int = 0, seed: int = 0,
                 num_orders: int = 30, autoreset: bool = True, with_infos: bool = False, decode_threads: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedFJSPEnv needs a CUDA device: the environment step exists only as sm_100a kernels")
        self._L = abi.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchedFJSPEnv runs on CUDA devices only (got %s)" % device)
        self.cfg = config if isinstance(config, abi.FjspConfig) else abi.config_from_dict(config)
        self.num_envs, self.first_env = int(num_envs), int(first_env)
        self.cells = int(self.cfg.num_cells)
        if not 1 <= self.cells <= abi.MAX_CELLS:
            raise ValueError("num_cells must be in 1..%d" % abi.MAX_CELLS)
        self.long_streams = bool(self.cfg.long_streams)   # the long order-stream layout (include/fjsp_b200.h)
        self.max_orders = abi.LONG_MAX_ORDERS if self.long_streams else abi.MAX_ORDERS
        self.shared_agvs = int(self.cfg.shared_agvs) if int(self.cfg.shared_agvs) >= 2 else 0   # shared floor (include/fjsp_b200.h)
        self.dims = abi.dims(self.cells, self.long_streams, self.shared_agvs)
        self.agent_ids, self.n_actions, self.obs_slices, self.mask_offsets = (
            shared_agent_layout(self.shared_agvs) if self.shared_agvs else agent_layout(self.cells))
        self.act_dim, self.obs_dim, self.mask_dim = self.dims["act"], self.dims["obs"], self.dims["mask"]
        self.seed, self.num_orders, self.autoreset = int(seed), int(num_orders), bool(autoreset)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        torch.cuda.init()
        with torch.cuda.device(dev_index):
            torch.zeros(1, device=self.device)  # make sure the primary context exists before the library uses it
            h = C.c_void_p()
            abi.check(self._L.fjsp_create(C.byref(self.cfg), self.num_envs, self.first_env, dev_index, C.byref(h)))
        self._h = h
        if decode_threads:  # host threads of the step_host decode (0 = all CPUs of the process)
            abi.check(self._L.fjsp_set_decode_threads(h, int(decode_threads)))
        n = self.num_envs
        kw = dict(device=self.device)
        self.obs = torch.zeros((n, self.obs_dim), dtype=torch.float32, **kw)
        self.masks = torch.zeros((n, self.mask_dim), dtype=torch.int8, **kw)
        self.rewards = torch.zeros((n, self.act_dim), dtype=torch.float32, **kw)
        self.flags = torch.zeros((n, 4), dtype=torch.uint8, **kw)
        self.results = torch.zeros((n, self.act_dim), dtype=torch.uint8, **kw) if with_infos else None
        self.infos = torch.zeros((n, 4), dtype=torch.int32, **kw) if with_infos else None
        self._term, self._trunc = self.flags[:, 0], self.flags[:, 1]
        self._actions = torch.zeros((n, self.act_dim), dtype=torch.uint8, **kw)
        self._stats = torch.zeros(8, dtype=torch.int64, **kw)
        self._t = 0
        self._host = None
        # constant argument block of fjsp_step for the env's own output tensors (built once: the eager call path is
        # launch-latency bound at small batch sizes, so per-call Python work matters)
        self._out_ptrs = (_ptr(self.obs), _ptr(self.masks), _ptr(self.rewards), _ptr(self.flags), _ptr(self.results),
                          _ptr(self.infos))
        self._fjsp_step = self._L.fjsp_step

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.fjsp_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launch_count(self) -> int:
        return int(self._L.fjsp_launch_count(self._h))

    @property
    def state_bytes_per_env(self) -> int:
        return int(self._L.fjsp_state_bytes(self._h))

    # ------------------------------------------------------------------ reset / step
    def reset(self, seed: int | None = None, num_orders: int | None = None, orders=None, env_mask=None):
        """FJSPParallelEnv.reset for all (or the masked) envs.

        orders: optional explicit order tables, uint32 tensor/array [N, 32] packed n | type<<8 | colour<<16
        (``abi.order_rec``; long order streams: [N, num_orders]) or int array [N, num_orders, 3]; default: Philox stream
        (seed, global env, episode 0).
        """
        if seed is not None:
            self.seed = int(seed)
        if num_orders is not None:
            self.num_orders = int(num_orders)
        d_orders = None
        if orders is not None:
            if self.autoreset:
                raise ValueError("explicit order tables describe ONE episode; with autoreset=True every later episode would "
                                 "silently draw Philox orders instead — create the env with autoreset=False to use them")
            d_orders = self._pack_orders(orders, env_mask)
        d_mask = None
        if env_mask is not None:
            d_mask = torch.as_tensor(env_mask, device=self.device).to(torch.uint8).contiguous()
            assert d_mask.numel() == self.num_envs
        # (a masked reset leaves the handle-wide seed / num_orders — used by auto-reset for every env — untouched)
        abi.check(self._L.fjsp_reset(self._h, _ptr(d_mask), self.seed, _ptr(d_orders), self.num_orders, _ptr(self.obs),
                                     _ptr(self.masks), self._stream()))
        self._keep = (d_orders, d_mask)  # keep the buffers alive until the stream has consumed them
        return self.obs, self.masks

    def _pack_orders(self, orders, env_mask=None):
        arr = orders.detach().cpu().numpy() if isinstance(orders, torch.Tensor) else np.asarray(orders)
        if arr.ndim == 3:  # [N, k, 3] -> packed [N, 32] (long order streams: [N, k])
            n, k, _ = arr.shape
            packed = np.zeros((n, max(k, 1) if self.long_streams else abi.MAX_ORDERS), dtype=np.uint32)
            packed[:, :k] = (arr[..., 0].astype(np.uint32) | (arr[..., 1].astype(np.uint32) << 8) |
                             (arr[..., 2].astype(np.uint32) << 16))
            if self.num_orders != k:
                self.num_orders = k
            arr = packed
        arr = np.ascontiguousarray(arr, dtype=np.uint32)
        if self.long_streams:
            assert arr.ndim == 2 and arr.shape[0] == self.num_envs and arr.shape[1] >= self.num_orders, arr.shape
            arr = np.ascontiguousarray(arr[:, :max(self.num_orders, 1)])
        else:
            assert arr.shape == (self.num_envs, abi.MAX_ORDERS), arr.shape
        live = arr[:, :self.num_orders]
        if env_mask is not None:  # only the rows of the envs being reset are read
            live = live[np.asarray(env_mask.cpu() if isinstance(env_mask, torch.Tensor) else env_mask).astype(bool).reshape(-1)]
        n, ty, co = live & 0xff, (live >> 8) & 0xff, (live >> 16) & 0xff
        if live.size and not (((n >= 1) & (n <= 9) & (ty >= 1) & (ty <= 3) & (co >= 1) & (co <= 3) & ((live >> 24) == 0)).all()):
            raise ValueError("order records must have n in 1..9, type in 1..3, colour in 1..3 (FJSPSimulation.py:107-112)")
        return torch.from_numpy(arr.view(np.int32)).to(self.device)

    def step(self, actions: torch.Tensor):
        """One lockstep step of all envs: ONE kernel launch.  Returns views of the env's output tensors."""
        a = actions
        if a.dtype != torch.uint8 or a.device != self.device or not a.is_contiguous():
            a = a.to(device=self.device, dtype=torch.uint8).contiguous()
        assert a.shape == (self.num_envs, self.act_dim), a.shape
        o = self._out_ptrs
        rc = self._fjsp_step(self._h, a.data_ptr(), o[0], o[1], o[2], o[3], o[4], o[5], self.autoreset,
                             torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            abi.check(rc)
        self._t += 1
        return self.obs, self.rewards, self._term, self._trunc, self.masks

    def step_into(self, actions: torch.Tensor, obs: torch.Tensor, masks: torch.Tensor, rewards: torch.Tensor,
                  flags: torch.Tensor):
        """Same launch, but the kernel writes straight into caller-owned tensors (e.g. slices of a rollout buffer,
        so the policy's input needs no copy).  Shapes/dtypes as the env's own output tensors; contiguous."""
        n = self.num_envs
        assert actions.dtype == torch.uint8 and actions.is_contiguous() and actions.shape == (n, self.act_dim)
        assert obs.dtype == torch.float32 and obs.is_contiguous() and obs.shape == (n, self.obs_dim)
        assert masks.dtype == torch.int8 and masks.is_contiguous() and masks.shape == (n, self.mask_dim)
        assert rewards.dtype == torch.float32 and rewards.is_contiguous() and rewards.shape == (n, self.act_dim)
        assert flags.dtype == torch.uint8 and flags.is_contiguous() and flags.shape == (n, 4)
        abi.check(self._L.fjsp_step(self._h, _ptr(actions), _ptr(obs), _ptr(masks), _ptr(rewards), _ptr(flags), None, None,
                                    int(self.autoreset), self._stream()))
        self._t += 1

    def step_wire(self, actions: torch.Tensor, wire: torch.Tensor):
        """Same launch with the results written as wire rows (include/fjsp_b200.h): `wire` int32 [N, dims["wire_words"]]
        on the env's device.  ``abi.lib().fjsp_wire_decode`` (host) turns rows into the tensors `step` returns."""
        n = self.num_envs
        assert actions.dtype == torch.uint8 and actions.is_contiguous() and actions.shape == (n, self.act_dim)
        assert wire.dtype == torch.int32 and wire.is_contiguous() and wire.shape == (n, self.dims["wire_words"])
        abi.check(self._L.fjsp_step_wire(self._h, _ptr(actions), _ptr(wire), _ptr(self.results), _ptr(self.infos),
                                         int(self.autoreset), self._stream()))
        self._t += 1

    def random_actions(self, t: int | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
        """a_i ~ U{0..n_i-1} from the Philox action stream at time index t (default: the env's step counter)."""
        out = self._actions if out is None else out
        abi.check(self._L.fjsp_random_actions(self._h, self.seed, self._t if t is None else int(t), _ptr(out), self._stream()))
        return out

    def rollout_random(self, steps: int, t0: int | None = None) -> torch.Tensor:
        """`steps` steps per launch with in-kernel random actions and auto-reset; returns the cumulative stats tensor
        (env_steps, episodes, orders_completed, products_packaged, faults, sum round(40*reward), 0, 0)."""
        t0 = self._t if t0 is None else int(t0)
        abi.check(self._L.fjsp_rollout_random(self._h, int(steps), self.seed, t0, _ptr(self._stats), self._stream()))
        self._t = t0 + int(steps)
        return self._stats

    # ------------------------------------------------------------------ host-buffer path (end-to-end)
    def host_buffers(self):
        """The pinned host tensors `step_host` reads actions from and writes results to (allocated on first use)."""
        if self._host is None:
            n = self.num_envs
            self._host = dict(
                actions=torch.zeros((n, self.act_dim), dtype=torch.uint8).pin_memory(),
                obs=torch.zeros((n, self.obs_dim), dtype=torch.float32).pin_memory(),
                masks=torch.zeros((n, self.mask_dim), dtype=torch.int8).pin_memory(),
                rewards=torch.zeros((n, self.act_dim), dtype=torch.float32).pin_memory(),
                flags=torch.zeros((n, 4), dtype=torch.uint8).pin_memory())
        return self._host

    def step_host(self, actions=None):
        """Same step through HOST buffers: H2D actions, kernel, D2H obs/masks/rewards/flags, synchronised on return.

        actions: None (use what the caller wrote into ``host_buffers()["actions"]``), a pinned uint8 CPU tensor [N,8]
        (read in place, no staging copy) or any array-like (copied into the pinned buffer first).
        Returns NumPy views of the pinned result buffers."""
        hb = self.host_buffers()
        src = hb["actions"]
        if actions is not None:
            if (isinstance(actions, torch.Tensor) and actions.device.type == "cpu" and actions.dtype == torch.uint8
                    and actions.is_contiguous() and actions.is_pinned()
                    and tuple(actions.shape) == (self.num_envs, self.act_dim)):
                src = actions
            else:
                hb["actions"].numpy()[...] = np.asarray(actions, dtype=np.uint8)
        abi.check(self._L.fjsp_step_host(self._h, _ptr(src), _ptr(hb["obs"]), _ptr(hb["masks"]), _ptr(hb["rewards"]),
                                         _ptr(hb["flags"]), int(self.autoreset), self._stream()))
        self._t += 1
        return hb["obs"].numpy(), hb["masks"].numpy(), hb["rewards"].numpy(), hb["flags"].numpy()

    def step_host_wire(self, actions: torch.Tensor):
        """Host-buffer step that delivers the compact wire rows (int32 [N, dims["wire_words"]], pinned) without decoding
        them; `actions`: pinned uint8 CPU tensor [N, act_dim].  ``abi.lib().fjsp_wire_decode`` turns rows into tensors."""
        hb = self.host_buffers()
        if "wire" not in hb:
            hb["wire"] = torch.zeros((self.num_envs, self.dims["wire_words"]), dtype=torch.int32).pin_memory()
        assert actions.device.type == "cpu" and actions.dtype == torch.uint8 and actions.is_contiguous()
        assert tuple(actions.shape) == (self.num_envs, self.act_dim)
        abi.check(self._L.fjsp_step_host_wire(self._h, _ptr(actions), _ptr(hb["wire"]), int(self.autoreset), self._stream()))
        self._t += 1
        return hb["wire"]

    # ------------------------------------------------------------------ snapshot / restore
    def save_state(self) -> torch.Tensor:
        """Copy of the whole packed state (int32 tensor on the env's device)."""
        nbytes = int(self._L.fjsp_state_total_bytes(self._h))
        buf = torch.empty(nbytes // 4, dtype=torch.int32, device=self.device)
        abi.check(self._L.fjsp_state_save(self._h, _ptr(buf), nbytes, self._stream()))
        return buf

    def load_state(self, buf: torch.Tensor):
        nbytes = int(self._L.fjsp_state_total_bytes(self._h))
        assert buf.device == self.device and buf.is_contiguous() and buf.numel() * buf.element_size() == nbytes
        abi.check(self._L.fjsp_state_load(self._h, _ptr(buf), nbytes, self._stream()))

    # ------------------------------------------------------------------ diagnostics
    def export_state(self, env: int, cell: int = 0) -> np.ndarray:
        """Canonical integer record S of one env (synchronises); for a scaled shop: the shared part + the given cell."""
        s = np.zeros((), dtype=abi.CANON_DT)
        abi.check(self._L.fjsp_export_state_cell(self._h, int(env), int(cell), C.c_void_p(s.ctypes.data)))
        return s

    def export_orders(self, env: int, first: int, count: int) -> np.ndarray:
        """[count, 4] int32: packaged_mask, processed_mask, is_complete, completion_step of orders first .. first+count-1
        (synchronises; conventions of the long layout in include/fjsp_b200.h fjsp_export_orders)."""
        out = np.zeros((count, 4), dtype=np.int32)
        abi.check(self._L.fjsp_export_orders(self._h, int(env), int(first), int(count), C.c_void_p(out.ctypes.data), None))
        return out

    def export_packed(self, env: int) -> np.ndarray:
        w = np.zeros(self.dims["state_words"], dtype=np.uint32)
        abi.check(self._L.fjsp_export_packed(self._h, int(env), C.c_void_p(w.ctypes.data)))
        return w


class CellViewEnv:
    """A K-cell ``BatchedFJSPEnv`` seen as N*K rows of the reference's own 8-agent layout (view row = env * K + cell:
    the pickup station + that cell's seven agents), so a trainer written for the reference shop — ``BatchedA2C`` —
    runs on the scaled shop unchanged, its 8 actors and the critic shared by the cells (include/fjsp_b200.h "cell
    views").  The pickup station acts through the row of cell 0; elsewhere its mask allows action 0 only.
    Two small kernels per step around the env's own single launch (pack actions, unpack views); K = 1 is the identity."""

    def __init__(self, env: BatchedFJSPEnv):
        self.env, self.cells = env, env.cells
        self.device, self.num_envs, self.first_env = env.device, env.num_envs * env.cells, env.first_env * env.cells
        self.seed = env.seed
        self._L = env._L
        self._actions = torch.zeros((env.num_envs, env.act_dim), dtype=torch.uint8, device=env.device)
        n = self.num_envs
        self.obs = torch.zeros((n, OBS_DIM), dtype=torch.float32, device=env.device)
        self.masks = torch.zeros((n, MASK_DIM), dtype=torch.int8, device=env.device)
        self.rewards = torch.zeros((n, 8), dtype=torch.float32, device=env.device)
        self.flags = torch.zeros((n, 4), dtype=torch.uint8, device=env.device)

    @property
    def launch_count(self):
        return self.env.launch_count

    def _unpack(self, obs, masks, rewards, flags):
        e = self.env
        abi.check(self._L.fjsp_cells_unpack_views(_ptr(e.obs), _ptr(e.masks), _ptr(e.rewards), _ptr(e.flags), _ptr(obs), _ptr(masks),
                                                  _ptr(rewards), _ptr(flags), e.num_envs, self.cells, e._stream()))

    def reset(self, **kw):
        self.env.reset(**kw)
        self._unpack(self.obs, self.masks, self.rewards, self.flags)
        return self.obs, self.masks

    def step_into(self, actions, obs, masks, rewards, flags):
        n, e = self.num_envs, self.env
        assert actions.dtype == torch.uint8 and actions.is_contiguous() and actions.shape == (n, 8)
        assert obs.is_contiguous() and obs.shape == (n, OBS_DIM) and masks.is_contiguous() and masks.shape == (n, MASK_DIM)
        assert rewards.is_contiguous() and rewards.shape == (n, 8) and flags.is_contiguous() and flags.shape == (n, 4)
        abi.check(self._L.fjsp_cells_pack_actions(_ptr(actions), _ptr(self._actions), e.num_envs, self.cells, e._stream()))
        e.step(self._actions)
        self._unpack(obs, masks, rewards, flags)

    def step(self, actions):
        self.step_into(actions, self.obs, self.masks, self.rewards, self.flags)
        return self.obs, self.rewards, self.flags[:, 0], self.flags[:, 1], self.masks

    def save_state(self):
        return self.env.save_state()

    def load_state(self, buf):
        self.env.load_state(buf)
