"""TEST-ONLY stand-in for BatchedFJSPEnv(num_envs=1) backed by the host build of the device step function
(tests/host_harness).  It lets the CPU-only suite run the facade's HOST logic (dict building, dtype conventions,
order drawing, infos, SimulationView) and the reference's own train.py / a2c.py against it in the build container,
where there is no GPU.  It is monkeypatched in by tests; the product never imports it."""
import ctypes as C

import numpy as np
import torch

from oracle.fjsp_oracle import FjspConfig as OracleCfg
from tests.host_harness.hostharness import HostEnv


class FakeBatchedFJSPEnv:
    def __init__(self, num_envs, config=None, device="cpu", first_env=0, seed=0, num_orders=30, autoreset=True,
                 with_infos=False):
        assert num_envs == 1
        ocfg = OracleCfg()
        C.memmove(C.addressof(ocfg), C.addressof(config), C.sizeof(ocfg))
        self._e = HostEnv(ocfg)
        self.num_envs, self.num_orders, self.autoreset = 1, num_orders, autoreset
        self.device = torch.device("cpu")
        k = max(1, int(config.num_cells))
        agents = 1 + 7 * k
        act, nobs, nmask = (agents + 7) // 8 * 8, 7 + 31 * k, (3 + 26 * k + 31) // 32 * 32
        self.cells = k
        self.long_streams = bool(config.long_streams)
        self.obs = torch.zeros((1, nobs)); self.masks = torch.zeros((1, nmask), dtype=torch.int8)
        self.rewards = torch.zeros((1, act)); self.flags = torch.zeros((1, 4), dtype=torch.uint8)
        self.results = torch.zeros((1, act), dtype=torch.uint8); self.infos = torch.zeros((1, 4), dtype=torch.int32)

    def reset(self, seed=None, num_orders=None, orders=None, env_mask=None):
        tab = np.asarray(orders, dtype=np.uint32)[0]
        o, m = self._e.reset(tab[:self.num_orders])
        self.obs[0] = torch.from_numpy(o); self.masks[0] = torch.from_numpy(m)
        return self.obs, self.masks

    def step(self, actions):
        o, m, r, f = self._e.step(actions.numpy()[0])
        self.obs[0] = torch.from_numpy(o); self.masks[0] = torch.from_numpy(m); self.rewards[0] = torch.from_numpy(r)
        self.flags[0] = torch.from_numpy(f); self.results[0] = torch.from_numpy(self._e.results.copy())
        self.infos[0] = torch.from_numpy(self._e.infos.copy())
        return self.obs, self.rewards, self.flags[:, 0], self.flags[:, 1], self.masks

    def export_state(self, env, cell=0):
        return self._e.export(cell)

    def export_orders(self, env, first, count):
        return self._e.export_orders(first, count)

    def close(self):
        pass
