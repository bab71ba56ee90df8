"""TEST-ONLY N-env stand-in for BatchedFJSPEnv on CPU tensors, backed by the host build of the device step function
(tests/host_harness).  Used to exercise the batched A2C trainer's host/torch logic without a GPU."""
import numpy as np
import torch

from oracle.fjsp_oracle import philox_orders
from tests.host_harness.hostharness import HostEnv


class FakeTensorEnv:
    def __init__(self, num_envs, first_env=0, seed=0, num_orders=30):
        self.num_envs, self.first_env, self.seed, self.num_orders = num_envs, first_env, seed, num_orders
        self.device = torch.device("cpu")
        self.envs = [HostEnv() for _ in range(num_envs)]
        self.episode = [0] * num_envs

    def reset(self):
        obs, masks = torch.zeros(self.num_envs, 38), torch.zeros(self.num_envs, 32, dtype=torch.int8)
        for i, e in enumerate(self.envs):
            o, m = e.reset(philox_orders(self.seed, self.first_env + i, 0, self.num_orders))
            obs[i], masks[i] = torch.from_numpy(o), torch.from_numpy(m)
        return obs, masks

    def step_into(self, actions, obs, masks, rewards, flags):
        a = actions.numpy()
        for i, e in enumerate(self.envs):
            o, m, r, f = e.step(a[i])
            if f[0] or f[1] or f[2]:
                self.episode[i] += 1
                o, m = e.reset(philox_orders(self.seed, self.first_env + i, self.episode[i], self.num_orders))
                f = f.copy()
                f[3] = 1
            obs[i], masks[i], rewards[i], flags[i] = (torch.from_numpy(x) for x in (o, m, r, f))
