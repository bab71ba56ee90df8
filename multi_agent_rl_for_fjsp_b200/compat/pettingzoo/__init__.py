"""Stand-in for `pettingzoo` (absent from this image).  The reference only subclasses
`pettingzoo.ParallelEnv` (`/root/reference/FJSPParallelEnvWrapper.py:6,9`) and its callers read
`env.unwrapped` (`a2c.py:298,353,575`)."""


class ParallelEnv:
    metadata = {}
    possible_agents = []
    agents = []

    @property
    def unwrapped(self):
        return self

    @property
    def num_agents(self):
        return len(self.agents)

    @property
    def max_num_agents(self):
        return len(self.possible_agents)

    def close(self):
        pass
