"""The packed-state step function (csrc/fjsp_core.h — the code the CUDA kernels run per env), built
for the host by tests/host_harness, against the same golden trajectories.  Integer state, observations
and masks bit-exact; rewards are fp32 on this path (tolerance 1e-6 relative, in tests/util.py)."""
import pytest

from tests.host_harness.hostharness import HostEnv
from tests.util import GOLDEN_FILES, replay_golden


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_hostcore_replays_golden(name):
    n = replay_golden(name, lambda cfg: HostEnv(cfg))
    assert n > 0
