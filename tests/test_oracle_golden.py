"""The C restatement (oracle/fjsp_oracle.c) against the golden trajectories recorded from the
unmodified reference (oracle/gen_golden.py).  Bit-exact: observations, masks, float64 rewards,
flags and the canonical integer state at every step."""
import pytest

from oracle.fjsp_oracle import OracleEnv
from tests.util import GOLDEN_FILES, replay_golden


def test_goldens_present():
    assert "config1_uniform" in GOLDEN_FILES and len(GOLDEN_FILES) >= 9


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_oracle_replays_golden(name):
    n = replay_golden(name, lambda cfg: OracleEnv(cfg), exact_rewards=True)
    assert n > 0
