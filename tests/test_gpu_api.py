"""Tensor-API behaviours of BatchedFJSPEnv on the device: snapshot/restore, masked reset, caller-owned output
tensors, argument checking, two handles on one device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_snapshot_restore_is_deterministic():
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    env = BatchedFJSPEnv(1000, seed=8, autoreset=True)
    env.reset()
    for t in range(30):
        env.step(env.random_actions(t))
    snap = env.save_state()
    outs = []
    for rep in range(2):
        env.load_state(snap)
        for t in range(30, 260):  # crosses an auto-reset
            env.step(env.random_actions(t))
        outs.append((env.obs.clone(), env.rewards.clone(), env.flags.clone(), env.save_state()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_masked_reset_touches_only_selected_envs():
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    n = 300
    env = BatchedFJSPEnv(n, seed=1, autoreset=False)
    env.reset()
    for t in range(20):
        env.step(env.random_actions(t))
    before = [env.export_packed(i) for i in (0, 1, 2, 299)]
    mask = np.zeros(n, np.uint8)
    mask[[1, 299]] = 1
    obs, masks = env.reset(env_mask=mask)
    after = [env.export_packed(i) for i in (0, 1, 2, 299)]
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[2], after[2])
    assert not np.array_equal(before[1], after[1]) and not np.array_equal(before[3], after[3])
    assert env.export_state(1)["current_step"] == 0 and env.export_state(0)["current_step"] == 20
    assert obs[1, 11:13].tolist() == [0.0, 0.0]


def test_step_into_writes_caller_tensors_and_matches_step():
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    a = BatchedFJSPEnv(513, seed=4)
    b = BatchedFJSPEnv(513, seed=4)
    a.reset(), b.reset()
    dev = a.device
    buf_obs = torch.zeros(3, 513, 38, device=dev)
    buf_m = torch.zeros(3, 513, 32, dtype=torch.int8, device=dev)
    buf_r = torch.zeros(3, 513, 8, device=dev)
    buf_f = torch.zeros(3, 513, 4, dtype=torch.uint8, device=dev)
    for t in range(3):
        acts = a.random_actions(t).clone()
        a.step(acts)
        b.step_into(acts, buf_obs[t], buf_m[t], buf_r[t], buf_f[t])
        assert torch.equal(a.obs, buf_obs[t]) and torch.equal(a.masks, buf_m[t])
        assert torch.equal(a.rewards, buf_r[t]) and torch.equal(a.flags, buf_f[t])


def test_argument_errors_are_reported_not_fatal():
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv, abi

    with pytest.raises(RuntimeError, match="num_envs"):
        BatchedFJSPEnv(0)
    cfg = abi.default_config()
    cfg.max_episode_steps = 1000
    with pytest.raises(RuntimeError, match="max_episode_steps"):
        BatchedFJSPEnv(8, config=cfg)
    env = BatchedFJSPEnv(8)
    with pytest.raises(RuntimeError, match="num_orders"):
        env.reset(num_orders=33)
    with pytest.raises(AssertionError):
        env.step(torch.zeros(7, 8, dtype=torch.uint8, device=env.device))
    env.reset(num_orders=30)
    env.step(torch.zeros(8, 8, dtype=torch.uint8, device=env.device))  # still usable after the errors


def test_out_of_range_actions_follow_reference_rules():
    """AGV action > 7 is an invalid action (-5); other agents' out-of-range actions silently do nothing."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

    env = BatchedFJSPEnv(2, autoreset=False)
    env.reset()
    acts = torch.tensor([[9, 200, 7, 5, 3, 3, 3, 3], [0, 0, 0, 0, 0, 0, 0, 0]], dtype=torch.uint8, device=env.device)
    obs, rew, term, trunc, masks = env.step(acts)
    r = rew.cpu().numpy()
    assert r[0, 1] == pytest.approx(-5.125) and r[0, 0] == pytest.approx(-0.125)   # pickup did nothing, not even idle
    assert r[1, 0] == pytest.approx(-1.125) and r[1, 1] == pytest.approx(-0.125)
    assert np.allclose(r[0, 2:], -0.125)
