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


def test_new_entry_points_report_argument_errors():
    """Wire rows, decode threads, cells: misuse is an error code + message, never a crash or a silent fallback."""
    import ctypes as C

    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv, abi

    L = abi.lib()
    cfg = abi.default_config()
    cfg.num_cells = 5
    h = C.c_void_p()
    assert L.fjsp_create(C.byref(cfg), 64, 0, 0, C.byref(h)) != 0 and b"num_cells" in L.fjsp_last_error()
    cfg.num_cells = 3
    env = BatchedFJSPEnv(200, config=cfg, seed=1)
    env.reset()
    assert L.fjsp_num_cells(env._h) == 3 and env.state_bytes_per_env == 4 * abi.dims(3)["state_words"]
    acts = env.random_actions(0)
    wire = torch.zeros((200, env.dims["wire_words"] + 1), dtype=torch.int32, device=env.device)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t, off=0: C.c_void_p(t.data_ptr() + off)  # noqa: E731
    assert L.fjsp_step_wire(env._h, p(acts), p(wire, 4), None, None, 1, s) != 0 and b"alignment" in L.fjsp_last_error()
    assert L.fjsp_step_wire(env._h, None, p(wire), None, None, 1, s) != 0
    s3 = np.zeros((), dtype=abi.CANON_DT)
    assert L.fjsp_export_state_cell(env._h, 0, 3, C.c_void_p(s3.ctypes.data)) != 0 and b"cell" in L.fjsp_last_error()
    assert L.fjsp_export_state_cell(env._h, 0, 2, C.c_void_p(s3.ctypes.data)) == 0
    assert L.fjsp_set_decode_threads(env._h, 65) != 0
    assert L.fjsp_set_decode_threads(env._h, 2) == 0
    env.step_host(acts.cpu().numpy())
    assert L.fjsp_set_decode_threads(env._h, 4) != 0 and b"before the first" in L.fjsp_last_error()
    assert L.fjsp_step_host_wire(env._h, None, None, 1, s) != 0
    # the env is still usable after the refused calls
    env.step(env.random_actions(1))
    torch.cuda.synchronize()


@pytest.mark.parametrize("cells,n", [(1, 4096), (1, 70000), (4, 3000)])
def test_programmatic_dependent_launches_change_nothing(cells, n, monkeypatch):
    """FJSP_PDL=1 (step / random-action kernels launched as programmatic dependents: the next launch's CTAs become resident
    while the current step drains and wait in griddepcontrol.wait) == FJSP_PDL=0, bit for bit, over back-to-back launches
    with nothing between them — the case where a missing wait would read the previous step's unfinished writes."""
    from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv, abi

    cfg = abi.default_config()
    cfg.num_cells = cells
    envs = []
    for pdl in ("0", "1"):
        monkeypatch.setenv("FJSP_PDL", pdl)   # read when the handle is created
        e = BatchedFJSPEnv(n, config=cfg, seed=21, num_orders=28, autoreset=True)
        e.reset()
        envs.append(e)
    a, b = envs
    outs = []
    for e in (a, b):
        acts = [e.random_actions(t, out=torch.empty((n, e.act_dim), dtype=torch.uint8, device=e.device)) for t in range(48)]
        torch.cuda.synchronize()
        for t in range(48):       # 48 launches in a row on one stream
            e.step(acts[t])
        outs.append((e.obs.clone(), e.rewards.clone(), e.masks.clone(), e.flags.clone(), e.save_state()))
    for x, y in zip(*outs):
        assert torch.equal(x, y)
