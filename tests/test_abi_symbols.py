"""The C-ABI library loads on a CPU-only box and exports every symbol include/fjsp_b200.h declares
(no compute calls without a GPU)."""
import ctypes as C
import os
import re

import pytest

from multi_agent_rl_for_fjsp_b200 import abi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    abi.build()
    return abi.lib()


def declared_functions():
    src = open(os.path.join(REPO, "include", "fjsp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fjsp_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    decl = declared_functions()
    assert sorted(abi.EXPORTS) == decl, (sorted(set(decl) ^ set(abi.EXPORTS)))
    for name in decl:
        assert hasattr(lib, name), name


def test_struct_sizes_match_header(lib, tmp_path):
    import subprocess

    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu\\n", sizeof(FjspConfig), sizeof(FjspCanonState));return 0;}\n'
                 % os.path.join(REPO, "include", "fjsp_b200.h"))
    exe = tmp_path / "sz"
    subprocess.run(["gcc", str(c), "-o", str(exe)], check=True)
    cfg_sz, canon_sz = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert cfg_sz == C.sizeof(abi.FjspConfig) and canon_sz == abi.CANON_DT.itemsize


def test_default_config_and_errors_without_gpu(lib):
    cfg = abi.default_config()  # host-only call
    assert (cfg.proc_small, cfg.proc_big, cfg.proc_pack, cfg.step_size, cfg.max_episode_steps) == (60, 120, 30, 10, 200)
    assert [tuple(cfg.pos[i]) for i in range(5)] == [(0, 0), (0, 3), (2, 3), (3, 0), (3, 5)]
    assert lib.fjsp_abi_version() == abi.ABI_VERSION == 4
    import torch

    if not torch.cuda.is_available():
        h = C.c_void_p()
        rc = lib.fjsp_create(C.byref(cfg), 64, 0, 0, C.byref(h))
        assert rc != 0 and not h.value  # fails loudly: no CPU path
        assert lib.fjsp_last_error()
        from multi_agent_rl_for_fjsp_b200 import BatchedFJSPEnv

        with pytest.raises(RuntimeError):
            BatchedFJSPEnv(4)


def test_built_for_sm100a_with_bulk_copies():
    import shutil
    import subprocess

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-sass", abi.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out, "step kernel must move its tiles with bulk async copies (TMA engine)"
    assert "UBLKPF" in out, "step kernels prefetch a later tile into L2 with a bulk prefetch"
    for k in (1, 2, 3, 4):  # every cell count has its own instantiations: thread-per-env, and cell-parallel for K >= 2
        assert "fjsp_step_kernelILi%dE" % k in out.replace("16fjsp_step_kernel", "fjsp_step_kernel")
    assert "fjsp_step_cells_kernelILi4ELb1" in out and "fjsp_rollout_cells_kernelILi2E" in out
