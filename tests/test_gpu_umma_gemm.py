"""tcgen05 grouped GEMM (fjsp_a2c_gemm, csrc/fjsp_umma.cuh) against plain PyTorch references on a B200.

The 3xTF32 path must agree with the fp32 reference of the same op (torch.matmul in fp32, itself checked against fp64)
to rtol 1e-5 of the row scale; the single-pass TF32 path to TF32 accuracy.  Shapes are the trainer's: the reference's
actor / critic layers (networks.py:22-61) forward, backward through a layer, and split-K weight gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda", 0)


def _close(got, ref64, scale, rtol):
    err = (got.double() - ref64).abs().max().item()
    assert err <= rtol * scale, "max abs err %.3e > %.1e * %.3e" % (err, rtol, scale)


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (4096, 256, 256), (1000, 128, 256), (130, 3, 256), (257, 8, 256), (4096, 256, 40)])
@pytest.mark.parametrize("passes", [3, 1])
def test_forward_layer_bias_relu(M, N, K, passes):
    from multi_agent_rl_for_fjsp_b200 import umma

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(M * 7 + N)
    x = torch.randn(M, K, device=dev, generator=g)
    w = torch.randn(K, N, device=dev, generator=g) / K ** 0.5
    b = torch.randn(N, device=dev, generator=g)
    y = torch.full((M, N), float("nan"), device=dev)
    umma.GemmTable(dev, umma.OP_KC, umma.OP_MC, passes).add(x, w, y, M, N, K, lda=K, ldb=N, csm=N, bias=b, relu=True).launch()
    ref64 = torch.relu(x.double() @ w.double() + b.double())
    scale = (x.double().abs() @ w.double().abs()).max().item() + 1.0
    _close(y, ref64, scale, 1e-5 if passes == 3 else 3e-3)
    if passes == 3:  # as good as the fp32 library GEMM
        ref32 = torch.relu(torch.addmm(b, x, w))
        assert torch.allclose(y, ref32, rtol=1e-5, atol=1e-5 * scale)


def test_grouped_forward_from_observation_slices():
    """Layer 1 of the 8 actors + critic in ONE launch: unaligned slices of the [B, 38] observation rows (scalar loads)."""
    from multi_agent_rl_for_fjsp_b200 import umma

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(5)
    B = 777
    obs = torch.randn(B, 38, device=dev, generator=g) * 3
    slices = [(0, 7), (7, 20)] + [(20 + 3 * i, 23 + 3 * i) for i in range(6)] + [(0, 38)]
    ws = [torch.randn(hi - lo, 256, device=dev, generator=g) for lo, hi in slices]
    bs = [torch.randn(256, device=dev, generator=g) for _ in slices]
    out = torch.zeros(len(slices), B, 256, device=dev)
    t = umma.GemmTable(dev, umma.OP_KCS, umma.OP_MC)
    for i, (lo, hi) in enumerate(slices):
        t.add(obs, ws[i], out, B, 256, hi - lo, lda=38, ldb=256, csm=256, a_off=lo, c_off=i * B * 256, bias=bs[i], relu=True)
    t.launch()
    for i, (lo, hi) in enumerate(slices):
        ref = torch.relu(obs[:, lo:hi].double() @ ws[i].double() + bs[i].double())
        _close(out[i], ref, (obs[:, lo:hi].double().abs() @ ws[i].double().abs()).max().item() + 1, 1e-5)


def test_backward_through_a_layer_mask_and_bias_gradient():
    """dx = (dy W^T) * (h > 0), colsum = sum_rows(dx): both operands K-contiguous."""
    from multi_agent_rl_for_fjsp_b200 import umma

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(9)
    B, I, O = 1500, 256, 256
    dy = torch.randn(B, O, device=dev, generator=g)
    w = torch.randn(I, O, device=dev, generator=g) / 16
    h = torch.randn(B, I, device=dev, generator=g)
    dx = torch.full((B, I), float("nan"), device=dev)
    cs = torch.zeros(I, device=dev)
    umma.GemmTable(dev, umma.OP_KC, umma.OP_KC).add(dy, w, dx, B, I, O, lda=O, ldb=O, csm=I, mask=h, colsum=cs).launch()
    ref = (dy.double() @ w.double().t()) * (h > 0)
    scale = (dy.double().abs() @ w.double().abs().t()).max().item()
    _close(dx, ref, scale, 1e-5)
    _close(cs, ref.sum(0), ref.abs().sum(0).max().item(), 1e-5)


@pytest.mark.parametrize("B,splitk", [(4096, 1), (131072, 37), (5000, 8)])
def test_split_k_weight_gradient(B, splitk):
    """dW = x^T dy accumulated with atomics over split-K parts; a second problem writes its result transposed."""
    from multi_agent_rl_for_fjsp_b200 import umma

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(B)
    x = torch.randn(B, 256, device=dev, generator=g)
    dy = torch.randn(B, 256, device=dev, generator=g)
    obs = torch.randn(B, 38, device=dev, generator=g)
    dw = torch.zeros(256, 256, device=dev)
    dw1 = torch.zeros(13, 256, device=dev)  # layer-1 gradient of the AGV actor: dW1[i, j] = sum_b obs[b, 7 + i] * dy[b, j]
    t = umma.GemmTable(dev, umma.OP_MC, umma.OP_MC)
    t.add(x, dy, dw, 256, 256, B, lda=256, ldb=256, csm=256, atomic=True, splitk=splitk)
    t.add(dy, obs, dw1, 256, 13, B, lda=256, ldb=38, csm=1, csn=256, b_off=7, atomic=True, splitk=splitk)
    t.launch()
    ref = x.double().t() @ dy.double()
    _close(dw, ref, (x.double().abs().t() @ dy.double().abs()).max().item(), 1e-5)
    ref1 = obs[:, 7:20].double().t() @ dy.double()
    _close(dw1, ref1, (obs[:, 7:20].double().abs().t() @ dy.double().abs()).max().item(), 1e-5)


def test_sass_has_tcgen05():
    """The library's GEMM is tensor-core code of this generation: UTC*MMA (tcgen05.mma) and LDTM (tcgen05.ld) in the SASS."""
    import shutil
    import subprocess

    from multi_agent_rl_for_fjsp_b200 import abi

    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([tool, "-sass", abi.SO_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass
