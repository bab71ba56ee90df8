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


@pytest.mark.parametrize("M,N,K,src_op", [(4096, 256, 256, "MC"), (1000, 128, 256, "MC"), (130, 3, 256, "MC"), (257, 8, 40, "MC"),
                                            (1500, 256, 256, "KC"), (900, 256, 3, "KC"), (513, 128, 1, "KC"), (4096, 256, 38, "MC")])
@pytest.mark.parametrize("passes", [3, 1])
def test_packed_weight_images_give_the_same_bits(M, N, K, src_op, passes):
    """FJSP_OP_PK: B read from the packed (hi, lo) image of fjsp_a2c_gemm_pack == B converted on the fly, bit for bit
    (same split, same MMA order), for weights stored [K][N] (forward) and [N][K] (dx = dy W^T), ragged N and K."""
    from multi_agent_rl_for_fjsp_b200 import umma

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(M + 3 * N + K)
    w = torch.randn(K, N, device=dev, generator=g) if src_op == "MC" else torch.randn(N, K, device=dev, generator=g)
    b = torch.randn(N, device=dev, generator=g)
    if K % 4 == 0:      # 16-byte loads of A
        x, lda, a_off, a_op = torch.randn(M, K, device=dev, generator=g), K, 0, umma.OP_KC
    else:               # unaligned rows: scalar loads of A
        x, lda, a_off, a_op = torch.randn(M, K + 2, device=dev, generator=g), K + 2, 1, umma.OP_KCS
    y0 = torch.full((M, N), float("nan"), device=dev)
    y1 = torch.full((M, N), float("nan"), device=dev)
    b_plain = umma.OP_MC if src_op == "MC" else a_op   # the unpacked orientation pairs the library has: (.,MC), (KC,KC), (KCS,KCS)
    umma.GemmTable(dev, a_op, b_plain, passes).add(x, w, y0, M, N, K, lda=lda, ldb=(N if src_op == "MC" else K), csm=N, a_off=a_off,
                                                   bias=b, relu=True).launch()
    pk = umma.PackTable(dev)
    off = pk.add(w, umma.OP_MC if src_op == "MC" else umma.OP_KCS, N if src_op == "MC" else K, N, K)
    pk.finalize().launch()
    umma.GemmTable(dev, a_op, umma.OP_PK, passes).add(x, pk.image, y1, M, N, K, lda=lda, ldb=0, csm=N, a_off=a_off, b_off=off, bias=b,
                                                      relu=True).launch()
    assert torch.equal(y0, y1)
    xv = x[:, a_off:a_off + K].double()
    ref64 = torch.relu(xv @ (w.double() if src_op == "MC" else w.double().t()) + b.double())
    scale = (xv.abs() @ (w.double().abs() if src_op == "MC" else w.double().abs().t())).max().item() + 1.0
    _close(y1, ref64, scale, 1e-5 if passes == 3 else 3e-3)


def test_packed_images_follow_the_weights():
    """Re-packing after the weights changed (an optimizer step) is one launch over the job table."""
    from multi_agent_rl_for_fjsp_b200 import umma

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(1)
    ws = [torch.randn(256, 256, device=dev, generator=g), torch.randn(256, 8, device=dev, generator=g)]
    x = torch.randn(640, 256, device=dev, generator=g)
    pk = umma.PackTable(dev)
    offs = [pk.add(ws[0], umma.OP_MC, 256, 256, 256), pk.add(ws[1], umma.OP_MC, 8, 8, 256)]
    pk.finalize().launch()
    ys = [torch.zeros(640, 256, device=dev), torch.zeros(640, 8, device=dev)]
    t = umma.GemmTable(dev, umma.OP_KC, umma.OP_PK)
    t.add(x, pk.image, ys[0], 640, 256, 256, lda=256, ldb=0, csm=256, b_off=offs[0])
    t.add(x, pk.image, ys[1], 640, 8, 256, lda=256, ldb=0, csm=8, b_off=offs[1])
    t.launch()
    for w, y in zip(ws, ys):
        _close(y, x.double() @ w.double(), (x.double().abs() @ w.double().abs()).max().item(), 1e-5)
    for w in ws:
        w.mul_(-0.5)
    pk.launch()
    t.launch()
    for w, y in zip(ws, ys):
        _close(y, x.double() @ w.double(), (x.double().abs() @ w.double().abs()).max().item(), 1e-5)


@pytest.mark.parametrize("B,nx,ny,transposed", [(5000, 256, 3, False), (1537, 256, 8, False), (4096, 256, 13, True), (3000, 256, 38, True),
                                                 (2049, 128, 1, False), (700, 256, 7, True), (513, 200, 16, False)])
def test_narrow_weight_gradients(B, nx, ny, transposed):
    """fjsp_a2c_wgrad_small: G[i, j] += sum_b X[b, i] * Y[b, j] (slices of wider rows, either storage order of G, accumulation
    into what G holds) against float64; plain fp32 FMAs: rtol 1e-5 of the absolute-value product."""
    from multi_agent_rl_for_fjsp_b200 import umma

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(B + nx + ny)
    xw, yw = torch.randn(B, nx + 5, device=dev, generator=g), torch.randn(B, ny + 9, device=dev, generator=g)
    x_off, y_off = 2, 4
    G = torch.randn(ny, nx, device=dev, generator=g) if transposed else torch.randn(nx, ny, device=dev, generator=g)
    G0 = G.clone()
    t = umma.WgradTable(dev)
    t.add(xw, yw, G, B, nx, ny, ldx=nx + 5, ldy=ny + 9, gsi=(1 if transposed else ny), gsj=(nx if transposed else 1), x_off=x_off, y_off=y_off)
    t.launch()
    X, Y = xw[:, x_off:x_off + nx].double(), yw[:, y_off:y_off + ny].double()
    ref = X.t() @ Y
    ref = G0.double() + (ref.t() if transposed else ref)
    scale = (X.abs().t() @ Y.abs()).max().item()
    _close(G, ref, scale, 1e-5)


def test_narrow_weight_gradients_grouped():
    """Jobs of different widths in one table (three launches at most), different batch sizes."""
    from multi_agent_rl_for_fjsp_b200 import umma

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(4)
    shapes = [(3000, 256, 3), (3000, 256, 8), (1000, 128, 1), (3000, 256, 13), (2500, 256, 38), (3000, 256, 7)]
    xs = [torch.randn(b, nx, device=dev, generator=g) for b, nx, ny in shapes]
    ys = [torch.randn(b, ny, device=dev, generator=g) for b, nx, ny in shapes]
    gs = [torch.zeros(nx, ny, device=dev) for b, nx, ny in shapes]
    t = umma.WgradTable(dev)
    for (b, nx, ny), x, y, G in zip(shapes, xs, ys, gs):
        t.add(x, y, G, b, nx, ny, ldx=nx, ldy=ny, gsi=ny, gsj=1)
    t.launch()
    for x, y, G in zip(xs, ys, gs):
        _close(G, x.double().t() @ y.double(), (x.double().abs().t() @ y.double().abs()).max().item(), 1e-5)


@pytest.mark.parametrize("M,N,R", [(4096, 256, 8), (1000, 256, 3), (130, 100, 1), (257, 136, 5)])
def test_fused_head_of_up_to_eight_columns(M, N, R):
    """``head_n`` = R: out[m, j] = sum_n relu(x W + b)[m, n] * Wh[n, j] + bh[j] in the epilogue of the layer (an actor's
    256 -> 3..8 logits layer, networks.py:22-38), written into columns [off, off + R) of wider rows; two problems in one
    launch, ragged M and N; the layer's own output is stored as before.  Against float64."""
    from multi_agent_rl_for_fjsp_b200 import umma

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(M + N + R)
    K, LD = 256, 32
    x = torch.randn(2, M, K, device=dev, generator=g)
    w = torch.randn(2, K, N, device=dev, generator=g) / 16
    b = torch.randn(2, N, device=dev, generator=g)
    wh = torch.randn(2, N, R, device=dev, generator=g) / 8
    bh = torch.randn(2, R, device=dev, generator=g)
    y = torch.zeros(2, M, N, device=dev)
    out = torch.full((M, LD), 7.0, device=dev)
    t = umma.GemmTable(dev, umma.OP_KC, umma.OP_MC)
    offs = (3, 3 + R + 2)
    for i in range(2):
        t.add(x, w, y, M, N, K, lda=K, ldb=N, csm=N, a_off=i * M * K, b_off=i * K * N, c_off=i * M * N, bias=b, bias_off=i * N, relu=True,
              rowdot_w=wh, rowdot_w_off=i * N * R, rowdot_bias=bh, rowdot_bias_off=i * R, rowdot_out=out, rowdot_out_off=offs[i], head_n=R,
              head_ld=LD)
    t.launch()
    touched = torch.zeros(LD, dtype=torch.bool, device=dev)
    for i in range(2):
        h = torch.relu(x[i].double() @ w[i].double() + b[i].double())
        _close(y[i], h, (x[i].double().abs() @ w[i].double().abs()).max().item(), 1e-5)
        ref = h @ wh[i].double() + bh[i].double()
        _close(out[:, offs[i]:offs[i] + R], ref, (h.abs() @ wh[i].double().abs()).max().item(), 1e-5)
        touched[offs[i]:offs[i] + R] = True
    assert (out[:, ~touched] == 7.0).all()   # nothing outside the head's columns is written


def test_short_k_first_layers_as_fma_kernel():
    """``fjsp_a2c_layer1``: relu(x W + b) for K = 3 / 7 / 13 / 38 from slices of 38-wide observation rows, N = 256 and a ragged
    N, ragged row counts, several jobs in one launch; plain fp32 FMAs against float64 (rtol 1e-6 of the absolute-value product)."""
    from multi_agent_rl_for_fjsp_b200 import umma

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(9)
    rows = 1000
    obs = torch.randn(rows, 38, device=dev, generator=g) * 3
    jobs = [(0, 7, 256, True), (7, 13, 256, True), (20, 3, 256, True), (0, 38, 256, True), (5, 4, 130, False)]
    tab, outs, refs = umma.Layer1Table(dev), [], []
    for lo, k, n, relu in jobs:
        w = torch.randn(k, n, device=dev, generator=g)
        b = torch.randn(n, device=dev, generator=g)
        y = torch.full((rows + 3, 256), -7.0, device=dev)
        r = rows if n == 256 else 333
        tab.add(obs, w, y, r, n, k, ldx=38, ldy=256, x_off=lo, bias=b, relu=relu)
        ref = obs[:r, lo:lo + k].double() @ w.double() + b.double()
        outs.append((y, r, n))
        refs.append((torch.relu(ref) if relu else ref, (obs[:r, lo:lo + k].double().abs() @ w.double().abs()).max().item()))
    tab.launch()
    for (y, r, n), (ref, scale) in zip(outs, refs):
        _close(y[:r, :n], ref, scale, 1e-6)
        assert (y[r:] == -7.0).all() and (y[:r, n:] == -7.0).all()


@pytest.mark.parametrize("B,n,na,ld", [(5000, 256, 8, 32), (1537, 256, 3, 32), (2049, 128, 1, 1), (700, 256, 5, 32)])
def test_head_backward_in_one_pass(B, n, na, ld):
    """``fjsp_a2c_head_backward``: dH = (dl W^T) * (H > 0), gb += column sums of dH, gW += H^T dl in one pass over H; dl is a
    column slice of wider rows; gW / gb are accumulated into; ragged row counts; plain fp32 FMAs against float64."""
    from multi_agent_rl_for_fjsp_b200 import umma

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(B + n + na)
    off = 3 if ld > na else 0
    dl = torch.randn(B, ld, device=dev, generator=g)
    W = torch.randn(n, na, device=dev, generator=g) / 4
    H = torch.relu(torch.randn(B + 2, n, device=dev, generator=g))
    dH = torch.full((B + 2, n), -7.0, device=dev)
    gW0, gb0 = torch.randn(n, na, device=dev, generator=g), torch.randn(n, device=dev, generator=g)
    gW, gb = gW0.clone(), gb0.clone()
    t = umma.HeadBwdTable(dev)
    t.add(dl, W, H, dH, gW, gb, B, n, na, ld, dl_off=off)
    t.launch()
    d = dl[:, off:off + na].double()
    ref = (d @ W.double().t()) * (H[:B] > 0)
    scale = (d.abs() @ W.double().abs().t()).max().item()
    _close(dH[:B], ref, scale, 1e-6)
    assert (dH[B:] == -7.0).all()
    _close(gb, gb0.double() + ref.sum(0), ref.abs().sum(0).max().item() + 1.0, 1e-5)
    _close(gW, gW0.double() + H[:B].double().t() @ d, (H[:B].double().abs().t() @ d.abs()).max().item() + 1.0, 1e-5)
