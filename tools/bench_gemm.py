"""Device-timed throughput of the tcgen05 grouped GEMM (fjsp_a2c_gemm) on the trainer's update shapes:
9 networks x [B, 256] x [256, 256], forward / dx / split-K dW; 3xTF32 and single-pass TF32, cuBLAS fp32 beside it."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multi_agent_rl_for_fjsp_b200 import umma

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
L = 9
x = torch.randn(L, B, 256, device=dev)
w = torch.randn(L, 256, 256, device=dev) / 16
b = torch.randn(L, 256, device=dev)
y = torch.empty(L, B, 256, device=dev)
dw = torch.zeros(L, 256, 256, device=dev)
cs = torch.zeros(L, 256, device=dev)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"B": B, "nets": L, "flop_per_launch": 2.0 * L * B * 256 * 256}
fwd = umma.GemmTable(dev, umma.OP_KC, umma.OP_MC)
dx = umma.GemmTable(dev, umma.OP_KC, umma.OP_KC)
dwt = umma.GemmTable(dev, umma.OP_MC, umma.OP_MC)
sk = max(1, min(64, B // 2048))
for i in range(L):
    o = i * B * 256
    fwd.add(x, w, y, B, 256, 256, 256, 256, 256, a_off=o, b_off=i * 65536, c_off=o, bias=b, bias_off=i * 256, relu=True)
    dx.add(x, w, y, B, 256, 256, 256, 256, 256, a_off=o, b_off=i * 65536, c_off=o, mask=x, mask_off=o, colsum=cs, colsum_off=i * 256)
    dwt.add(x, y, dw, 256, 256, B, 256, 256, 256, a_off=o, b_off=o, c_off=i * 65536, atomic=True, splitk=sk)
# the same two with the weights as packed (hi, lo) images (FJSP_OP_PK): what the trainer runs
pk = umma.PackTable(dev)
offs_f = [pk.add(w, umma.OP_MC, 256, 256, 256, i * 65536) for i in range(L)]
offs_b = [pk.add(w, umma.OP_KCS, 256, 256, 256, i * 65536) for i in range(L)]
pk.finalize().launch()
out["pack_18_weights_ms"] = timed(pk.launch)
fwdp = umma.GemmTable(dev, umma.OP_KC, umma.OP_PK)
dxp = umma.GemmTable(dev, umma.OP_KC, umma.OP_PK)
for i in range(L):
    o = i * B * 256
    fwdp.add(x, pk.image, y, B, 256, 256, 256, 0, 256, a_off=o, b_off=offs_f[i], c_off=o, bias=b, bias_off=i * 256, relu=True)
    dxp.add(x, pk.image, y, B, 256, 256, 256, 0, 256, a_off=o, b_off=offs_b[i], c_off=o, mask=x, mask_off=o, colsum=cs, colsum_off=i * 256)
for name, t in (("forward_bias_relu", fwd), ("forward_bias_relu_packed_w", fwdp), ("dx_mask_colsum", dx), ("dx_mask_colsum_packed_w", dxp),
                ("dw_splitk%d_atomic" % sk, dwt)):
    for p in (3, 1):
        ms = timed(lambda: t.launch(passes=p))
        out["%s_passes%d" % (name, p)] = {"ms": ms, "tflops_fp32_equiv": out["flop_per_launch"] / ms / 1e9,
                                          "tf32_mma_tflops": out["flop_per_launch"] * p / ms / 1e9}
torch.backends.cuda.matmul.allow_tf32 = False
ms = timed(lambda: torch.baddbmm(b[:, None, :], x, w, out=y))
out["cublas_fp32_baddbmm"] = {"ms": ms, "tflops": out["flop_per_launch"] / ms / 1e9}
torch.backends.cuda.matmul.allow_tf32 = True
ms = timed(lambda: torch.baddbmm(b[:, None, :], x, w, out=y))
out["cublas_tf32_baddbmm"] = {"ms": ms, "tflops": out["flop_per_launch"] / ms / 1e9}
print(json.dumps(out, indent=1))
