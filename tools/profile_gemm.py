"""ncu driver: a few launches of the tcgen05 grouped GEMM on the trainer's update shape (9 x [B,256] x [256,256])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multi_agent_rl_for_fjsp_b200 import umma

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
mode = sys.argv[2] if len(sys.argv) > 2 else "fwd"
L = 9
x = torch.randn(L, B, 256, device=dev)
w = torch.randn(L, 256, 256, device=dev) / 16
b = torch.randn(L, 256, device=dev)
y = torch.empty(L, B, 256, device=dev)
dw = torch.zeros(L, 256, 256, device=dev)
t = umma.GemmTable(dev, *{"fwd": (umma.OP_KC, umma.OP_MC), "dx": (umma.OP_KC, umma.OP_KC), "dw": (umma.OP_MC, umma.OP_MC),
                          "fwdpk": (umma.OP_KC, umma.OP_PK)}[mode])
if mode == "fwdpk":   # what the trainer runs: the weights as packed (hi, lo) images, moved by the B loader's bulk copies
    pk = umma.PackTable(dev)
    offs = [pk.add(w, umma.OP_MC, 256, 256, 256, i * 65536) for i in range(L)]
    pk.finalize().launch()
for i in range(L):
    o = i * B * 256
    if mode == "fwdpk":
        t.add(x, pk.image, y, B, 256, 256, 256, 0, 256, a_off=o, b_off=offs[i], c_off=o, bias=b, bias_off=i * 256, relu=True)
    elif mode == "fwd":
        t.add(x, w, y, B, 256, 256, 256, 256, 256, a_off=o, b_off=i * 65536, c_off=o, bias=b, bias_off=i * 256, relu=True)
    elif mode == "dx":
        t.add(x, w, y, B, 256, 256, 256, 256, 256, a_off=o, b_off=i * 65536, c_off=o, mask=x, mask_off=o)
    else:
        t.add(x, y, dw, 256, 256, B, 256, 256, 256, a_off=o, b_off=o, c_off=i * 65536, atomic=True, splitk=max(1, B // 2048))
for _ in range(4):
    t.launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4):
    t.launch()
e1.record()
torch.cuda.synchronize()
print("%s B=%d: %.3f ms/launch" % (mode, B, e0.elapsed_time(e1) / 4))
