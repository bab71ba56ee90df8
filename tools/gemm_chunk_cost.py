"""Fixed cost and per-K-chunk cost of ONE single-wave grouped launch of fjsp_a2c_gemm (what a rollout step runs): 9 problems x
[M, K] x [K, N] with M rows (default 4096: 288 CTAs, <= 2 per SM), K swept, N = 256 / 128 / 8; CUDA-graph replay of 50
back-to-back launches, device-timed.  Prints one JSON line: microseconds per launch by (N, K) and the fitted slope per 16-wide
K chunk."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multi_agent_rl_for_fjsp_b200 import umma

dev = torch.device("cuda", 0)
M = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
L, KMAX = 9, 512
x = torch.randn(L, M, KMAX, device=dev)
w = torch.randn(L, KMAX, 256, device=dev) / 16
y = torch.empty(L, M, 256, device=dev)
b = torch.randn(L, 256, device=dev)
out = {"rows": M, "problems": L, "packed_b": "--packed" in sys.argv, "us_per_launch": {}}
PACKED = "--packed" in sys.argv   # B as packed (hi, lo) weight images (FJSP_OP_PK) instead of fp32 weights split on the fly
for passes in (3, 1):
    for N in (256, 128, 8):
        pts = []
        for K in (16, 64, 128, 256, 512):
            t = umma.GemmTable(dev, umma.OP_KC, umma.OP_PK if PACKED else umma.OP_MC, passes)
            if PACKED:
                pk = umma.PackTable(dev)
                offs = [pk.add(w, umma.OP_MC, 256, N, K, i * KMAX * 256) for i in range(L)]
                pk.finalize().launch()
            for i in range(L):
                if PACKED:
                    t.add(x, pk.image, y, M, N, K, lda=KMAX, ldb=0, csm=256, a_off=i * M * KMAX, b_off=offs[i], c_off=i * M * 256, bias=b,
                          bias_off=i * 256, relu=True)
                else:
                    t.add(x, w, y, M, N, K, lda=KMAX, ldb=256, csm=256, a_off=i * M * KMAX, b_off=i * KMAX * 256, c_off=i * M * 256, bias=b,
                          bias_off=i * 256, relu=True)
            t.finalize()
            for _ in range(3):
                t.launch()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(50):
                    t.launch()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 200
            pts.append((K // 16, us))
            out["us_per_launch"]["p%d_N%d_K%d" % (passes, N, K)] = round(us, 2)
        (c0, u0), (c1, u1) = pts[1], pts[-1]
        out["us_per_launch"]["p%d_N%d_us_per_chunk" % (passes, N)] = round((u1 - u0) / (c1 - c0), 3)
print(json.dumps(out))
