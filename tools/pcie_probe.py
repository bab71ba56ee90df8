"""What the PCIe link of this box gives: pinned H2D / D2H copies of the e2e payload sizes."""
import torch, time
n = 1 << 20
for name, nbytes in (("obs 152B/env", n * 152), ("all outputs 220B/env", n * 220), ("actions 8B/env", n * 8)):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for direction in ("d2h", "h2d"):
        for _ in range(3):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True)); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(10):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
        print("%-22s %s %7.3f ms  %6.1f GB/s" % (name, direction, dt * 1e3, nbytes / dt / 1e9))
