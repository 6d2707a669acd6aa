import time, torch, sys
sys.path.insert(0, ".")
import numpy as np, math
from online_convex_optimization_b200 import engine
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); print("pinned H2D GB/s", 3 * n / (time.perf_counter() - t0) / 1e9)
hp = torch.empty(n, dtype=torch.uint8)
d.copy_(hp); torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(hp); torch.cuda.synchronize()
print("pageable H2D GB/s", n / (time.perf_counter() - t0) / 1e9)
T, dd, B = 10000, 64, 512
z = torch.randn(B, T, dd, dtype=torch.float64).mul_(0.1).pin_memory().numpy(); y = torch.ones(B, T, dtype=torch.float64).pin_memory().numpy()
engine.simulate_batch(z[:32], y[:32], [0], [1.0], comparator="linear")
for _ in range(2):
    t0 = time.perf_counter(); engine.simulate_batch(z, y, [0], [1.0], comparator="linear"); dt = time.perf_counter() - t0
    print("simulate_batch host pinned: GB/s", (z.nbytes + y.nbytes) / dt / 1e9, "sec", dt)
