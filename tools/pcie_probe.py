"""Host<->device copy rates seen by the e2e path (pinned memory, 1-D vs pitched 2-D)."""
import time
import torch

n = 4 * 1024 ** 3 // 4  # 4 GiB of float32
d = torch.empty(n, dtype=torch.float32, device="cuda")
h = torch.empty(n, dtype=torch.float32, pin_memory=True)
for name, fn in (("D2H 1-D", lambda: h.copy_(d, non_blocking=True)), ("H2D 1-D", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%s: %.1f GB/s" % (name, n * 4 / dt / 1e9))
rows, width, pitch = 120, 8639201, 8639232
d2 = torch.empty((rows, pitch), dtype=torch.float32, device="cuda")
h2 = torch.empty((rows, width), dtype=torch.float32, pin_memory=True)
fn = lambda: h2.copy_(d2[:, :width], non_blocking=True)
fn(); torch.cuda.synchronize()
t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("D2H pitched 2-D (torch): %.1f GB/s" % (rows * width * 4 / dt / 1e9))
