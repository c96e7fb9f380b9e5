import sys, numpy as np, torch
sys.path.insert(0, ".")
import padeops_b200 as pdo
n = 2048
d = 2 * np.pi / n
c = pdo.cd10(); assert c.init(n, d) == 0
f = torch.empty((n, n, n), dtype=torch.float64, device="cuda"); f.uniform_()
df = torch.empty_like(f)
for nm, fn in (("dd3", c.dd3), ("dd2", c.dd2), ("dd3", c.dd3)):
    for _ in range(2): fn(f, df)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): fn(f, df)
    e1.record(); torch.cuda.synchronize()
    print(nm, e0.elapsed_time(e1) / 4, "ms", pdo.lib().pdo_debug_last_variant() if hasattr(pdo.lib(), "pdo_debug_last_variant") else "")
