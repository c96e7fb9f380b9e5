#!/usr/bin/env python
"""CD10 on a full 2048^3 cube (64 GiB per array, 2^33 points: past every 32-bit element / byte offset) on one GPU: time of the three
passes, and a spot check of each full-cube result against the same operator applied to small contiguous sub-blocks cut out near the
far end of the array (the operators act line by line, so a sub-block that holds whole lines must reproduce its part exactly)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import padeops_b200 as pdo

n = 2048
d = 2 * np.pi / n
c = pdo.cd10()
assert c.init(n, d) == 0
f = torch.empty((n, n, n), dtype=torch.float64, device="cuda")
f.uniform_()
df = torch.empty_like(f)


def timed(fn, reps=3):
    for _ in range(2):
        fn(f, df)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn(f, df)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for nm, fn in (("dd1", c.dd1), ("dd2", c.dd2), ("dd3", c.dd3)):
    ms = timed(fn)
    # sub-blocks holding whole lines of this axis, at the far end of the other two (field index order is [z][y][x])
    if nm == "dd1":
        sub, got = f[n - 3:, n - 5:, :].contiguous(), df[n - 3:, n - 5:, :]
    elif nm == "dd2":
        sub, got = f[n - 3:, :, n - 64:].contiguous(), df[n - 3:, :, n - 64:]
    else:
        sub, got = f[:, n - 3:, n - 64:].contiguous(), df[:, n - 3:, n - 64:]
    ref = fn(sub)
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"{nm}: {ms:.2f} ms  ({16.0 * n ** 3 / ms / 1e6:.0f} GB/s)  max rel difference vs sub-block result {err:.2e}", flush=True)
