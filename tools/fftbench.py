#!/usr/bin/env python
"""ms per 2-D / 3-D transform through the public fft_3d calls, hand-written passes (default) or cuFFT (PDO_FFT=cufft).
Usage: python tools/fftbench.py [n ...]      bytes: a pass reads its input once and writes its output once (8 B / point each way)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import padeops_b200 as pdo


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [256, 512]
    pdo.decomp_2d.comm_init()
    peak = 6552.0
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for n in sizes:
        ft = pdo.fft_3d()
        assert ft.init(n, n, n, "x", 0.1, 0.1, 0.1) == 0
        f = torch.randn(n, n, n, device="cuda", dtype=torch.float64)
        spec2 = ft.fft2_x2y(f)
        out2 = torch.empty_like(spec2)
        back = torch.empty_like(f)
        unit = 8.0 * n ** 3 / 1e6   # MB per array pass (real array; the half spectrum is the same size + one column)
        rows = {}
        rows["fft2_x2y"] = (timeit(lambda: ft.fft2_x2y(f, out2)), 4)          # x pass + y pass, read + write each
        rows["ifft2_y2x"] = (timeit(lambda: ft.ifft2_y2x(spec2, back)), 4)
        spec3 = ft.fft3_x2z(f)
        out3 = torch.empty_like(spec3)
        rows["fft3_x2z"] = (timeit(lambda: ft.fft3_x2z(f, out3)), 6)
        rows["ifft3_z2x"] = (timeit(lambda: ft.ifft3_z2x(spec3, back)), 6)
        for k, (ms, passes) in rows.items():
            gbs = passes * unit / ms
            print(json.dumps({"n": n, "op": k, "impl": os.environ.get("PDO_FFT", "handwritten"), "ms": round(ms, 4),
                              "GBps_algorithmic": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3)}), flush=True)
        ft.destroy()


if __name__ == "__main__":
    main()
