#!/usr/bin/env python
"""Prints the measured max relative difference between the CUDA igrid step and the CPU oracle after each time step
(broadband, non-solenoidal start; the shape of tests/test_igrid_gpu.py::test_igrid_substep_matches_oracle_broadband)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import padeops_b200 as pdo
from conftest import broadband
from oracle import igrid_oracle as IG

nx, ny, nz = 24, 16, 32
L = (2 * np.pi,) * 3
u, v = broadband((nz, ny, nx), 1), broadband((nz, ny, nx), 2)
w = broadband((nz + 1, ny, nx), 3)
w[nz] = w[0]
for scheme in (1, 2):
    ref = IG.IGrid(nx, ny, nz, *L, 50.0, u, v, w, TimeSteppingScheme=scheme)
    g = pdo.igrid()
    g.init(nx, ny, nz, *L, 50.0, u, v, w, TimeSteppingScheme=scheme)
    for it in range(3):
        ref.timeAdvance(0.01)
        g.timeAdvance(0.01)
        errs = {nm: float(np.abs(g.get(nm) - getattr(ref, nm)).max() / np.abs(getattr(ref, nm)).max()) for nm in ("u", "v", "w", "uhat", "what")}
        print(json.dumps({"scheme": "TVD-RK3" if scheme == 1 else "SSP-RK45", "substeps_done": (it + 1) * (3 if scheme == 1 else 5),
                          "max_rel_diff": errs, "max_divergence": g.maxDivergence()}), flush=True)
