#!/usr/bin/env python
"""tools/csp_probe.py -- sort_verlet_by_distance(12) + centro-symmetry(12) on the 10 M-atom frame of BASELINE
configs[1] (cut-off list at 4.2 A, M ~ 19), device resident: time of k_sort_rows(_staged) and k_csp(_fixed).
MDB_SORT=global / MDB_CSP=generic select the round-1 kernels.   python tools/csp_probe.py [n=136]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tools"))
import helpers as H  # noqa: E402
from bench_configs import lattice_dev  # noqa: E402
from mdapy_b200.device import DeviceSystem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 136
dev = torch.device("cuda", 0)
(x, y, z), box = lattice_dev(H.FCC, 3.615, n, 0.02, 1, dev)
N = x.numel()
ds = DeviceSystem(0)
o, b = np.zeros(3), np.array([1, 1, 1], np.int32)
best = {}
for rep in range(3):
    ds.set_atoms_device(x, y, z, box, o, b)
    M, mx = ds.build_neighbor(4.2)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ds.sort_neighbor(12)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ds.csp(12, fetch=False)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    best["sort"] = min(best.get("sort", 1e9), t1 - t0)
    best["csp"] = min(best.get("csp", 1e9), t2 - t1)
c = ds.csp(12)
print(f"N={N} M={M} sort12 {best['sort']*1e3:.2f} ms  csp12 {best['csp']*1e3:.2f} ms  csp sum {float(np.asarray(c).sum()):.9e}")
