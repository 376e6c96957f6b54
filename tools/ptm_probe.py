#!/usr/bin/env python
"""PTM / kNN timing probe: rattled BCC Fe, kNN(18) + PTM (default structures), device resident.
Usage: python tools/ptm_probe.py [n_cells=100] [structure=fcc-hcp-bcc]"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
structure = sys.argv[2] if len(sys.argv) > 2 else "fcc-hcp-bcc"
dev = torch.device("cuda", 0)
(x, y, z), box = lattice_dev(H.BCC, 2.8665, n, 0.03, 7, dev)
N = x.numel()
ds = DeviceSystem(0)
ds.set_atoms_device(x, y, z, box, np.zeros(3), [1, 1, 1])
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ds.build_knn(18)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ds.ptm(structure, fetch=False)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"N={N} knn18 {1e3*(t1-t0):.2f} ms ({N/(t1-t0)/1e6:.1f} M/s)  ptm {1e3*(t2-t1):.2f} ms ({N/(t2-t1)/1e6:.2f} M/s)", flush=True)
out, _ = ds.ptm(structure, fetch=True)
print("types", np.bincount(out[:, 0].astype(int), minlength=9))
