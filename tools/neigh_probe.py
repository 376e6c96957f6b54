#!/usr/bin/env python
"""tools/neigh_probe.py -- neighbour-kernel time only (CUDA events inside the library), for kernel work.

    python tools/neigh_probe.py [--n 292] [--sigma 0.0] [--reps 5] [--max-neigh 0] [--cna]
Environment switches of the library (MDB_NEIGHBOR, MDB_TILE, MDB_CELLS_NT) apply.
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import A_AL, RC_RATIO, fcc_slab_torch  # noqa: E402
from mdapy_b200.device import DeviceSystem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=292)
ap.add_argument("--sigma", type=float, default=0.0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--max-neigh", type=int, default=0)
ap.add_argument("--rc", type=float, default=RC_RATIO * A_AL)
ap.add_argument("--cna", action="store_true")
ap.add_argument("--shear", type=float, default=0.0, help="triclinic frame: xy tilt factor (xz = xy/2, yz = -xy)")
ap.add_argument("--fused", action="store_true", help="time the fused neighbour + CNA kernel instead")
args = ap.parse_args()
dev = torch.device("cuda", 0)
n, a = args.n, A_AL
x, y, z = fcc_slab_torch(n, a, 0, n, dev)
if args.sigma > 0:
    g = torch.Generator(device=dev).manual_seed(1)
    for t in (x, y, z):
        t += torch.randn(t.shape, generator=g, device=dev, dtype=torch.float64) * args.sigma
box = np.diag([n * a] * 3).astype(float)
if args.shear:
    F = np.array([[1.0, 0.0, 0.0], [args.shear, 1.0, 0.0], [args.shear / 2, -args.shear, 1.0]])
    x, y, z = (x + F[1, 0] * y + F[2, 0] * z).contiguous(), (y + F[2, 1] * z).contiguous(), z
    box = box @ F
ds = DeviceSystem(0)
ds.set_profiling(True)
ts, tc = [], []
for r in range(args.reps + 2):
    ds.set_atoms_device(x, y, z, box, np.zeros(3), np.array([1, 1, 1], np.int32), stream=torch.cuda.current_stream().cuda_stream)
    if args.fused:
        lab, used = ds.fused_cna(args.rc, fetch=False)
        assert used
        M, mx = 0, 0
    else:
        M, mx = ds.build_neighbor(args.rc, args.max_neigh or None)
    t = ds.last_times()
    if args.cna:
        ds.fcna(args.rc, fetch=False)
        tc.append(ds.last_times()["cna_ms"])
    if r >= 2:
        ts.append(t["neighbor_ms"])
N = x.numel()
env = {k: v for k, v in os.environ.items() if k.startswith("MDB_")}
if os.environ.get("MDB_STRIP"):
    print("STRIPPED kernel (tile tables + staging + fp32 copy + row stores, no search): floor measurement")
print(f"N={N} sigma={args.sigma} M={M} max={mx} env={env} neighbor_ms={np.mean(ts):.3f} (min {np.min(ts):.3f}) "
      f"GB/s={(28 + 12 * M) * N / np.mean(ts) / 1e6:.0f}" + (f" cna_ms={np.mean(tc[2:]):.3f}" if tc else ""), flush=True)
