#!/usr/bin/env python
"""Build the CPU harness of the PTM core (tests/host/ptm_host_harness.cpp) and compare it with the
reference's own C++ (oracle/_ref) on rattled lattices of all five supported structures.
Usage: python tools/ptm_host_check.py [sigma=0.04]"""
import ctypes as C
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers as H  # noqa: E402
from oracle import pipeline as P  # noqa: E402
from oracle import ref  # noqa: E402

dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
out = Path(tempfile.mkdtemp()) / "libptm_host.so"
subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(out),
                str(ROOT / "tests" / "host" / "ptm_host_harness.cpp")], check=True)
lib = C.CDLL(str(out))
sigma = float(sys.argv[1]) if len(sys.argv) > 1 else 0.04


def hcp(a, n):
    c = np.sqrt(8 / 3) * a
    frac = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 5 / 6, 0.5], [0, 1 / 3, 0.5]])
    cell = np.array([a, np.sqrt(3) * a, c])
    g = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 3)
    return ((frac[None] + g[:, None]) * cell).reshape(-1, 3), np.diag(cell * n)


def graphene(a, n):
    """Honeycomb sheet in a box that is open along z."""
    frac = np.array([[0, 0, 0.5], [0.5, 1 / 6, 0.5], [0.5, 0.5, 0.5], [0, 2 / 3, 0.5]])
    cell = np.array([a, np.sqrt(3) * a, 20.0])
    g = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(1), indexing="ij"), -1).reshape(-1, 3)
    return ((frac[None] + g[:, None]) * cell).reshape(-1, 3), np.diag(cell * np.array([n, n, 1]))


cases = {"fcc": H.fcc(3.6, 10), "bcc": H.bcc(2.87, 12), "hcp": hcp(2.9, 9), "sc": H.lattice(np.zeros((1, 3)), 2.6, 14, 14, 14),
         "dcub": H.diamond(3.567, 7), "dhex": H.hex_diamond(2.52, 7, 4, 4), "graphene": graphene(2.46, 24)}
STRUCT = sys.argv[2] if len(sys.argv) > 2 else "all"
FLAGV = {"all": 255, "fcc-hcp-bcc-ico-sc": 31, "dcub-dhex": 96, "graphene": 128}[STRUCT]
FLAGS = {"fcc": 1, "hcp": 2, "bcc": 4, "ico": 8, "sc": 16}
tot = bad = 0
for name, (pos, box) in cases.items():
    for seed, sg in ((1, sigma), (2, 3 * sigma)):
        fr = P.Frame(H.rattle(pos, sg, seed), box, [1, 1, 0] if name == "graphene" else [1, 1, 1])
        f3, idx, _ = P.nearest(ref, fr, 18)
        N = f3.N
        t = (np.random.default_rng(seed).integers(1, 3, N)).astype(np.int32)
        ro, ri = ref.ptm(STRUCT, *f3.geom(), idx, t, 10.0)
        b, o, pb = np.ascontiguousarray(f3.box), np.ascontiguousarray(f3.origin), np.ascontiguousarray(f3.boundary, np.int32)
        idx = np.ascontiguousarray(idx, np.int32)
        go = np.zeros((N, 8)); gi = np.zeros((N, 18), np.int32)
        t0 = time.perf_counter()
        lib.ptmh_index(f3.x.ctypes.data_as(dp), f3.y.ctypes.data_as(dp), f3.z.ctypes.data_as(dp), N,
                       b.ctypes.data_as(dp), o.ctypes.data_as(dp), pb.ctypes.data_as(ip), idx.ctypes.data_as(ip), 18,
                       t.ctypes.data_as(ip), FLAGV, C.c_double(10.0), go.ctypes.data_as(dp), gi.ctypes.data_as(ip))
        dt = time.perf_counter() - t0
        same = (go[:, 0] == ro[:, 0]) & (go[:, 1] == ro[:, 1])
        m = ro[:, 0] > 0
        dq = np.minimum(np.abs(ro[m, 4:] - go[m, 4:]).max(1), np.abs(ro[m, 4:] + go[m, 4:]).max(1))
        dr = np.abs(go[:, 2] - ro[:, 2]).max()
        dd = np.abs(go[:, 3] - ro[:, 3]).max()
        sets = np.array([set(a[a >= 0]) == set(b_[b_ >= 0]) for a, b_ in zip(gi, ri)])
        tot += N; bad += int((~same).sum())
        print(f"{name:4s} sigma={sg:.2f} N={N} types {np.bincount(ro[:,0].astype(int), minlength=9)} mismatch {int((~same).sum())} "
              f"drmsd {dr:.2e} ddist {dd:.2e} dq {dq.max() if dq.size else 0:.2e} index-sets equal {sets.mean():.4f}  host {N/dt/1e3:.1f} k atoms/s")
print("total", tot, "mismatching", bad)
