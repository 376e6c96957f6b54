"""CPU test of the PTM core arithmetic and its generated tables: tests/host/ptm_host_harness.cpp builds the
__host__ __device__ core of mdapy_b200/csrc/ptm_core.cuh for the CPU (test-only) and the result is compared
with the golden vectors (upstream OVITO `ptm` labels and reference-run `ref_ptm_output`).  Also pins the
generated look-up tables: 1 / 8 / 16 / 1 / 218 template triangulation classes (what the reference ships in
extern/ptm/ptm_graph_data.h:28-34) and rotation groups of order 24 / 12 / 60."""
import ctypes as C
import glob
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import pipeline as P
from oracle import port

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"
SA = sorted(glob.glob(str(GOLD / "sa_*.npz")))
EXACT_TIES = {"vacancy_fcc", "interstitial_fcc", "slab_fcc", "wire_fcc", "perfect_diamond"}
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("ptm") / "libptm_host.so"
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(out),
                    str(ROOT / "tests" / "host" / "ptm_host_harness.cpp")], check=True)
    return C.CDLL(str(out))


def test_generated_tables(harness):
    c = (C.c_int * 12)()
    harness.ptmh_tables_info(c)
    assert list(c)[:8] == [1, 8, 16, 1, 218, 12, 24, 0]      # ptm_graph_data.h:28-34
    assert list(c)[8:11] == [24, 12, 60]
    t = np.zeros((17, 3))
    for s, n in enumerate([6, 12, 12, 12, 14, 16, 16, 9]):
        harness.ptmh_template(s, t.ctypes.data_as(dp))
        assert np.allclose(t[: n + 1].sum(axis=0), 0, atol=1e-12)                     # barycentre 0
        assert abs(np.linalg.norm(t[1: n + 1], axis=1).mean() - 1) < 1e-12            # mean distance 1


@pytest.mark.skipif(not port.available(), reason="port not built")
@pytest.mark.parametrize("path", SA, ids=[Path(p).stem[3:] for p in SA])
def test_host_core_vs_golden(harness, path):
    d = np.load(path)
    fr = P.Frame(d["pos"], d["box"], d["boundary"])
    rep = P.safe_repeat(fr.box, fr.boundary)
    f2 = fr.replicate(port, *rep) if rep.sum() != 3 else fr
    f3, idx, _ = P.nearest(port, f2, 18)
    N = f3.N
    b, o, pb = np.ascontiguousarray(f3.box), np.ascontiguousarray(f3.origin), np.ascontiguousarray(f3.boundary, np.int32)
    idx = np.ascontiguousarray(idx, np.int32)
    t = np.ones(N, np.int32)
    out = np.zeros((N, 8))
    ind = np.zeros((N, 18), np.int32)
    harness.ptmh_index(f3.x.ctypes.data_as(dp), f3.y.ctypes.data_as(dp), f3.z.ctypes.data_as(dp), N,
                       b.ctypes.data_as(dp), o.ctypes.data_as(dp), pb.ctypes.data_as(ip), idx.ctypes.data_as(ip), 18,
                       t.ctypes.data_as(ip), 7, C.c_double(0.1), out.ctypes.data_as(dp), ind.ctypes.data_as(ip))
    out = out[: fr.N]
    assert np.array_equal(out[:, 0].astype(np.int32), d["ptm"])
    ro = d["ref_ptm_output"]
    m = ro[:, 0] > 0
    assert np.allclose(out[m, 2], ro[m, 2], rtol=1e-6, atol=1e-7) and np.allclose(out[m, 3], ro[m, 3], rtol=1e-6)
    dq = np.minimum(np.abs(ro[m, 4:] - out[m, 4:]).max(axis=1), np.abs(ro[m, 4:] + out[m, 4:]).max(axis=1))
    assert dq.size == 0 or dq.max() < 1e-6
    if Path(path).stem[3:] not in EXACT_TIES:
        assert np.allclose(out[:, 2], ro[:, 2], rtol=1e-6, atol=1e-7)
    # all eight structures (flags 255) against the reference run with structure="all"
    out2 = np.zeros((N, 8))
    harness.ptmh_index(f3.x.ctypes.data_as(dp), f3.y.ctypes.data_as(dp), f3.z.ctypes.data_as(dp), N,
                       b.ctypes.data_as(dp), o.ctypes.data_as(dp), pb.ctypes.data_as(ip), idx.ctypes.data_as(ip), 18,
                       t.ctypes.data_as(ip), 255, C.c_double(0.1), out2.ctypes.data_as(dp), ind.ctypes.data_as(ip))
    ra = d["ref_ptm_all_output"]
    assert np.array_equal(out2[: fr.N, 0], ra[:, 0])
    m = ra[:, 0] > 0
    assert np.allclose(out2[: fr.N][m, 2:4], ra[m, 2:4], rtol=1e-6, atol=1e-7)
