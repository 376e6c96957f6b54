"""CPU tests of the text readers / writers: LAMMPS dump and XYZ, mirroring the conventions
of the reference's load_save.py (box bounds incl. restricted triclinic tilt conversion, boundary flags, scaled
coordinates, extended-XYZ properties, .gz)."""
import gzip

import numpy as np
import pytest

import mdapy_b200 as mp
from mdapy_b200 import load_save as LS

DUMP_ORTHO = """ITEM: TIMESTEP
40
ITEM: NUMBER OF ATOMS
4
ITEM: BOX BOUNDS pp pp ff
-1.5 6.5
0.0 8.0
2.0 12.0
ITEM: ATOMS id type x y z vx
3 2 0.25 1.0 3.0 -0.5
1 1 -1.25 7.5 11.0 0.125
2 1 6.0 0.0 2.5 1e-3
4 2 2.0 4.0 6.0 0
"""

DUMP_TRI_SCALED = """ITEM: TIMESTEP
0
ITEM: NUMBER OF ATOMS
2
ITEM: BOX BOUNDS xy xz yz pp pp pp
-1.0 11.5 1.5
0.0 10.0 -1.0
0.0 10.0 0.5
ITEM: ATOMS id type xs ys zs
1 1 0.0 0.0 0.0
2 1 0.5 0.5 0.5
"""


def test_read_dump_ortho(tmp_path):
    p = tmp_path / "a.dump"
    p.write_text(DUMP_ORTHO)
    data, box, info = LS.read_dump(str(p))
    assert info["timestep"] == 40 and data.shape[0] == 4
    assert data.columns == ["id", "type", "x", "y", "z", "vx"]
    assert np.asarray(data["id"]).dtype == np.int32 and np.asarray(data["type"]).dtype == np.int32
    assert np.array_equal(np.asarray(data["id"]), [3, 1, 2, 4])            # file order is kept (no sort)
    assert np.allclose(box.box, np.diag([8.0, 8.0, 10.0])) and np.allclose(box.origin, [-1.5, 0.0, 2.0])
    assert list(box.boundary) == [1, 1, 0]
    assert np.asarray(data["vx"])[2] == 1e-3
    system = mp.System(str(p))
    assert system.N == 4 and np.array_equal(np.asarray(system.data["x"]), [0.25, -1.25, 6.0, 2.0])


def test_read_dump_restricted_triclinic_scaled(tmp_path):
    p = tmp_path / "t.dump.gz"
    with gzip.open(p, "wt") as f:
        f.write(DUMP_TRI_SCALED)
    data, box, _ = LS.read_dump(str(p))
    # load_save.py:109-126: xlo = xlo_bound - min(0, xy, xz, xy+xz) ...
    xy, xz, yz = 1.5, -1.0, 0.5
    xlo = -1.0 - min(0.0, xy, xz, xy + xz)
    xhi = 11.5 - max(0.0, xy, xz, xy + xz)
    ylo, yhi = 0.0 - min(0.0, yz), 10.0 - max(0.0, yz)
    want = np.array([[xhi - xlo, 0, 0], [xy, yhi - ylo, 0], [xz, yz, 10.0]])
    assert np.allclose(box.box, want) and np.allclose(box.origin, [xlo, ylo, 0.0])
    pos = np.stack([np.asarray(data[c]) for c in "xyz"], axis=1)
    assert np.allclose(pos[0], box.origin) and np.allclose(pos[1], box.origin + 0.5 * want.sum(axis=0))
    assert "xs" not in data.columns


def test_xyz_roundtrip_and_multiframe_dump(tmp_path):
    rng = np.random.default_rng(0)
    pos = rng.random((50, 3)) * 9.0
    box = mp.Box(np.array([[9.0, 0, 0], [1.0, 9.0, 0], [0.5, -0.7, 9.0]]), [1, 0, 1], [-1.0, 2.0, 0.25])
    frame = mp.Frame({"x": pos[:, 0], "y": pos[:, 1], "z": pos[:, 2], "element": np.array(["Al", "Cu"] * 25, dtype=object),
                      "type": (np.arange(50) % 2 + 1).astype(np.int32), "q": rng.standard_normal(50)})
    px = tmp_path / "f.xyz"
    LS.write_xyz(str(px), box, frame)
    d2, b2, _ = LS.read_xyz(str(px))
    for c in ("x", "y", "z", "q"):
        assert np.array_equal(np.asarray(d2[c]), np.asarray(frame[c])), c      # repr() round-trips doubles exactly
    assert list(d2["element"]) == list(frame["element"]) and np.asarray(d2["type"]).dtype == np.int32
    assert np.array_equal(b2.box, box.box) and np.array_equal(b2.origin, box.origin) and list(b2.boundary) == [1, 0, 1]
    pd_ = tmp_path / "traj.dump"
    with open(pd_, "w") as f:
        for t in range(3):
            one = tmp_path / f"one{t}.dump"
            LS.write_dump(str(one), box, frame.with_columns(x=np.asarray(frame["x"]) + t), timestep=10 * t,
                          columns=["type", "x", "y", "z", "q"])
            f.write(one.read_text())
    frames = list(LS.iter_dump_frames(str(pd_)))
    assert [i["timestep"] for _, _, i in frames] == [0, 10, 20]
    for t, (d, b, _) in enumerate(frames):
        assert np.array_equal(np.asarray(d["x"]), np.asarray(frame["x"]) + t)
        assert np.allclose(b.box, box.box) and np.allclose(b.origin, box.origin) and list(b.boundary) == [1, 0, 1]


def test_classical_xyz_and_errors(tmp_path):
    p = tmp_path / "c.xyz"
    p.write_text("2\ncomment without a cell\nAr 0.0 0.0 0.0\nAr 1.5 2.0 -1.0\n")
    d, b, _ = LS.read_xyz(str(p))
    assert d.columns == ["element", "x", "y", "z"] and list(b.boundary) == [0, 0, 0]
    with pytest.raises(ValueError):
        LS.parse_dump_frame(["ITEM: TIMESTEP\n", "0\n"], "short")
    with pytest.raises(NotImplementedError):
        LS.from_file(str(tmp_path / "x.poscar"))


def test_mp_roundtrip_and_system_helpers(tmp_path):
    pytest.importorskip("pyarrow")
    rng = np.random.default_rng(3)
    pos = rng.random((40, 3)) * 7.0
    box = mp.Box(np.array([[7.0, 0, 0], [0.5, 7.0, 0], [0, 0, 8.0]]), [1, 1, 0], [0.5, -1.0, 2.0])
    frame = mp.Frame({"id": np.arange(1, 41, dtype=np.int32), "type": (np.arange(40) % 3 + 1).astype(np.int32),
                      "x": pos[:, 0], "y": pos[:, 1], "z": pos[:, 2], "element": np.array(["Fe", "Ni"] * 20, dtype=object)})
    p = tmp_path / "model.mp"
    LS.write_mp(str(p), box, frame, {"timestep": 7, "ignored": "x"})
    d, b, info = LS.read_mp(str(p))
    assert info == {"timestep": "7"}
    for c in ("x", "y", "z"):
        assert np.array_equal(np.asarray(d[c]), np.asarray(frame[c]))
    assert np.asarray(d["type"]).dtype == np.int32 and list(d["element"]) == list(frame["element"])
    assert np.array_equal(b.box, box.box) and np.array_equal(b.origin, box.origin) and list(b.boundary) == [1, 1, 0]
    system = mp.System(str(p))
    assert system.N == 40 and system.global_info == {"timestep": "7"}
    system.replicate(2, 1, 3)
    assert system.N == 240 and np.allclose(system.box.box, box.box * np.array([2, 1, 3]).reshape(3, 1))
    # the replicas are the original shifted by whole cell vectors (repeat_cell.cpp:41-59)
    x = np.asarray(system.data["x"])
    assert np.array_equal(x[:40], pos[:, 0])


def test_dump_general_triclinic_and_unwrapped(tmp_path):
    """'abc origin' bounds (general triclinic) round-trip through the writer; xu/yu/zu are promoted to x/y/z
    when no wrapped coordinates are present (load_save.py:96-106, 186-197)."""
    rng = np.random.default_rng(5)
    cell = np.array([[8.0, 0.3, -0.2], [1.0, 7.5, 0.4], [0.2, -0.6, 9.0]])
    box = mp.Box(cell, [1, 1, 1], [1.0, 2.0, 3.0])
    pos = rng.random((10, 3)) @ cell + np.array([1.0, 2.0, 3.0])
    frame = mp.Frame({"id": np.arange(1, 11, dtype=np.int32), "x": pos[:, 0], "y": pos[:, 1], "z": pos[:, 2]})
    p = tmp_path / "g.dump"
    LS.write_dump(str(p), box, frame, timestep=3)
    assert "abc origin" in p.read_text().splitlines()[4]
    d, b, info = LS.read_dump(str(p))
    assert info["timestep"] == 3 and np.array_equal(b.box, cell) and np.array_equal(b.origin, [1.0, 2.0, 3.0])
    assert np.array_equal(np.asarray(d["y"]), pos[:, 1])
    q = tmp_path / "u.dump"
    q.write_text("ITEM: TIMESTEP\n0\nITEM: NUMBER OF ATOMS\n2\nITEM: BOX BOUNDS pp pp pp\n0 5\n0 5\n0 5\n"
                 "ITEM: ATOMS id xu yu zu\n1 6.5 -1.0 2.0\n2 1.0 1.0 1.0\n")
    d, b, _ = LS.read_dump(str(q))
    assert d.columns == ["id", "x", "y", "z"] and np.asarray(d["x"])[0] == 6.5
