"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol that
include/mdapy_b200.h declares, the ctypes prototypes cover the header, host-only entry points work,
and -- with no GPU -- the product fails loudly instead of falling back to anything."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "mdapy_b200.h").read_text()


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    names = re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_a_sane_surface():
    names = declared_functions()
    for must in ("mdb_build_neighbor", "mdb_build_neighbor_without_max_neigh", "mdb_sort_verlet_by_distance",
                 "mdb_knn", "mdb_fcna", "mdb_acna", "mdb_get_csp", "mdb_compute_aja", "mdb_get_sq",
                 "mdb_identify_solid_liquid", "mdb_rdf", "mdb_rdf_single_species", "mdb_rdf_streaming",
                 "mdb_system_create", "mdb_system_build_neighbor"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from mdapy_b200 import _lib

    lib = C.CDLL(str(_lib.LIB_PATH))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_prototypes_cover_the_header():
    from mdapy_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == declared_functions()
    _lib.lib()  # binds every prototype


def test_host_only_entry_points():
    from mdapy_b200 import _lib
    from mdapy_b200.distributed import cell_grid, slab_bounds

    assert b"sm_100a" in _lib.lib().mdb_version()
    box = np.diag([72.3, 36.0, 10.0])
    assert cell_grid(box, np.zeros(3), [1, 1, 1], 3.0857) == [23, 11, 3]
    # triclinic thickness
    tri = np.array([[10.0, 0, 0], [5.0, 10.0, 0], [0, 0, 10.0]])
    n = cell_grid(tri, np.zeros(3), [1, 1, 1], 2.0)
    assert n == [int(np.floor(10.0 / np.sqrt(1.25) / 2.0)), 5, 5]
    assert slab_bounds(23, 4) == [0, 5, 11, 17, 23]


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from mdapy_b200.device import DeviceSystem

    with pytest.raises(RuntimeError, match="no CUDA device|CUDA error"):
        DeviceSystem(0)
    import mdapy_b200 as mp

    s = mp.System(pos=np.random.rand(50, 3) * 10, box=10.0)
    with pytest.raises(RuntimeError):
        s.build_neighbor(3.0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under mdapy_b200/ may reference it."""
    for p in (ROOT / "mdapy_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h") and p.is_file():
            txt = p.read_text()
            assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, p


def test_host_mirror_classes():
    """Box / Frame / replicate host logic (no GPU involved)."""
    import mdapy_b200 as mp
    from mdapy_b200 import tool_function as tool
    from oracle import port

    b = mp.Box([10, 20, 30], boundary=[1, 0, 1], origin=[1, 2, 3])
    assert not b.triclinic and abs(b.volume - 6000.0) < 1e-9
    assert np.allclose(b.get_thickness(), [10, 20, 30])
    assert list(b.check_small_box(6.0)) == [2, 1, 1]
    t = mp.Box(np.array([[10.0, 0, 0], [5, 10, 0], [0, 0, 10], [1, 1, 1]]))
    assert t.triclinic and np.allclose(t.origin, 1)
    with pytest.raises(ValueError):
        mp.Box(np.zeros((2, 2)))
    fr = mp.Frame({"x": np.arange(4.0), "y": np.arange(4.0), "z": np.arange(4.0), "type": np.arange(4)})
    rep, rb = tool.replicate(fr, mp.Box(5.0), 2, 1, 3)
    assert rep.shape[0] == 24 and np.allclose(np.diag(rb.box), [10, 5, 15])
    assert np.array_equal(np.asarray(rep["type"]), np.tile(np.arange(4), 6))
    if port.available():
        old = np.stack([np.asarray(fr[c]) for c in "xyz"], 1)
        assert np.array_equal(tool.repeat_cell(np.eye(3) * 5.0, old, 2, 1, 3),
                              port.repeat_cell(np.eye(3) * 5.0, old, 2, 1, 3))


def test_devices_keyword_and_group_without_gpu():
    """System(devices=[...]) is host-side bookkeeping until a cal_* call runs; the device group fails loudly (no
    CPU path) when there is no GPU, and rejects an empty device list."""
    import torch

    import mdapy_b200 as mp

    pos = np.random.default_rng(0).random((60, 3)) * 12.0
    with pytest.raises(ValueError):
        mp.System(pos=pos, box=12.0, devices=[])
    s = mp.System(pos=pos, box=12.0, devices=[2, 0, 1])
    assert s._devices == [2, 0, 1] and s._device == 2 and s._group is None      # every other call uses devices[0]
    assert mp.System(pos=pos, box=12.0, device=3)._devices is None
    if torch.cuda.is_available():
        return
    from mdapy_b200.device import DeviceGroup

    with pytest.raises((RuntimeError, ValueError)):
        DeviceGroup([0, 1])
    with pytest.raises(ValueError):
        DeviceGroup([])
    with pytest.raises((RuntimeError, ValueError)):
        s.cal_common_neighbor_analysis(3.0)
