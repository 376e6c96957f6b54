"""FCC planar faults (src/identify_fcc_planar_faults.cpp:43) on the device: the reference's own fixture
(tests/fixtures/misc/fcc_planar_faults.npz on input_files/ISF.dump: 103,056 atoms with intrinsic stacking faults,
twin boundaries and multi-layer faults) through this library's PTM + planar-fault kernels, and the drop-in entry
point fed with the REFERENCE's ptm_indices (reference template order) against the reference C++."""
from pathlib import Path

import numpy as np
import pytest

from oracle import checker as K
from oracle import pipeline as P

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _frame():
    d = np.load(GOLD / "fcc_planar_faults.npz")
    return d, P.Frame(d["pos"], d["box"], d["boundary"], d["origin"])


def test_reference_fixture_through_the_system_api():
    import mdapy_b200 as mp

    d, _ = _frame()
    system = mp.System(pos=d["pos"], box=mp.Box(d["box"], d["boundary"], d["origin"]))
    system.cal_polyhedral_template_matching("all", identify_fcc_planar_faults=True, identify_esf=False)
    got = np.asarray(system.data["pft"])
    assert np.array_equal(got, d["pft"]), (f"PFT differs ({int((got != d['pft']).sum())}/{got.size}): "
                                           f"{np.bincount(got, minlength=6)} vs {np.bincount(d['pft'], minlength=6)}")


@pytest.mark.parametrize("esf", [False, True])
def test_drop_in_with_reference_index_order(esf):
    from mdapy_b200.identify_fcc_planar_faults import IdentifyFccPlanarFaults

    d, fr = _frame()
    out, ind = P.cal_ptm(K, fr, "all", 0.1)
    st = out[:, 0].astype(np.int32)
    ref = K.identify_sftb_fcc(st, ind[:, 1:13], identify_esf=esf)
    ours = IdentifyFccPlanarFaults(st, ind[:, 1:13], esf, index_order="reference")
    ours.compute()
    assert np.array_equal(ours.fault_types, ref)


def test_own_ptm_order_equals_reference_pipeline_with_esf():
    """This library's PTM (own template point order) + own tables == reference PTM + reference tables."""
    import mdapy_b200 as mp

    d, fr = _frame()
    out, ind = P.cal_ptm(K, fr, "all", 0.1)
    ref = K.identify_sftb_fcc(out[:, 0].astype(np.int32), ind[:, 1:13], identify_esf=True)
    system = mp.System(pos=d["pos"], box=mp.Box(d["box"], d["boundary"], d["origin"]))
    system.cal_polyhedral_template_matching("all", identify_fcc_planar_faults=True, identify_esf=True)
    assert np.array_equal(np.asarray(system.data["ptm"]), out[:, 0].astype(np.int32))
    assert np.array_equal(np.asarray(system.data["pft"]), ref)
