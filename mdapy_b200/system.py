"""``System`` façade for the hot path, mirroring the reference's ``mdapy.System``
(src/mdapy/system.py): ``build_neighbor`` 1108-1166, ``build_nearest_neighbor`` 1226-1263,
``cal_ackland_jones_analysis`` 1605-1636, ``cal_steinhardt_bond_orientation`` 1716-1861,
``cal_polyhedral_template_matching`` 1863-1970, ``cal_centro_symmetry_parameter`` 1972-2003,
``cal_common_neighbor_analysis`` 2005-2064, ``cal_radial_distribution_function`` 2235-2361,
``_get_compute_view`` 765-784 and the cache invalidation of 232-245 / 748-763.

Same neighbour-list reuse policy and result columns; the difference is where
the data lives: coordinates and the neighbour list stay in HBM (one
``DeviceSystem`` per compute view) and ``verlet_list`` / ``distance_list`` /
``neighbor_number`` are materialised as NumPy arrays only when read.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import tool_function as tool
from .ackland_jones_analysis import AcklandJonesAnalysis
from .atomic_temperature import AtomicTemperature
from .bond_analysis import AngularDistributionFunction, BondAnalysis
from .box import Box
from .centro_symmetry_parameter import CentroSymmetryParameter
from .chill_plus import ChillPlus
from .cluster_analysis import ClusterAnalysis
from .common_neighbor_analysis import CommonNeighborAnalysis
from .common_neighbor_parameter import CommonNeighborParameter
from .identify_diamond_structure import IdentifyDiamondStructure
from .device import LIST_CUTOFF, LIST_KNN, DeviceSystem
from .frame import Frame
from .knn import NearestNeighbor
from .neighbor import Neighbor
from .polyhedral_template_matching import PolyhedralTemplateMatching
from .radial_distribution_function import RadialDistributionFunction
from .voronoi import Voronoi
from .steinhardt_bond_orientation import SteinhardtBondOrientation
from .structure_entropy import StructureEntropy
from .warren_cowley_parameter import WarrenCowleyParameter

_LIST_ATTRS = ("verlet_list", "distance_list", "neighbor_number")


class System:
    def __init__(self, filename=None, data=None, pos=None, box=None, device: int = 0, devices=None, **_unused):
        """``device``: the GPU every call runs on.  ``devices=[...]`` (two or more GPUs, one process):
        ``cal_common_neighbor_analysis(rc)`` shards the frame into x slabs over those GPUs (csrc/group.cu: chunked
        upload, peer-store routing over NVLink, fused neighbour + CNA per slab); every other call, and the
        neighbour list a later call may read, uses ``devices[0]``."""
        self.global_info = {}
        if filename is not None:
            # text readers of SURVEY.md 8f.3 (LAMMPS dump, XYZ, optionally .gz); system.py:186-198
            from .load_save import from_file

            self._data, box, self.global_info = from_file(filename)
        elif data is not None and box is not None:
            self._data = Frame.from_any(data)
        elif pos is not None and box is not None:
            pos = np.asarray(pos, np.float64)
            assert pos.ndim == 2 and pos.shape[1] == 3, "pos must be (N, 3)"
            self._data = Frame({"x": pos[:, 0].copy(), "y": pos[:, 1].copy(), "z": pos[:, 2].copy()})
        else:
            raise RuntimeError(
                "One must at least provide filename or [data, box] or [pos, box] or ase_atom or ovito_atom."
            )
        self._box = box if isinstance(box, Box) else Box(box)
        self._devices = None if devices is None else [int(d) for d in devices]
        if self._devices is not None and len(self._devices) == 0:
            raise ValueError("devices must name at least one GPU")
        self._device = int(device) if self._devices is None else self._devices[0]
        self._group = None                           # DeviceGroup over `devices` (created on first use)
        self._dev: Optional[DeviceSystem] = None     # device copy of the compute view
        self._dev_enlarged = False
        self._host_list = {}                         # lazily fetched / user supplied NumPy arrays
        self._host_dirty = False                     # user assigned a list on the host side
        self._has_list = False
        # cut-off of a list the reference would hold at this point but nobody has read yet: the fused
        # neighbour + CNA kernel classified the frame without writing the list (see _materialize_pending)
        self._pending_rc: Optional[float] = None

    def wrap_pos(self) -> None:
        """system.py:854-856: wrap positions into the box for periodic boundaries (resets the neighbour list)."""
        self.update_data(tool.wrap_pos(self.data, self.box), reset_neighbor=True)

    def write_dump(self, filename: str, timestep: int = 0) -> None:
        from .load_save import write_dump

        write_dump(filename, self.box, self.data, timestep)

    def write_xyz(self, filename: str) -> None:
        from .load_save import write_xyz

        write_xyz(filename, self.box, self.data)

    def write_mp(self, filename: str) -> None:
        from .load_save import write_mp

        write_mp(filename, self.box, self.data, self.global_info)

    def replicate(self, nx: int, ny: int, nz: int) -> None:
        """system.py:858-890: replace the frame by its nx x ny x nz supercell (resets the neighbour list)."""
        data, box = tool.replicate(self.data, self.box, nx, ny, nz)
        self._box = box
        self.update_data(data, reset_neighbor=True)

    # ------------------------------------------------------------------ data / box
    @property
    def data(self) -> Frame:
        return self._data

    @property
    def N(self) -> int:
        return self._data.shape[0]

    @property
    def box(self) -> Box:
        return self._box

    @box.setter
    def box(self, value):
        self._box = value if isinstance(value, Box) else Box(value)
        self._reset_neighbor()

    def update_data(self, data, reset_neighbor: bool = False, **_kw) -> None:
        """system.py:700-763: replace the frame; positions changed => caller passes reset_neighbor."""
        self._data = Frame.from_any(data)
        if reset_neighbor:
            self._reset_neighbor()

    def _materialize_pending(self):
        """Build the cut-off list a fused call left out (same rc, automatic width): from here on the state
        is exactly the reference's after ``cal_common_neighbor_analysis(rc)``."""
        rc = self.__dict__.get("_pending_rc")
        if rc is not None:
            self._pending_rc = None
            self.build_neighbor(rc)

    def _reset_neighbor(self):
        self._pending_rc = None
        self._has_list = False
        self._host_list = {}
        self._host_dirty = False
        self._dev = None
        self._dev_enlarged = False
        for a in ("rc", "_enlarge_box", "_enlarge_data"):
            if a in self.__dict__:
                del self.__dict__[a]

    def _get_compute_view(self):
        if "_enlarge_data" in self.__dict__:
            return self._enlarge_box, self._enlarge_data
        return self.box, self.data

    # ------------------------------------------------------------------ lazy neighbour-list attributes
    def __getattr__(self, name):
        # only reached when normal lookup fails
        if name in _LIST_ATTRS or name == "rc":
            if self.__dict__.get("_pending_rc") is not None:
                self._materialize_pending()       # the list is built on first access
                return getattr(self, name)
        if name in _LIST_ATTRS:
            d = self.__dict__
            if not d.get("_has_list", False):
                raise AttributeError(name)
            host = d["_host_list"]
            if name not in host:
                k = _LIST_ATTRS.index(name)
                want = [False, False, False]
                want[k] = True
                host[name] = d["_dev"].fetch_neighbor(*want)[k]
            return host[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in _LIST_ATTRS:
            self._host_list[name] = value
            self._host_dirty = True
            self._has_list = True
            return
        object.__setattr__(self, name, value)

    def __delattr__(self, name):
        if name in _LIST_ATTRS:
            self._host_list.pop(name, None)
            if not self._host_list:
                self._has_list = False
            return
        object.__delattr__(self, name)

    def _device_view(self) -> DeviceSystem:
        """Device copy of the current compute view (uploads once per view)."""
        enlarged = "_enlarge_data" in self.__dict__
        if self._dev is None or self._dev_enlarged != enlarged:
            box, data = self._get_compute_view()
            dev = DeviceSystem(self._device)
            dev.set_atoms(data["x"], data["y"], data["z"], box.box, box.origin, box.boundary)
            self._dev, self._dev_enlarged = dev, enlarged
        return self._dev

    def _group_cna(self, rc: float):
        """FixedCNA labels through the device group (one unpartitioned host frame in, original order out)."""
        from .device import shared_group

        if self._group is None:
            self._group = shared_group(self._devices)   # streams, peer mappings and buffers outlive the frame
        box, data = self.box, self.data
        with self._group.lock:
            self._group.set_atoms(data["x"], data["y"], data["z"], box.box, box.origin, box.boundary)
            return self._group.fused_cna(rc)

    def _device_list(self) -> DeviceSystem:
        """Device view with the cached list in place (pushes a host-assigned list first)."""
        dev = self._device_view()
        if self._host_dirty:
            h = self._host_list
            if "verlet_list" not in h:
                # the reference's consumers read self.verlet_list: a list needs at least its index array
                raise AttributeError("verlet_list")
            dev.put_neighbor(h["verlet_list"], h.get("distance_list"), h.get("neighbor_number"),
                             rc=float(self.__dict__.get("rc", -1.0)),
                             kind=LIST_CUTOFF if "rc" in self.__dict__ else LIST_KNN)
            self._host_dirty = False
        return dev

    def _min_neighbor_number(self) -> int:
        if self._host_dirty and "neighbor_number" in self._host_list:
            return int(np.min(self._host_list["neighbor_number"]))
        return self._device_list().min_count()

    def _sort_neighbor(self, k: int):
        """tool.sort_neighbor on the cached list (mutates it, as the reference does: Appendix D.3)."""
        dev = self._device_list()
        min_number = self._min_neighbor_number()
        assert min_number >= k, f"The min neighbor number {min_number} is lower than k {k}."
        dev.sort_neighbor(k)
        self._host_list.pop("verlet_list", None)
        self._host_list.pop("distance_list", None)

    # ------------------------------------------------------------------ list builders
    def build_neighbor(self, rc: float, max_neigh: Optional[int] = None) -> None:
        self._pending_rc = None    # whatever was pending is replaced by this list
        neigh = Neighbor(rc, self.box, self.data, max_neigh, device=self._device)
        dev = None
        if "_enlarge_data" not in self.__dict__ and sum(self.box.check_small_box(float(rc))) == 3:
            dev = self._device_view()
        try:
            neigh.compute(dev=dev, fetch=False)
        except Exception:
            # (max_neigh too small, ...) the shared device handle may already hold the truncated rows of
            # THIS build: no cached list survives a failed build
            if dev is not None:
                self._reset_neighbor()
            raise
        self.rc = rc
        if hasattr(neigh, "_enlarge_box"):
            self._enlarge_box = neigh._enlarge_box
            self._enlarge_data = neigh._enlarge_data
            self._dev_enlarged = True
        elif "_enlarge_data" in self.__dict__ and dev is None:
            # previous enlarged view is stale for this rc: the list now indexes the original atoms
            del self.__dict__["_enlarge_box"], self.__dict__["_enlarge_data"]
            self._dev_enlarged = False
        self._dev = neigh.dev
        self._host_list = {}
        self._host_dirty = False
        self._has_list = True

    def build_nearest_neighbor(self, k: int) -> None:
        if self._pending_rc is not None:
            # the reference keeps self.rc of the earlier cut-off build (system.py:1262-1263) while the
            # k-nearest list replaces the arrays: no need to build the list that is about to be replaced
            self.rc, self._pending_rc = self._pending_rc, None
        kdt = NearestNeighbor(self.data, self.box, k, device=self._device)
        dev = None
        if "_enlarge_data" not in self.__dict__ and sum(kdt._check_repeat_nearest()) == 3:
            dev = self._device_view()
        kdt.compute(dev=dev, fetch=False)
        if hasattr(kdt, "_enlarge_box"):
            self._enlarge_box = kdt._enlarge_box
            self._enlarge_data = kdt._enlarge_data
            self._dev_enlarged = True
        elif "_enlarge_data" in self.__dict__ and dev is None:
            del self.__dict__["_enlarge_box"], self.__dict__["_enlarge_data"]
            self._dev_enlarged = False
        self._dev = kdt.dev
        self._host_list = {}
        self._host_dirty = False
        self._has_list = True
        # like the reference (system.py:1262-1263), self.rc is left untouched

    def _safe_repeat(self, safe_L=15):
        repeat = np.ceil(safe_L / self.box.get_thickness()).astype(int)
        for i in range(3):
            if self.box.boundary[i] == 0:
                repeat[i] = 1
        return repeat

    # ------------------------------------------------------------------ descriptors
    def cal_common_neighbor_analysis(self, rc: Optional[float] = None, max_neigh: Optional[int] = None):
        use_cached = False
        repeat = self._safe_repeat()
        # Fused path: fixed cut-off, automatic width, no replication, and either no cached list or one with
        # a smaller cut-off -- exactly the cases in which the reference builds a fresh list for rc and
        # classifies from it.  The labels are the same (tests/test_gpu_fused.py); the list itself is built
        # on first access (verlet_list / distance_list / neighbor_number / rc, or any cal_* that reuses it).
        pend = self._pending_rc
        if (rc is not None and max_neigh is None and sum(repeat) == 3 and "_enlarge_data" not in self.__dict__
                and sum(self.box.check_small_box(float(rc))) == 3
                and ((pend is None and "rc" not in self.__dict__) or (pend is not None and pend <= rc)
                     or ("rc" in self.__dict__ and pend is None and self.rc < rc))):
            if self._devices is not None and len(self._devices) > 1:
                labels, used = self._group_cna(float(rc)), True
            else:
                labels, used = self._device_view().fused_cna(float(rc))
            if used:
                self._pending_rc = float(rc) if (pend is None or pend < rc) else pend
                if "rc" in self.__dict__:      # the smaller cached list is replaced (reference: build_neighbor(rc))
                    del self.__dict__["rc"]
                self._has_list = False
                self._host_list = {}
                self._host_dirty = False
                self.update_data(self._data.with_columns(cna=labels[: self.N]))
                return
        self._materialize_pending()
        if sum(repeat) == 3:
            if "rc" in self.__dict__:
                if rc is None:
                    if self._min_neighbor_number() >= 14:
                        self._sort_neighbor(14)
                        use_cached = True
                else:
                    if self.rc < rc:
                        self.build_neighbor(rc, max_neigh)
                        use_cached = True
            else:
                if rc is not None:
                    self.build_neighbor(rc, max_neigh)
                    use_cached = True
        box, data = self._get_compute_view()
        cna = CommonNeighborAnalysis(data, box, rc=rc, dev=self._device_list() if use_cached else None,
                                     device=self._device)
        cna.compute()
        self.update_data(self._data.with_columns(cna=cna.pattern[: self.N]))

    def cal_chill_plus(self, cutoff: float = 3.5) -> None:
        """system.py:1531-1570 -> data['chill_plus'] (0 other, 1 hexagonal ice, 2 cubic ice, 3 interfacial ice,
        4 gas hydrate, 5 interfacial gas hydrate); the frame must hold molecule centres only."""
        has_neigh = "rc" in self.__dict__ and self.rc >= cutoff
        if not has_neigh:
            self.build_neighbor(cutoff)
        box, data = self._get_compute_view()
        cp = ChillPlus(data, box, cutoff, dev=self._device_list())
        cp.compute()
        self.update_data(self._data.with_columns(chill_plus=cp.pattern[: self.N]))

    def build_bond(self, rc, max_neigh: Optional[int] = None) -> np.ndarray:
        """system.py:1330-1411: bond pairs (Nbond, 2), 0-based, i < j, sorted and unique.  ``rc``: one cut-off,
        a {(type_i, type_j): cut-off} / {(element_i, element_j): cut-off} dict, or a matrix over the sorted
        unique types."""
        if np.isscalar(rc):
            max_rc = float(rc)
        elif isinstance(rc, dict):
            max_rc = float(np.max(np.asarray(list(rc.values()), float)))
        else:
            assert "type" in self.data.columns, "Data must contain type column for matrix bond cutoff."
            max_rc = float(np.max(np.asarray(rc, float)))
        assert max_rc > 0, "rc should be larger than 0."
        if not ("rc" in self.__dict__ and self.rc >= max_rc):
            self.build_neighbor(max_rc, max_neigh)
        _, data = self._get_compute_view()
        compact_type, cutoff_matrix = self._normalize_bond_cutoff(rc, data)
        bond = self._device_list().build_bond(compact_type, cutoff_matrix)
        if bond.size == 0:
            self.bond = np.empty((0, 2), np.int32)
            return self.bond
        if "_enlarge_data" in self.__dict__:
            bond %= self.N
        bond.sort(axis=1)
        bond = bond[bond[:, 0] != bond[:, 1]]
        self.bond = np.unique(bond, axis=0) if bond.size else np.empty((0, 2), np.int32)
        return self.bond

    @staticmethod
    def _normalize_bond_cutoff(rc, data):
        """system.py:1266-1328: (compact type per atom, symmetric cut-off matrix)."""
        n = data.shape[0]
        if np.isscalar(rc):
            return np.zeros(n, np.int32), np.array([[float(rc)]], float)
        if isinstance(rc, dict):
            assert len(rc) > 0, "pairwise rc should not be empty."
            first = next(iter(rc))[0]
            col = "element" if isinstance(first, str) else "type"
            assert col in data.columns, f"Data must contain {col} column for pair bond cutoff."
            labels = np.asarray(data[col])
            uniq = np.unique(labels)
            compact = np.searchsorted(uniq, labels).astype(np.int32)
            index = {v: k for k, v in enumerate(uniq.tolist())}
            cm = np.full((uniq.shape[0], uniq.shape[0]), -1.0, float)
            for (a, b), c in rc.items():
                assert a in index and b in index, f"type/element pair {(a, b)} is not in current system."
                assert float(c) > 0, "pairwise rc should be larger than 0."
                cm[index[a], index[b]] = cm[index[b], index[a]] = float(c)
            assert not np.any(cm < 0), "every type pair needs a bond cutoff"
            return compact, cm
        labels = np.asarray(data["type"])
        uniq = np.unique(labels)
        cm = np.asarray(rc, float)
        assert cm.shape == (uniq.shape[0], uniq.shape[0]), "cutoff matrix must be (ntype, ntype)"
        assert np.allclose(cm, cm.T), "cutoff matrix must be symmetric"
        return np.searchsorted(uniq, labels).astype(np.int32), cm

    def cal_identify_diamond_structure(self):
        """system.py:1493-1529 -> data['ids'] (0 other, 1-3 cubic diamond + shells, 4-6 hexagonal)."""
        dev = None
        safe_L = 15
        repeat = np.ceil(safe_L / self.box.get_thickness()).astype(int)
        for i in range(3):
            if self.box.boundary[i] == 0:
                repeat[i] = 1
        if sum(repeat) == 3 and self._has_list and "rc" in self.__dict__:
            if self._min_neighbor_number() >= 4:
                self._sort_neighbor(4)
                dev = self._device_list()
        box, data = self._get_compute_view()
        ids = IdentifyDiamondStructure(data, box, dev=dev, device=self._device)
        ids.compute()
        self.update_data(self._data.with_columns(ids=ids.pattern[: self.N]))

    def cal_centro_symmetry_parameter(self, N: int):
        assert N % 2 == 0 and N > 0, f"N must be a positive even number: {N}."
        if self.N <= N and sum(self.box.boundary) == 0:
            res = np.full(self.N, 10000, float)
        else:
            has_verlet = False
            if self._has_list:
                if self._min_neighbor_number() >= N and "rc" in self.__dict__:
                    self._sort_neighbor(N)
                    has_verlet = True
            if not has_verlet:
                self.build_nearest_neighbor(N)
            box, data = self._get_compute_view()
            csp = CentroSymmetryParameter(data, box, N, dev=self._device_list())
            csp.compute()
            res = csp.csp[: self.N]
        self.update_data(self.data.with_columns(csp=res))

    def cal_ackland_jones_analysis(self) -> None:
        N_neigh = 14
        if self.data.shape[0] < N_neigh and sum(self.box.boundary) == 0:
            self.update_data(self._data.with_columns(aja=np.zeros(self.N, np.int32)))
            return
        if self._has_list and self._min_neighbor_number() >= N_neigh:
            self._sort_neighbor(N_neigh)
        else:
            self.build_nearest_neighbor(N_neigh)
        box, data = self._get_compute_view()
        aja = AcklandJonesAnalysis(data, box, dev=self._device_list())
        aja.compute()
        self.update_data(self._data.with_columns(aja=aja.aja[: self.N]))

    # ---- further list consumers
    def _ensure_cutoff_list(self, rc: float, max_neigh: Optional[int] = None):
        """system.py:1591-1596 / 1666-1671: reuse the cached list when it reaches at least rc."""
        has_neigh = "rc" in self.__dict__ and self.rc >= rc
        if not has_neigh:
            self.build_neighbor(rc, max_neigh)

    def cal_common_neighbor_parameter(self, rc: float, max_neigh: Optional[int] = None) -> None:
        """system.py:1572-1603 -> data['cnp']."""
        self._ensure_cutoff_list(rc, max_neigh)
        box, data = self._get_compute_view()
        cnp = CommonNeighborParameter(data, box, rc, dev=self._device_list(), device=self._device)
        cnp.compute()
        self.update_data(self._data.with_columns(cnp=cnp.cnp[: self.N]))

    def cal_warren_cowley_parameter(self, rc: float, max_neigh: Optional[int] = None) -> WarrenCowleyParameter:
        """system.py:1638-1676: returns the object holding the (Ntype, Ntype) matrix ``WCP``."""
        self._ensure_cutoff_list(rc, max_neigh)
        _, data = self._get_compute_view()
        wcp = WarrenCowleyParameter(None, None, data, dev=self._device_list(), device=self._device)
        wcp.compute()
        return wcp

    def average_by_neighbor(self, average_rc: float, property_name: str, include_self: bool = True,
                            output_name: Optional[str] = None, max_neigh: Optional[int] = None) -> None:
        """system.py:2363-2414 -> data[output_name or f'{property_name}_ave']."""
        assert property_name in self._data.columns, f"{property_name} not in data."
        if "rc" in self.__dict__:
            if self.rc < average_rc:
                self.build_neighbor(average_rc, max_neigh)
        else:
            self.build_neighbor(average_rc, max_neigh)
        assert "_enlarge_data" not in self.__dict__, (
            "average_by_neighbor only supports systems whose box is large enough "
            "that no replica was built (i.e. self._enlarge_data must not exist)."
        )
        value = np.asarray(self._data[property_name], np.float64)
        ave = self._device_list().average_by_neighbor(average_rc, value, include_self)
        name = output_name if output_name is not None else f"{property_name}_ave"
        self.update_data(self._data.with_columns(**{name: np.asarray(ave[: self.N]).copy()}))

    def cal_cluster_analysis(self, rc=5.0, max_neigh: Optional[int] = None) -> None:
        """system.py:2416-2478 -> data['cluster_id']; rc is a number or a dict like {'1-1': 1.5, '1-2': 1.3}."""
        if isinstance(rc, (int, float, np.integer, np.floating)):
            max_rc = float(rc)
        elif isinstance(rc, dict):
            max_rc = max([i for i in rc.values()])
        else:
            raise TypeError("rc should be a positive number, or a dict like {'1-1':1.5, '1-2':1.3}")
        if "rc" in self.__dict__:
            if self.rc < max_rc:
                self.build_neighbor(max_rc, max_neigh)
        else:
            self.build_neighbor(max_rc, max_neigh)
        type_list = None
        if isinstance(rc, dict):
            assert "type" in self.data.columns, "Must have type for multi rc cluster calculation."
            _, view = self._get_compute_view()
            type_list = np.asarray((view if "type" in view.columns else self.data)["type"]).astype(np.int32)
        ca = ClusterAnalysis(rc, type_list=type_list, dev=self._device_list())
        ca.compute()
        self.update_data(self.data.with_columns(cluster_id=np.asarray(ca.particleClusters[: self.N]).copy()))

    def cal_structure_entropy(self, rc: float, sigma: float, use_local_density: bool = False, average_rc: float = 0.0,
                              max_neigh: Optional[int] = None) -> None:
        """system.py:2480-2542 -> data['entropy'] (+ data['entropy_ave'] when average_rc > 0)."""
        if "rc" in self.__dict__:
            if self.rc < rc:
                self.build_neighbor(rc, max_neigh)
        else:
            self.build_neighbor(rc, max_neigh)
        box, _ = self._get_compute_view()
        SE = StructureEntropy(box, rc=rc, sigma=sigma, use_local_density=use_local_density, average_rc=average_rc,
                              dev=self._device_list())
        SE.compute()
        cols = {"entropy": np.asarray(SE.entropy[: self.N]).copy()}
        if average_rc > 0:
            cols["entropy_ave"] = np.asarray(SE.entropy_ave[: self.N]).copy()
        self.update_data(self.data.with_columns(**cols))

    def cal_atomic_temperature(self, rc: float, factor: float = 1.0, max_neigh: Optional[int] = None) -> None:
        """system.py:1678-1714 -> data['atomic_temp'] in K (velocities in A/fs times ``factor``)."""
        self._ensure_cutoff_list(rc, max_neigh)
        _, data = self._get_compute_view()
        at = AtomicTemperature(data, rc=rc, factor=factor, dev=self._device_list())
        at.compute()
        self.update_data(self.data.with_columns(atomic_temp=np.asarray(at.T[: self.N]).copy()))

    def cal_bond_analysis(self, rc: float, nbin: int, max_neigh: Optional[int] = None) -> BondAnalysis:
        """system.py:2130-2178: bond-length and bond-angle histograms of the cut-off list."""
        self._ensure_cutoff_list(rc, max_neigh)
        box, data = self._get_compute_view()
        ba = BondAnalysis(data, box, rc, nbin, dev=self._device_list(), device=self._device)
        ba.compute()
        return ba

    def cal_angular_distribution_function(self, rc_dict, nbin: int,
                                          max_neigh: Optional[int] = None) -> AngularDistributionFunction:
        """system.py:2180-2233: angle histograms per 'A-B-C' element triplet, rc_dict[key] = [rij_min, rij_max, rik_min, rik_max]."""
        rc = float(np.array(list(rc_dict.values())).max())
        self._ensure_cutoff_list(rc, max_neigh)
        box, data = self._get_compute_view()
        adf = AngularDistributionFunction(data, box, rc_dict, nbin, dev=self._device_list(), device=self._device)
        adf.compute()
        return adf

    def cal_steinhardt_bond_orientation(self, llist, use_voronoi: bool = False, nnn: int = 0, rc: float = -1.0,
                                        average: bool = False, use_weight: bool = False, weight=None,
                                        wl: bool = False, wlhat: bool = False, a_face_area_threshold: float = -1,
                                        r_face_area_threshold: float = -1, identify_liquid: bool = False,
                                        threshold: float = 0.7, n_bond: int = 7, max_neigh: Optional[int] = None):
        if use_voronoi:
            # system.py:1781-1789: the Voronoi rows replace the cut-off / k-nearest list for this call only
            self.build_voronoi_neighbor(a_face_area_threshold, r_face_area_threshold)
            if use_weight and weight is None:
                weight = self.voro_face_area
            box, data = self._get_compute_view()
            SBO = SteinhardtBondOrientation(box, data, np.asarray(llist, int), nnn, rc, average, True, use_weight,
                                            weight, self.voro_verlet_list, self.voro_distance_list,
                                            self.voro_neighbor_number, wl=wl, wlhat=wlhat,
                                            identify_liquid=identify_liquid, threshold=threshold, n_bond=n_bond,
                                            device=self._device)
            SBO.compute()
            return self._store_steinhardt(SBO, llist, wl, wlhat, identify_liquid)
        if nnn > 0:
            has_sort_neigh = False
            if self._has_list and self._min_neighbor_number() >= nnn:
                self._sort_neighbor(nnn)
                has_sort_neigh = True
            if not has_sort_neigh:
                self.build_nearest_neighbor(nnn)
        else:
            assert rc > 0, "At least use voronoi, or set positive nnn, or positive rc."
            if "rc" in self.__dict__:
                if self.rc < rc:
                    self.build_neighbor(rc, max_neigh)
            else:
                self.build_neighbor(rc, max_neigh)
        box, data = self._get_compute_view()
        SBO = SteinhardtBondOrientation(box, data, np.asarray(llist, int), nnn, rc, average, use_voronoi, use_weight,
                                        weight, wl=wl, wlhat=wlhat, identify_liquid=identify_liquid,
                                        threshold=threshold, n_bond=n_bond, dev=self._device_list())
        SBO.compute()
        return self._store_steinhardt(SBO, llist, wl, wlhat, identify_liquid)

    def _store_steinhardt(self, SBO, llist, wl, wlhat, identify_liquid):
        cols = {}
        if SBO.qnarray.shape[1] > 1:
            names = [f"ql{i}" for i in llist]
            if wl:
                names.extend(f"wl{i}" for i in llist)
            if wlhat:
                names.extend(f"wlh{i}" for i in llist)
            for i, name in enumerate(names):
                cols[name] = SBO.qnarray[: self.N, i].copy()
        else:
            cols[f"ql{llist[0]}"] = SBO.qnarray.flatten()[: self.N]
        if identify_liquid:
            cols["solidliquid"] = SBO.solidliquid[: self.N]
            cols["nbond"] = SBO.nbond[: self.N]
        self.update_data(self.data.with_columns(**cols))
        return SBO

    def build_voronoi_neighbor(self, a_face_area_threshold: float = -1.0, r_face_area_threshold: float = -1.0) -> None:
        """system.py:1168-1224 -> voro_verlet_list / voro_distance_list / voro_face_area / voro_neighbor_number
        (rows list each cell's faces; walls and faces under the area threshold hold -1)."""
        vor = Voronoi(self.box, self.data, dev=self._device_view() if "_enlarge_data" not in self.__dict__ else None,
                      device=self._device)
        (self.voro_verlet_list, self.voro_distance_list, self.voro_face_area,
         self.voro_neighbor_number) = vor.get_neighbor(a_face_area_threshold, r_face_area_threshold)
        if hasattr(vor, "_enlarge_box"):
            self._enlarge_box = vor._enlarge_box
            self._enlarge_data = vor._enlarge_data
            self._dev = None            # the device copy described the original frame
            self._dev_enlarged = False

    def cal_voronoi_volume(self) -> None:
        """system.py:2544-2573 -> data['volume'], data['neighbor_number'], data['cavity_radius']."""
        vor = Voronoi(self.box, self.data, dev=self._device_view() if "_enlarge_data" not in self.__dict__ else None,
                      device=self._device)
        volume, neighbor_number, cavity_radius = vor.get_volume()
        self.update_data(self._data.with_columns(volume=volume, neighbor_number=neighbor_number,
                                                 cavity_radius=cavity_radius))

    def cal_radial_distribution_function(self, rc: float, nbin: int = 200, max_neigh: Optional[int] = None,
                                         streaming: Optional[bool] = None) -> RadialDistributionFunction:
        box, data = self._get_compute_view()
        if streaming is None:
            thickness = box.get_thickness()
            per = [thickness[i] for i in range(3) if box.boundary[i]]
            min_thick = min(per) if per else float("inf")
            streaming = rc >= min_thick / 3.0

        def _species_labels(view):
            if "element" in view.columns:
                return np.asarray(view["element"])
            if "type" in view.columns:
                return np.asarray(view["type"])
            return np.zeros(view.shape[0], np.int32)

        type_list = _species_labels(data)
        if streaming:
            repeat = self.box.check_small_box(rc)
            if sum(repeat) != 3:
                rep_data, rep_box = tool.replicate(data, box, *repeat)
                rdf = RadialDistributionFunction(rc, nbin, rep_box, type_list=_species_labels(rep_data), streaming=True,
                                                 x=rep_data["x"], y=rep_data["y"], z=rep_data["z"],
                                                 device=self._device)
            else:
                rdf = RadialDistributionFunction(rc, nbin, box, type_list=type_list, streaming=True,
                                                 dev=self._device_view(), device=self._device)
        else:
            has_neigh = "rc" in self.__dict__ and self.rc >= rc
            if not has_neigh:
                self.build_neighbor(rc, max_neigh)
            box, data = self._get_compute_view()
            type_list = _species_labels(data)
            rdf = RadialDistributionFunction(rc, nbin, box, type_list=type_list, dev=self._device_list(),
                                             device=self._device)
        rdf.compute()
        return rdf

    def cal_polyhedral_template_matching(self, structure="fcc-hcp-bcc", rmsd_threshold=0.1, return_ordering=False,
                                         return_rmsd=False, return_atomic_distance=False, return_orientation=False,
                                         identify_fcc_planar_faults=False, identify_esf=True):
        use_cached = False
        repeat = self._safe_repeat()
        if sum(repeat) == 3 and self._has_list and self._min_neighbor_number() >= 18:
            self._sort_neighbor(18)
            use_cached = True
        box, data = self._get_compute_view()
        ptm = PolyhedralTemplateMatching(structure, data, box, rmsd_threshold,
                                         dev=self._device_list() if use_cached else None, device=self._device)
        ptm.compute()
        output = ptm.output[: self.N]
        cols = {"ptm": output[:, 0].astype(np.int32)}
        if output.shape[1] >= 8:
            if return_ordering:
                cols["ordering"] = output[:, 1].copy()
            if return_rmsd:
                cols["rmsd"] = output[:, 2].copy()
            if return_atomic_distance:
                cols["interatomic_distance"] = output[:, 3].copy()
            if return_orientation:
                cols.update(qx=output[:, 5].copy(), qy=output[:, 6].copy(), qz=output[:, 7].copy(), qw=output[:, 4].copy())
        if identify_fcc_planar_faults:
            # system.py:1963-1968: planar faults from the structure types and the 12 matched neighbours of
            # every atom (this library's template point order, see identify_fcc_planar_faults.py)
            from .identify_fcc_planar_faults import IdentifyFccPlanarFaults

            ifpt = IdentifyFccPlanarFaults(np.asarray(ptm.output[:, 0], np.int32),
                                           np.ascontiguousarray(ptm.ptm_indices[:, 1:13]), identify_esf)
            ifpt.compute()
            cols["pft"] = ifpt.fault_types[: self.N].copy()
        self.update_data(self._data.with_columns(**cols))
        return ptm


def _after_pending(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        self._materialize_pending()
        return fn(self, *a, **k)

    return wrapper


for _name in list(vars(System)):
    if (_name.startswith("cal_") and _name != "cal_common_neighbor_analysis") or _name in (
            "average_by_neighbor", "build_bond", "_ensure_cutoff_list", "_min_neighbor_number", "_sort_neighbor",
            "_device_list"):
        setattr(System, _name, _after_pending(getattr(System, _name)))
del _name
