"""mdapy_b200 -- B200-native (sm_100a) drop-in for mdapy's neighbour + structural-descriptor hot path.

Host side mirrors the reference's Python surface (System.cal_*, Neighbor, NearestNeighbor ...);
all computation runs in hand-written CUDA behind the C ABI of include/mdapy_b200.h.
"""
__version__ = "0.1.0"

from .box import Box  # noqa: E402,F401
from .frame import Frame  # noqa: E402,F401


def __getattr__(name):
    # heavier modules (they dlopen the CUDA library) are imported on first use
    import importlib

    table = {
        "System": ("system", "System"),
        "Neighbor": ("neighbor", "Neighbor"),
        "NearestNeighbor": ("knn", "NearestNeighbor"),
        "DeviceSystem": ("device", "DeviceSystem"),
        "IdentifyDiamondStructure": ("identify_diamond_structure", "IdentifyDiamondStructure"),
        "CommonNeighborParameter": ("common_neighbor_parameter", "CommonNeighborParameter"),
        "WarrenCowleyParameter": ("warren_cowley_parameter", "WarrenCowleyParameter"),
        "ClusterAnalysis": ("cluster_analysis", "ClusterAnalysis"),
        "StructureEntropy": ("structure_entropy", "StructureEntropy"),
        "AtomicTemperature": ("atomic_temperature", "AtomicTemperature"),
        "BondAnalysis": ("bond_analysis", "BondAnalysis"),
        "AngularDistributionFunction": ("bond_analysis", "AngularDistributionFunction"),
        "ChillPlus": ("chill_plus", "ChillPlus"),
        "IdentifyFccPlanarFaults": ("identify_fcc_planar_faults", "IdentifyFccPlanarFaults"),
        "Voronoi": ("voronoi", "Voronoi"),
        "DeviceGroup": ("device", "DeviceGroup"),
        "build_crystal": ("lattice", "build_crystal"),
        "CreatePolycrystal": ("create_polycrystal", "CreatePolycrystal"),
    }
    if name == "empty_cache":
        from ._lib import empty_cache

        return empty_cache
    if name in table:
        mod, attr = table[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(name)
