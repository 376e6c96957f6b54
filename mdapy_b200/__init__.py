"""mdapy_b200 -- B200-native (sm_100a) drop-in for mdapy's neighbour + structural-descriptor hot path.

Host side mirrors the reference's Python surface (System.cal_*, Neighbor, NearestNeighbor ...);
all computation runs in hand-written CUDA behind the C ABI of include/mdapy_b200.h.
"""
__version__ = "0.1.0"
