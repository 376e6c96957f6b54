import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import helpers as H
import mdapy_b200 as mp
from mdapy_b200.device import DeviceSystem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 136
pos, box = H.fcc(4.05, n)
x, y, z = (torch.from_numpy(np.ascontiguousarray(pos[:, k])).pin_memory().numpy() for k in range(3))
del pos
rc = 0.8536 * 4.05
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T(); ds = DeviceSystem(0); t1 = T()
    ds.set_atoms(x, y, z, box, np.zeros(3), [1, 1, 1]); ds.synchronize(); t2 = T()
    ds.build_neighbor(rc); ds.synchronize(); t3 = T()
    c = ds.fcna(rc); t4 = T()
    ds.close(); t5 = T()
    print(f"create {1e3*(t1-t0):.1f} set_atoms {1e3*(t2-t1):.1f} build {1e3*(t3-t2):.1f} fcna+d2h {1e3*(t4-t3):.1f} close {1e3*(t5-t4):.1f}")
    t0 = T(); s = mp.System(data={"x": x, "y": y, "z": z}, box=mp.Box(box)); t1 = T()
    s.cal_common_neighbor_analysis(rc); t2 = T()
    print(f"System ctor {1e3*(t1-t0):.1f} cal_cna {1e3*(t2-t1):.1f}")
