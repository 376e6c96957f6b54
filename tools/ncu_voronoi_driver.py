import sys, numpy as np
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import helpers as H
from mdapy_b200.device import DeviceSystem
p,b=H.fcc(3.615,60); pos=H.rattle(p,0.05,0)
x,y,z=(np.ascontiguousarray(pos[:,k]) for k in range(3))
ds=DeviceSystem(0); ds.set_atoms(x,y,z,b,np.zeros(3),np.array([1,1,1],np.int32))
ds.voronoi_volume(); ds.voronoi_volume()
