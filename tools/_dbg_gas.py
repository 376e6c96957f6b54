import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers as H
from oracle import checker as K
from mdapy_b200.device import DeviceSystem
g,bg=H.random_gas(3000,30.0,7)
x,y,z=(np.ascontiguousarray(g[:,k]) for k in range(3))
o=np.zeros(3); b=[1,1,1]
for mn in (None,60):
    rv,rd,rn = K.build_neighbor_auto(x,y,z,bg,o,b,4.0) if mn is None else K.build_neighbor(x,y,z,bg,o,b,4.0,mn)
    ds=DeviceSystem(0); ds.set_atoms(x,y,z,bg,o,b)
    M,mx=ds.build_neighbor(4.0,mn)
    v,d,n=ds.fetch_neighbor()
    print(mn,'M',M,rv.shape[1],'mx',mx,rn.max(),'nn equal',np.array_equal(n,rn), 'bad nn', np.flatnonzero(n!=rn)[:10], n[n!=rn][:10], rn[n!=rn][:10])
    if v.shape==rv.shape:
        badrows=np.flatnonzero((v!=rv).any(axis=1)); print('bad rows',badrows.size, badrows[:10])
        if badrows.size:
            i=badrows[0]; print(v[i]); print(rv[i])
