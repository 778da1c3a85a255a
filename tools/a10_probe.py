import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import montecarlooptionspricer_b200 as m
from oracle import oracle as O
port=O.port()
eng=m.Engine(0)
for N,n,seed in ((100000,50,1),(20000,50,2),(250,62,3)):
    rng=np.random.default_rng(seed)
    z=rng.standard_normal((N,n)).astype(np.float32).astype(np.float64)
    paths=port.gbm_paths(100.0,0.05,0.2,1.0/n,n,z).astype(np.float32).astype(np.float64)
    for p in (3,4,5,6):
        want=port.lsm(paths,0.05,100.0,1.0,1.0/n,False,p)
        ps=eng.upload_paths(paths,dtype=m.MCP_F32)
        got=eng.lsm_price(ps,0.05,100.0,1.0,1.0/n,False,p,carry=m.MCP_F64,want_first_exercise=True,want_v0=True)
        got32=eng.lsm_price(ps,0.05,100.0,1.0,1.0/n,False,p,carry=m.MCP_F32) if N>4096 else got
        ps.close()
        mism=np.count_nonzero(got.first_exercise!=want["first_ex"])
        print(f"N={N} p={p}: oracle {want['price']:.9f} gpu64 {got.price:.9f} rel {abs(got.price-want['price'])/want['price']:.2e} gpu32 rel {abs(got32.price-want['price'])/want['price']:.2e} "
              f"idx mismatches {mism} ({mism/N:.2e}) maxdV0 {np.max(np.abs(got.v0-want['V0'])):.2e} min_gap {want['min_gap']:.2e}")
eng.close()
