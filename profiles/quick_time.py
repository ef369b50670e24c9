import sys, time, torch
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, R+'/nvalchemi-toolkit-ops_b200', R+'/tests', R+'/oracle']
from systems import bench_box
from nvalchemiops_b200.neighborlist import _engine, neighbor_list
dev='cuda:0'
sizes = [int(a) for a in sys.argv[1:]] or [50_000, 1_000_000]
for n in sizes:
    pos, cell, pbc = bench_box(n, seed=4)
    pos, cell, pbc = pos.to(dev), cell.to(dev), pbc.to(dev)
    csq = 36.0
    for it in range(3):
        h = _engine.build(pos, 6.0, cell, pbc)
        num, ptr = _engine.count(h, csq)
    tot = _engine.status(h)
    print(n, 'status', tot)
    P = tot[0]
    ei = torch.empty((2,P),dtype=torch.int32,device=dev); sh=torch.empty((P,3),dtype=torch.int32,device=dev)
    ev=[torch.cuda.Event(enable_timing=True) for _ in range(5)]
    for it in range(4):
        torch.cuda.synchronize()
        ev[0].record(); h=_engine.build(pos,6.0,cell,pbc)
        ev[1].record(); num,ptr=_engine.count(h,csq)
        ev[2].record(); _engine.fill_coo(h,csq,ptr,ei,sh,P)
        ev[3].record(); torch.cuda.synchronize()
        print(n,'build %.3f count+scan %.3f fill %.3f ms'%(ev[0].elapsed_time(ev[1]),ev[1].elapsed_time(ev[2]),ev[2].elapsed_time(ev[3])))
    B = 12*n+20*P+4*(n+1)
    t = ev[0].elapsed_time(ev[3])*1e-3
    print('total %.3f ms  %.2f GB/s alg  pairs/s %.3e atoms/s %.3e'%(t*1e3,B/t/1e9,P/t,n/t))
