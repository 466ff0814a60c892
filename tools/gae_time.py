import torch, numpy as np, sys
sys.path.insert(0,'/root/repo')
from egopose_b200 import lib
dev='cuda:0'
for NB in (4096*300, 65536*300):
    rb = (torch.rand(NB, dtype=torch.float64, device=dev), (torch.rand(NB, dtype=torch.float64, device=dev) > 0.11).double(), torch.randn(NB, dtype=torch.float64, device=dev))
    ob = (torch.empty(NB, dtype=torch.float64, device=dev), torch.empty(NB, dtype=torch.float64, device=dev), torch.empty(3, dtype=torch.float64, device=dev))
    wb = torch.empty(lib.load().egp_gae_work_bytes(NB), dtype=torch.uint8, device=dev)
    for mn in (0, 1<<62):
        lib.gae_set_onepass_min(mn)
        bt=[]
        for rep in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); lib.gae(rb[0], rb[1], rb[2], 0.95, 0.95, work=wb, out=ob); b.record(); b.synchronize()
            bt.append(a.elapsed_time(b))
        ms=float(np.median(bt[2:]))
        print(NB, 'onepass' if mn==0 else 'twopass', round(ms*1e3,1),'us', round(40*NB/ms/1e6,1),'GB/s')
