import sys
sys.path.insert(0, '/root/repo')
import torch
from egopose_b200 import lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 300
dev = 'cuda:0'
r = torch.rand(n, dtype=torch.float64, device=dev); m = (torch.rand(n, dtype=torch.float64, device=dev) > 0.11).double(); v = torch.randn(n, dtype=torch.float64, device=dev)
for _ in range(3):
    lib.gae(r, m, v, 0.95, 0.95)
torch.cuda.synchronize()
