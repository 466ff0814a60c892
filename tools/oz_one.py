"""one Ozaki GEMM launch sequence for ncu:  python tools/oz_one.py [M] [N] [K] [S]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from egopose_b200 import lib  # noqa: E402

M, N, K, S = [int(v) for v in (sys.argv[1:5] + ['1228800', '300', '243', '6'][len(sys.argv) - 1:])]
torch.manual_seed(0)
x = torch.randn(M, K, device='cuda', dtype=torch.float64)
w = torch.randn(N, K, device='cuda', dtype=torch.float64)
a, ea = lib.oz_slice_rows(x, S)
b, eb = lib.oz_slice_rows(w, S)
out = torch.empty(M, N, device='cuda', dtype=torch.float64)
for _ in range(3):
    lib.oz_gemm(a, ea, b, eb, out=out)
torch.cuda.synchronize()
print('done')
