"""one launch of each slicing kernel at the 8-wave chunk size, for ncu"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from egopose_b200 import lib  # noqa: E402
M, F, S = 151552, 300, 6
x = torch.relu(torch.randn(M, F, device='cuda', dtype=torch.float64))
cm = torch.zeros(F, device='cuda', dtype=torch.float64)
for _ in range(2):
    cm.zero_()
    a = lib.oz_slice_rows(x, S, colmax=cm)
    b = lib.oz_slice_colsT(x, S, cm, ones_row=True)
torch.cuda.synchronize()
print('done')
