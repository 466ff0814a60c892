#!/usr/bin/env python
"""Times the float64 GEMM shapes of the PPO update (cuBLAS through torch.mm) and a square DGEMM for scale."""
import torch

torch.set_default_dtype(torch.float64)
dev = 'cuda'
N = 4096 * 300


def t(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / n


sq = torch.randn(8192, 8192, device=dev)
ms = t(lambda: torch.mm(sq, sq))
print('square 8192^3: %.2f ms  %.1f TFLOP/s' % (ms, 2 * 8192 ** 3 / ms / 1e9))
x = torch.randn(N, 243, device=dev)
h1 = torch.randn(N, 300, device=dev)
h2 = torch.randn(N, 300, device=dev)
mu = torch.randn(N, 52, device=dev)
W1, W2, W3 = torch.randn(300, 243, device=dev), torch.randn(300, 300, device=dev), torch.randn(52, 300, device=dev)
o1, o2, o3 = torch.empty_like(h1), torch.empty_like(h2), torch.empty_like(mu)
g1, g2, g3 = torch.empty_like(W1), torch.empty_like(W2), torch.empty_like(W3)
cases = [('fwd1 [N,243]x[243,300]', lambda: torch.mm(x, W1.t(), out=o1), 2 * N * 243 * 300),
         ('fwd2 [N,300]x[300,300]', lambda: torch.mm(h1, W2.t(), out=o2), 2 * N * 300 * 300),
         ('fwd3 [N,300]x[300,52]', lambda: torch.mm(h2, W3.t(), out=o3), 2 * N * 300 * 52),
         ('dW3 [52,N]x[N,300]', lambda: torch.mm(mu.t(), h2, out=g3), 2 * N * 300 * 52),
         ('dh2 [N,52]x[52,300]', lambda: torch.mm(mu, W3, out=o2), 2 * N * 300 * 52),
         ('dW2 [300,N]x[N,300]', lambda: torch.mm(h2.t(), h1, out=g2), 2 * N * 300 * 300),
         ('dh1 [N,300]x[300,300]', lambda: torch.mm(h2, W2, out=o1), 2 * N * 300 * 300),
         ('dW1 [300,N]x[N,243]', lambda: torch.mm(h1.t(), x, out=g1), 2 * N * 300 * 243)]
tot = 0
for name, fn, fl in cases:
    ms = t(fn)
    tot += ms
    print('%-26s %.2f ms  %.1f TFLOP/s' % (name, ms, fl / ms / 1e9))
print('one policy fwd+bwd: %.1f ms' % tot)
