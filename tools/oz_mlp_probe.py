"""Time one chunked MLP step (forward + PPO loss + backward) on the int8 tensor cores vs the cuBLAS trunk.
python tools/oz_mlp_probe.py [n_chunks] [S]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from egopose_b200 import lib  # noqa: E402

nch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mult = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dims = (243, 300, 300, 52)
dev = 'cuda'
torch.manual_seed(0)
base = lib.load().egp_oz_mlp_chunk_rows()
oz = lib.OzMlp(*dims, n_slices=S, chunk_rows=base * mult, device=dev)
n = base * nch
i, h1, h2, o = dims
W = [torch.randn(s, device=dev, dtype=torch.float64) * 0.05 for s in [(h1, i), (h1,), (h2, h1), (h2,), (o, h2), (o,)]]
x = torch.randn(n, i, device=dev, dtype=torch.float64)
log_std = torch.full((o,), -2.3, device=dev, dtype=torch.float64)
actions = torch.randn(n, o, device=dev, dtype=torch.float64) * 0.1
adv = torch.randn(n, device=dev, dtype=torch.float64)
exps = torch.ones(n, device=dev, dtype=torch.float64)
stats = torch.tensor([float(n), 0.0, float(n - 1)], device=dev, dtype=torch.float64)
grads = [torch.zeros_like(w) for w in W]
loss = torch.zeros(1, device=dev, dtype=torch.float64)
cache = oz.new_cache(n)
mu = oz.step(W, x, cache=cache)
logp0 = lib.gauss_logp(mu, actions, log_std)
ls = dict(kind='ppo', actions=actions, log_std=log_std, adv=adv, stats=stats, logp0=logp0, exps=exps, clip_eps=0.2, inv_count=1.0 / n,
          dlogstd=None, loss=loss)
for _ in range(2):
    oz.step(W, x, grads=grads, loss=ls, cache=cache)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
reps = 3
for _ in range(reps):
    oz.step(W, x, grads=grads, loss=ls, cache=cache)
t_host = (time.perf_counter() - t0) / reps
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = 2.0 * n * (i * h1 + h1 * h2 + h2 * o) * 3 - 2.0 * n * i * h1
print('chunk x%d' % mult, end=' ')
print('oz mlp step: n %d (%d base chunks) S %d | %.3f ms GPU (%.1f us/chunk), host enqueue %.3f ms | %.1f TFLOP/s f64-equivalent'
      % (n, nch, S, ms, ms * 1e3 / nch, t_host * 1e3, flops / ms * 1e-9), flush=True)
e0.record()
for _ in range(reps):
    oz.step(W, x, cache=cache)
e1.record()
torch.cuda.synchronize()
print('oz mlp forward only: %.3f ms (%.1f us/chunk)' % (e0.elapsed_time(e1) / reps, e0.elapsed_time(e1) / reps * 1e3 / nch))
