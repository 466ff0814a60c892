"""Per-launch device time of ONE chunk of one optimisation step (policy: PPO loss, value: MSE loss) on the int8 tensor-core
path, in launch order (torch.profiler / CUPTI activity records, no replay):
   python tools/oz_step_trace.py [waves] [S] > profiles/r2_oz_step_trace.md"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from egopose_b200 import lib  # noqa: E402

waves = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = 'cuda'
torch.manual_seed(0)
base = lib.load().egp_oz_mlp_chunk_rows()
n = base * waves


def short(name):
    name = name.replace('void ', '').replace('egp::oz::', '').replace('egp::', '')
    return name.split('(')[0]


def run(dims, kind):
    i, h1, h2, o = dims
    oz = lib.OzMlp(*dims, n_slices=S, chunk_rows=n, device=dev)
    W = [torch.randn(s, device=dev, dtype=torch.float64) * 0.05 for s in [(h1, i), (h1,), (h2, h1), (h2,), (o, h2), (o,)]]
    x = torch.randn(n, i, device=dev, dtype=torch.float64)
    grads = [torch.zeros_like(w) for w in W]
    loss = torch.zeros(1, device=dev, dtype=torch.float64)
    cache = oz.new_cache(n)
    y = oz.step(W, x, cache=cache)
    if kind == 'ppo':
        log_std = torch.full((o,), -2.3, device=dev, dtype=torch.float64)
        actions = torch.randn(n, o, device=dev, dtype=torch.float64) * 0.1
        stats = torch.tensor([float(n), 0.0, float(n - 1)], device=dev, dtype=torch.float64)
        ls = dict(kind='ppo', actions=actions, log_std=log_std, adv=torch.randn(n, device=dev, dtype=torch.float64), stats=stats,
                  logp0=lib.gauss_logp(y, actions, log_std), exps=torch.ones(n, device=dev, dtype=torch.float64), clip_eps=0.2,
                  inv_count=1.0 / n, dlogstd=None, loss=loss)
    else:
        ls = dict(kind='value', returns=torch.randn(n, device=dev, dtype=torch.float64), inv_n=1.0 / n, loss=loss)
    for _ in range(3):
        oz.step(W, x, grads=grads, loss=ls, cache=cache)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        oz.step(W, x, grads=grads, loss=ls, cache=cache)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    print('\n## %s net %s, one chunk of %d rows, S = %d\n' % (kind, dims, n, S))
    print('| # | start us | us | kernel |\n|---|---|---|---|')
    tot = {}
    for k, e in enumerate(ev):
        d = e.time_range.end - e.time_range.start
        print('| %d | %.0f | %.1f | %s |' % (k, e.time_range.start - t0, d, short(e.name)))
        tot[short(e.name)] = tot.get(short(e.name), 0.0) + d
    span = ev[-1].time_range.end - t0
    print('\nspan %.1f us, sum of kernels %.1f us' % (span, sum(tot.values())))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print('* %s: %.1f us' % (k, v))


print('# launch-ordered kernel times of one chunk-step (B200)')
run((243, 300, 300, 52), 'ppo')
run((243, 300, 300, 1), 'value')
