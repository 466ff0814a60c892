#!/usr/bin/env python
"""Per-kernel device time of ONE iteration (rollout + update) at a bench configuration, from torch.profiler (CUPTI activity
records: no replay, so the whole 13 k-launch update is covered, unlike an ncu launch list):
   python tools/update_profile.py [--config 2] > profiles/r2_iteration_kernel_shares.md"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench
    sys.argv = [sys.argv[0]] + [a for a in sys.argv[1:]] + ['--no-variants', '--no-cpu-baseline', '--no-e2e']
    args = bench.parse()
    device = torch.device('cuda', 0)
    torch.cuda.set_device(device)
    _, takes, cnn = bench.synthetic_problem(args)
    agent, cfg = bench.build_agent(args, device, takes, cnn)
    N = args.envs * args.horizon

    def iteration():
        batch, log = agent.sample(N, to_host=False)
        agent.env.end_reward = log.avg_c_reward * cfg.gamma / (1 - cfg.gamma)
        agent.update_params(batch)
        torch.cuda.synchronize()

    for _ in range(2):
        iteration()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        iteration()
    rows = []
    total = 0.0
    for e in prof.key_averages():
        t = getattr(e, 'device_time_total', None)
        if t is None:
            t = getattr(e, 'cuda_time_total', 0.0)
        if t <= 0:
            continue
        rows.append((t / 1e3, e.count, e.key))
        total += t / 1e3
    rows.sort(reverse=True)
    print('# kernel shares of one iteration (%s): torch.profiler / CUPTI activity records\n' % args.workload)
    print('sum of kernel + memcpy device time %.1f ms (streams overlap: the wall time of the iteration is shorter); %d launches\n'
          % (total, sum(r[1] for r in rows)))
    print('| kernel | launches | total ms | share |\n|---|---|---|---|')
    for t, c, k in rows[:40]:
        print('| %s | %d | %.2f | %.1f %% |' % (k[:110].replace('|', '/'), c, t, 100 * t / total))


if __name__ == '__main__':
    main()
