#!/usr/bin/env python
"""SURVEY 8e test on real NCCL: a 2-rank data-parallel iteration (environments sharded, per-epoch gradient
all-reduce, global advantage moments / exps count) must produce the same parameters as one rank processing the
concatenated batch.  Launch:  torchrun --nproc-per-node 2 tools/check_multi_gpu.py   (prints PASS/FAIL on rank 0).
EGP_CHECK_BACKEND=gloo EGP_CHECK_SAME_DEVICE=1 lets all ranks share cuda:0 (tests/test_gpu_scale.py on a 1-GPU lease)."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egopose_b200.agent import AgentEgo  # noqa: E402
from egopose_b200.config import Config  # noqa: E402
from egopose_b200.env import HumanoidEnv  # noqa: E402
from egopose_b200.mjcf import load_builtin  # noqa: E402
from egopose_b200.nets import MLP, FrameContext, PolicyGaussian, Value  # noqa: E402
from egopose_b200.synthetic import synthetic_cnn_feat, synthetic_takes  # noqa: E402
from egopose_b200.trajbatch import TrajBatchEgo  # noqa: E402


def build(device, E, T, EPL, CD):
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(3)                              # identical initial weights on every rank / run
    cfg = Config('subject_03')
    cfg.env_episode_len = EPL
    env = HumanoidEnv(cfg, device=device.index)
    md = load_builtin()
    env.set_expert_qpos(['a', 'b', 'c'], synthetic_takes(md, 3, 70, seed=4), synthetic_cnn_feat(3, 70, dim=CD))
    S, nu = env.obs_dim, md.nu
    pol = PolicyGaussian(MLP(S + CD, (64, 48), 'relu'), nu, log_std=-2.3, fix_std=True).to(device)
    val = Value(MLP(S + CD, (64, 48), 'relu')).to(device)
    agent = AgentEgo(env=env, dtype=torch.float64, device=device, running_state=None, custom_reward=None,
                     policy_net=pol, policy_vs_net=FrameContext(CD), value_net=val, value_vs_net=FrameContext(CD),
                     optimizer_policy=torch.optim.Adam(pol.parameters(), lr=5e-4),
                     optimizer_value=torch.optim.Adam(val.parameters(), lr=3e-3), opt_num_epochs=3, gamma=0.95, tau=0.95,
                     clip_epsilon=0.2, policy_grad_clip=[(list(pol.parameters()), 0.5)], num_envs=E, horizon=T)
    return agent, pol, val, S, nu


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=32)
    ap.add_argument('--horizon', type=int, default=12)
    args = ap.parse_args()
    world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    backend = os.environ.get('EGP_CHECK_BACKEND', 'nccl')
    if os.environ.get('EGP_CHECK_SAME_DEVICE') == '1':
        local = 0
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if backend == 'nccl':
        dist.init_process_group('nccl', device_id=device)
    else:
        dist.init_process_group(backend)
    E, T, EPL, CD = args.envs, args.horizon, 10, 16
    rng = np.random.RandomState(7)
    rt, rs = rng.randint(0, 3, size=(E, T)), rng.randint(10, 70 - EPL - 10, size=(E, T))
    eps = rng.randn(E * T, 52)
    mean_flag = (rng.rand(E * T) < 0.15).astype(np.uint8)
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)  # noqa: E731

    # ---- data-parallel run: each rank rolls out its shard of the environments
    agent, pol, val, S, nu = build(device, E // world, T, EPL, CD)
    lo, hi = rank * (E // world), (rank + 1) * (E // world)
    par = dict(eps=cu(eps[lo * T:hi * T]), reset_take=cu(rt[lo:hi], torch.int32), reset_start=cu(rs[lo:hi], torch.int32),
               mean_flag=cu(mean_flag[lo * T:hi * T], torch.uint8))
    batch, log = agent.sample(E // world * T, to_host=False, parity=par)
    agent.update_params(batch)
    dp_losses = agent.losses()
    exchange = 'peer memory (csrc/p2p.cu), barrier status %d' % agent._peer.error() if agent._peer is not None else 'torch.distributed all_reduce'
    dp_params = torch.cat([p.detach().reshape(-1) for p in list(pol.parameters()) + list(val.parameters())]).clone()
    gathered = [torch.empty_like(dp_params) for _ in range(world)]
    dist.all_gather(gathered, dp_params)
    replicas_identical = all(torch.equal(gathered[0], g) for g in gathered)
    dist.barrier()
    dist.destroy_process_group()

    # ---- single-process run over the whole batch (rank 0 only, no process group)
    if rank == 0:
        agent1, pol1, val1, _, _ = build(device, E, T, EPL, CD)
        par1 = dict(eps=cu(eps), reset_take=cu(rt, torch.int32), reset_start=cu(rs, torch.int32), mean_flag=cu(mean_flag, torch.uint8))
        batch1, log1 = agent1.sample(E * T, to_host=False, parity=par1)
        agent1.update_params(batch1)
        l1 = agent1.losses()
        p1 = torch.cat([p.detach().reshape(-1) for p in list(pol1.parameters()) + list(val1.parameters())])
        perr = float((dp_params - p1).abs().max() / p1.abs().max())
        lerr = float(np.abs(dp_losses['surr_loss'] - l1['surr_loss']).max())
        verr = float(np.abs(dp_losses['value_loss'] - l1['value_loss']).max() / np.abs(l1['value_loss']).max())
        ok = replicas_identical and perr < 1e-9 and lerr < 1e-10 and verr < 1e-10 and log.num_steps == E * T
        print('world %d: replicas bit-identical %s | params rel err vs 1-GPU %.2e | surr abs err %.2e | vloss rel err %.2e | '
              'logger steps %d avg_c_reward %.6f vs %.6f | gradient exchange: %s -> %s'
              % (world, replicas_identical, perr, lerr, verr, log.num_steps, log.avg_c_reward, log1.avg_c_reward, exchange,
                 'PASS EQUIVALENT' if ok else 'FAIL'))
        sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
