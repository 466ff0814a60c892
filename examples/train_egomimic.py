#!/usr/bin/env python
"""Training loop with the structure of ego_pose/ego_mimic.py:93-144 (pre_iter_update -> sample -> end_reward ->
update_params -> log -> checkpoint) on the B200-native path, with synthetic experts / CNN features because the
EgoPose dataset is not redistributable.  With the real dataset and the reference checkout, run the reference's own
script through the import shim instead.

  python examples/train_egomimic.py --iters 20 --envs 1024 --horizon 50
  torchrun --nproc-per-node 8 examples/train_egomimic.py ...        (environments sharded, gradients all-reduced)
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from egopose_b200 import checkpoint  # noqa: E402
from egopose_b200.agent import AgentEgo  # noqa: E402
from egopose_b200.config import Config  # noqa: E402
from egopose_b200.env import HumanoidEnv  # noqa: E402
from egopose_b200.nets import MLP, FrameContext, PolicyGaussian, Value  # noqa: E402
from egopose_b200.synthetic import synthetic_cnn_feat, synthetic_takes  # noqa: E402
from egopose_b200.torch_utils import set_optimizer_lr  # noqa: E402
from egopose_b200.zfilter import ZFilter  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default='subject_03')
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--envs', type=int, default=1024)
    ap.add_argument('--horizon', type=int, default=50)
    ap.add_argument('--takes', type=int, default=8)
    ap.add_argument('--save', default='')
    ap.add_argument('--physics', choices=['smooth', 'full'], default='smooth',
                    help='full = joint limits + floor contact (MuJoCo soft constraints, one-warp rollout kernel)')
    ap.add_argument('--standing', action='store_true',
                    help='expert = standing still on the floor (T-pose, soles on z = 0, random heading per take): a physically '
                         'consistent target for --physics full, where the policy has to learn to keep the balance')
    ap.add_argument('--policy-lr', type=float, default=0.0, help='override cfg.policy_lr')
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', 1), ('RANK', 0), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=device)
    dtype = torch.float64
    torch.set_default_dtype(dtype)
    cfg = Config(args.cfg)
    cfg.env_episode_len = args.horizon
    np.random.seed(cfg.seed)
    torch.manual_seed(cfg.seed)

    env = HumanoidEnv(cfg, device=local)
    env.seed(cfg.seed + 1000 * rank)
    L = args.horizon + 2 * cfg.fr_margin + 64
    if args.physics == 'full':
        env.kernel.set_joint_limits(True)
        env.kernel.set_contacts(True)
    takes = synthetic_takes(env.md, args.takes, L, seed=1)
    if args.standing:
        rng = np.random.RandomState(3)
        for q in takes:
            yaw = rng.uniform(-np.pi, np.pi)
            q[:] = np.asarray(env.md.qpos0)
            q[:, 2] = 0.8665                                            # soles of the foot boxes on the floor
            q[:, 3:7] = [np.cos(yaw / 2), 0.0, 0.0, np.sin(yaw / 2)]
    env.set_expert_qpos(['take_%d' % i for i in range(args.takes)], takes, synthetic_cnn_feat(args.takes, L))
    state_dim, action_dim = env.observation_space.shape[0], env.action_space.shape[0]
    running_state = ZFilter((state_dim,), clip=5)
    policy_vs_net, value_vs_net = FrameContext(128, cfg.policy_v_hdim, cfg.fr_margin), FrameContext(128, cfg.value_v_hdim, cfg.fr_margin)
    policy_net = PolicyGaussian(MLP(state_dim + cfg.policy_v_hdim, cfg.policy_hsize, cfg.policy_htype), action_dim,
                                log_std=cfg.log_std, fix_std=cfg.fix_std).to(device)
    value_net = Value(MLP(state_dim + cfg.value_v_hdim, cfg.value_hsize, cfg.value_htype)).to(device)
    optimizer_policy = torch.optim.Adam(policy_net.parameters(), lr=cfg.policy_lr, weight_decay=cfg.policy_weightdecay)
    optimizer_value = torch.optim.Adam(value_net.parameters(), lr=cfg.value_lr, weight_decay=cfg.value_weightdecay)
    agent = AgentEgo(env=env, dtype=dtype, device=device, running_state=running_state, custom_reward=None,
                     num_threads=1, policy_net=policy_net, policy_vs_net=policy_vs_net, value_net=value_net,
                     value_vs_net=value_vs_net, optimizer_policy=optimizer_policy, optimizer_value=optimizer_value,
                     opt_num_epochs=cfg.num_optim_epoch, gamma=cfg.gamma, tau=cfg.tau, clip_epsilon=cfg.clip_epsilon,
                     policy_grad_clip=[(list(policy_net.parameters()), 40)], num_envs=args.envs, horizon=args.horizon)

    for i_iter in range(args.iters):
        cfg.update_adaptive_params(i_iter)                              # ego_mimic.py:93-99
        agent.set_noise_rate(cfg.adp_noise_rate)
        if args.policy_lr > 0:
            cfg.adp_policy_lr = args.policy_lr
        set_optimizer_lr(optimizer_policy, cfg.adp_policy_lr)
        if cfg.fix_std:
            policy_net.action_log_std.fill_(cfg.adp_log_std)
        batch, log = agent.sample(args.envs * args.horizon, to_host=False)
        agent.env.end_reward = log.avg_c_reward * cfg.gamma / (1 - cfg.gamma)
        t_update = agent.update_params(batch)
        losses = agent.losses()
        if rank == 0:
            print('{}\tT_sample {:.3f}\tT_update {:.3f}\tR_avg {:.4f} {}\tR_range ({:.4f}, {:.4f})\teps_len_avg {:.2f}\t'
                  'surr {:+.5f}->{:+.5f}\tvloss {:.4f}->{:.4f}'.format(
                      i_iter, log.sample_time, t_update, log.avg_c_reward,
                      np.array2string(log.avg_c_info, formatter={'all': lambda v: '%.4f' % v}, separator=','),
                      log.min_c_reward, log.max_c_reward, log.num_steps / max(1, log.num_episodes), losses['surr_loss'][0],
                      losses['surr_loss'][-1], losses['value_loss'][0], losses['value_loss'][-1]), flush=True)
    if args.save and rank == 0:
        checkpoint.save_checkpoint(args.save, policy_net, policy_vs_net, value_net, value_vs_net, running_state)
        print('saved', args.save)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    t0 = time.time()
    main()
    print('done in %.1f s' % (time.time() - t0))
