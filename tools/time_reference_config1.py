#!/usr/bin/env python
"""BASELINE config 1 as SURVEY 8d defines it, timed in the BUILD container (needs /root/reference): the UNMODIFIED reference
Python - agents/agent.py Agent.sample (forked sampler workers), ego_pose/envs/humanoid_v1.py, reward_function.py,
ego_pose/core/agent_ego.py AgentEgo.update_params (VideoStateNet BiLSTM context, GAE, 10 PPO epochs) - running on the restated
physics (oracle/mujoco_shim.py in place of mujoco_py: MuJoCo itself is not installable here), CPU only, float64:

    4 sampler workers x 300 steps (min_batch_size 1200, env_episode_len 300), policy / value MLP [300, 300] (BASELINE) or
    [300, 200] (the reference yml), OMP_NUM_THREADS=1 for sampling (README.md:25-27), all cores for the update.

Prints one JSON line; the numbers are recorded in BASELINE.md.  Label: "reference Python + restated physics".
   python tools/time_reference_config1.py [--workers 4] [--steps 300] [--iters 3] [--hidden 300 300]
"""
import argparse
import json
import os
import pickle
import sys
import tempfile
import time

os.environ.setdefault('OMP_NUM_THREADS', '1')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workers', type=int, default=4)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--hidden', type=int, nargs=2, default=[300, 300])
    args = ap.parse_args()
    from oracle import cphys, mujoco_shim, refimport
    refimport.install()
    import torch
    import yaml
    orc = cphys.Oracle()
    mujoco_shim.install(orc)
    work = tempfile.mkdtemp(prefix='egopose_cfg1_')
    os.symlink(os.path.join(refimport.REF, 'config'), os.path.join(work, 'config'))
    os.symlink(os.path.join(refimport.REF, 'assets'), os.path.join(work, 'assets'))
    os.makedirs(os.path.join(work, 'datasets', 'meta'))
    os.makedirs(os.path.join(work, 'datasets', 'features'))
    names = ['take_%d' % i for i in range(8)]
    yaml.safe_dump({'train': names, 'test': names}, open(os.path.join(work, 'datasets', 'meta', 'meta_subject_03.yml'), 'w'))
    os.chdir(work)

    from agents.agent import Agent  # noqa: F401
    from core.critic import Value
    from core.policy_gaussian import PolicyGaussian
    from ego_pose.core.agent_ego import AgentEgo
    from ego_pose.core.reward_function import reward_func
    from ego_pose.envs.humanoid_v1 import HumanoidEnv
    from ego_pose.utils.egomimic_config import Config
    from models.mlp import MLP
    from models.video_state_net import VideoStateNet
    from utils.zfilter import ZFilter

    dtype = torch.float64
    torch.set_default_dtype(dtype)
    cfg = Config('subject_03', create_dirs=False)
    cfg.env_episode_len = args.steps
    cfg.min_batch_size = args.workers * args.steps
    np.random.seed(cfg.seed)
    torch.manual_seed(cfg.seed)
    L = args.steps + 2 * cfg.fr_margin + 64
    takes = cphys.synthetic_takes(orc.md, len(names), L, seed=1)
    # expert features by the restated gen_expert pipeline (same rows the GPU path uploads), CNN features N(0, 1)
    X = cphys.X
    ex = {}
    for n, q in zip(names, takes):
        rows, lb = orc.expert_features(q)
        col = lambda k, w: rows[:, X[k]:X[k] + w].copy()  # noqa: E731
        ex[n] = {'qpos': q, 'qvel': col('QVEL', 58), 'rlinv_local': col('RLINV_LOCAL', 3), 'rangv': col('RANGV', 3),
                 'rq_rmh': col('RQ_RMH', 4), 'ee_pos': col('EE_POS', 15), 'bquat': col('BQUAT', 84), 'bangvel': col('BANGVEL', 63),
                 'len': q.shape[0], 'height_lb': q[:, 2].min(), 'head_height_lb': lb}
    cnn = {n: np.random.RandomState(100 + i).randn(L, 128) for i, n in enumerate(names)}
    pickle.dump(ex, open(cfg.expert_feat_file, 'wb'))
    pickle.dump((cnn, {}), open(cfg.cnn_feat_file, 'wb'))

    env = HumanoidEnv(cfg)
    env.seed(cfg.seed)
    env.load_experts(names, cfg.expert_feat_file, cfg.cnn_feat_file)
    env.set_fix_head_lb(-10.0)              # contact-less bodies fall after ~9 steps; keep every worker on full 300-step episodes
    sd, ad = env.observation_space.shape[0], env.action_space.shape[0]
    running_state = ZFilter((sd,), clip=5)
    pvs = VideoStateNet(128, cfg.policy_v_hdim, cfg.fr_margin, cfg.policy_v_net, cfg.policy_v_net_param, cfg.causal)
    vvs = VideoStateNet(128, cfg.value_v_hdim, cfg.fr_margin, cfg.value_v_net, cfg.value_v_net_param, cfg.causal)
    pol = PolicyGaussian(MLP(sd + cfg.policy_v_hdim, args.hidden, cfg.policy_htype), ad, log_std=cfg.log_std, fix_std=cfg.fix_std)
    val = Value(MLP(sd + cfg.value_v_hdim, args.hidden, cfg.value_htype))
    pparams = list(pol.parameters()) + list(pvs.parameters())
    vparams = list(val.parameters()) + list(vvs.parameters())
    opt_p = torch.optim.Adam(pparams, lr=cfg.policy_lr)
    opt_v = torch.optim.Adam(vparams, lr=cfg.value_lr)
    agent = AgentEgo(env=env, dtype=dtype, device=torch.device('cpu'), running_state=running_state,
                     custom_reward=reward_func[cfg.reward_id], mean_action=False, render=False, num_threads=args.workers,
                     policy_net=pol, policy_vs_net=pvs, value_net=val, value_vs_net=vvs, optimizer_policy=opt_p,
                     optimizer_value=opt_v, opt_num_epochs=cfg.num_optim_epoch, gamma=cfg.gamma, tau=cfg.tau,
                     clip_epsilon=cfg.clip_epsilon, policy_grad_clip=[(pparams, 40)])
    cores = os.cpu_count() or 1
    ts, tu, ns = [], [], []
    for it in range(args.iters + 1):
        torch.set_num_threads(1)
        t0 = time.perf_counter()
        batch, log = agent.sample(cfg.min_batch_size)
        t1 = time.perf_counter()
        torch.set_num_threads(cores)
        agent.update_params(batch)
        t2 = time.perf_counter()
        if it > 0:                      # first iteration = warm-up
            ts.append(t1 - t0); tu.append(t2 - t1); ns.append(log.num_steps)
        print('iter %d: %d steps, T_sample %.2f s, T_update %.2f s' % (it, log.num_steps, t1 - t0, t2 - t1), file=sys.stderr, flush=True)
    n = float(np.mean(ns))
    line = {'config': 'BASELINE config 1: subject_03 egomimic, %d sampler workers x %d steps, MLP %s, VideoStateNet context, 10 PPO epochs, '
                      'float64, CPU' % (args.workers, args.steps, args.hidden),
            'what': 'UNMODIFIED reference Python (agents/agent.py, humanoid_v1.py, reward_function.py, agent_ego.py, agent_ppo.py) on the '
                    'restated physics (oracle/mujoco_shim.py; MuJoCo not installable offline)',
            'host_cores': cores, 'env_steps_per_iteration': n, 't_sample_s': float(np.mean(ts)), 't_update_s': float(np.mean(tu)),
            'env_steps_per_s': n / float(np.mean(ts) + np.mean(tu)), 'rollout_env_steps_per_s': n / float(np.mean(ts)),
            'update_samples_per_s': n / float(np.mean(tu)), 'iters_timed': args.iters}
    print(json.dumps(line))


if __name__ == '__main__':
    main()
