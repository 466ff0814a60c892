#!/usr/bin/env python
"""Evaluation script with the command line of ego_pose/ego_mimic_eval.py (--cfg --iter --data --fail-safe) on the
B200-native path: loads a checkpoint written by ego_mimic.py / examples/train_egomimic.py, rolls every take of the
chosen split out in the fused kernel (egopose_b200.evaluate.eval_takes) and writes the (results, meta) pickle that
ego_pose/eval_pose.py:31 reads, under the reference's file name (ego_mimic_eval.py:189-192).

The state-regression net that supplies the fail-safe states in the reference (models/video_reg_net.py) is outside the
hot path: pass its per-take predictions with --state-pred (pickle: {take: [len - 2 fr_margin, 115]}), else the
expert's own observations are used.  --synthetic runs on seeded synthetic takes (the EgoPose dataset is not
redistributable).

  python examples/eval_egomimic.py --synthetic --fail-safe valuefs --out /tmp/res
  python examples/eval_egomimic.py --cfg subject_03 --iter 3000 --data test        (inside an EgoPose checkout with data)
"""
import argparse
import os
import pickle
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from egopose_b200 import checkpoint, evaluate, metrics  # noqa: E402
from egopose_b200.config import Config  # noqa: E402
from egopose_b200.env import HumanoidEnv  # noqa: E402
from egopose_b200.nets import MLP, FrameContext, PolicyGaussian, Value, VideoStateNet  # noqa: E402
from egopose_b200.synthetic import synthetic_cnn_feat, synthetic_takes  # noqa: E402
from egopose_b200.zfilter import ZFilter  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default='subject_03')
    ap.add_argument('--iter', type=int, default=0)
    ap.add_argument('--data', default='test')
    ap.add_argument('--fail-safe', default='valuefs', choices=['valuefs', 'naivefs', 'none'])
    ap.add_argument('--show-noise', action='store_true')
    ap.add_argument('--state-pred', default='', help='pickle {take: [len - 2 fr_margin, obs_dim]} of state-net predictions')
    ap.add_argument('--checkpoint', default='', help='default: <model_dir>/iter_%%04d.p when it exists')
    ap.add_argument('--identity-video-net', action='store_true', help='FrameContext (raw CNN features) instead of VideoStateNet')
    ap.add_argument('--synthetic', action='store_true')
    ap.add_argument('--takes', type=int, default=3)
    ap.add_argument('--len', type=int, default=80)
    ap.add_argument('--out', default='', help='output directory (default cfg.result_dir)')
    ap.add_argument('--device', type=int, default=0)
    args = ap.parse_args(argv)
    torch.cuda.set_device(args.device)
    device = torch.device('cuda', args.device)
    torch.set_default_dtype(torch.float64)
    cfg = Config(args.cfg)
    np.random.seed(cfg.seed)
    torch.manual_seed(cfg.seed)

    env = HumanoidEnv(cfg, device=args.device)
    if args.synthetic:
        names = ['take_%d' % i for i in range(args.takes)]
        lens = [args.len + 7 * i for i in range(args.takes)]                 # ragged takes
        takes = [synthetic_takes(env.md, 1, L, seed=11 + i)[0] for i, L in enumerate(lens)]
        cnn = [synthetic_cnn_feat(1, L, seed=31 + i)[0] for i, L in enumerate(lens)]
        env.set_expert_qpos(names, takes, cnn)
    else:
        env.load_experts(cfg.takes[args.data], cfg.expert_feat_file, cfg.cnn_feat_file)      # ego_mimic_eval.py:41
    feat_dim = env.cnn_feat[0].shape[-1]
    state_dim, action_dim = env.observation_space.shape[0], env.action_space.shape[0]
    if args.identity_video_net or args.synthetic:
        policy_vs_net, value_vs_net, hp, hv = FrameContext(feat_dim), FrameContext(feat_dim), feat_dim, feat_dim
    else:
        policy_vs_net = VideoStateNet(feat_dim, cfg.policy_v_hdim, cfg.fr_margin, cfg.policy_v_net, cfg.policy_v_net_param, cfg.causal).to(device)
        value_vs_net = VideoStateNet(feat_dim, cfg.value_v_hdim, cfg.fr_margin, cfg.value_v_net, cfg.value_v_net_param, cfg.causal).to(device)
        hp, hv = cfg.policy_v_hdim, cfg.value_v_hdim
    policy_net = PolicyGaussian(MLP(state_dim + hp, cfg.policy_hsize, cfg.policy_htype), action_dim,
                                log_std=cfg.log_std, fix_std=cfg.fix_std).to(device)
    value_net = Value(MLP(state_dim + hv, cfg.value_hsize, cfg.value_htype)).to(device)
    running_state = ZFilter((state_dim,), clip=5)
    cp_path = args.checkpoint or '%s/iter_%04d.p' % (cfg.model_dir, args.iter)
    if os.path.exists(cp_path):
        vs = isinstance(policy_vs_net, VideoStateNet)
        _, running_state = checkpoint.load_checkpoint(cp_path, policy_net, policy_vs_net if vs else None, value_net,
                                                      value_vs_net if vs else None)
        print('loaded', cp_path)
    else:
        print('no checkpoint at %s: randomly initialised nets' % cp_path)
        running_state = None
    state_pred = None
    if args.state_pred:
        fm = cfg.fr_margin
        sp = pickle.load(open(args.state_pred, 'rb'))
        table = evaluate.expert_obs_table(env.kernel)
        off = np.asarray(env.kernel.take_off)
        state_pred = []
        for k, take in enumerate(env.expert_list):                            # predictions cover frames fm .. len - fm
            t = table[off[k]:off[k + 1]].copy()
            t[fm:fm + sp[take].shape[0]] = sp[take]
            state_pred.append(t)
    results, meta, info = evaluate.eval_takes(env, policy_net, policy_vs_net, running_state, state_pred=state_pred,
                                              fail_safe=args.fail_safe, show_noise=args.show_noise, value_net=value_net,
                                              value_vs_net=value_vs_net if isinstance(value_vs_net, VideoStateNet) else None)
    out_dir = args.out or cfg.result_dir
    os.makedirs(out_dir, exist_ok=True)
    fs_tag = '' if args.fail_safe == 'valuefs' else '_' + args.fail_safe                     # ego_mimic_eval.py:189-191
    res_path = '%s/iter_%04d_%s%s.p' % (out_dir, args.iter, args.data, fs_tag)
    evaluate.save_results(results, meta, res_path)
    for take in env.expert_list:
        d = np.linalg.norm(results['traj_pred'][take][:, 7:] - results['traj_orig'][take][:, 7:], axis=1).mean()
        print('%-12s frames %4d  mean joint-angle distance to the expert %.4f  mean reward %.4f'
              % (take, results['traj_pred'][take].shape[0], d, info['rewards'][take].mean()))
    m = metrics.compute_metrics(results)                                                    # eval_pose.py:31-66 ('stats' mode)
    print('all - pose dist: %.4f, vel dist: %.4f, accels: %.4f' % (m['pose_dist'], m['vel_dist'], m['smoothness']))
    print('num reset: %d' % meta['num_reset'])
    print('saved results to %s' % res_path)
    env.close()
    return res_path


if __name__ == '__main__':
    main()
