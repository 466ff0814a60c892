"""Where does the fused rollout's time go?  Times egp_rollout_f64 alone (CUDA events) at config-2 geometry for
several policy widths and sub-step counts: the difference between a 2x300 and a 2x8 policy is the per-step policy
MLP, the slope over frame_skip is one physics sub-step, the intercept is obs / reward / bookkeeping.
usage: python tools/rollout_split.py [--envs 4096] [--horizon 60]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import helpers
    from egopose_b200 import lib
    from egopose_b200.config import Config
    from egopose_b200.mjcf import load_builtin
    from egopose_b200.synthetic import synthetic_cnn_feat, synthetic_takes
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=4096)
    ap.add_argument('--horizon', type=int, default=60)
    ap.add_argument('--reps', type=int, default=3)
    args = ap.parse_args()
    md = load_builtin()
    T = args.horizon
    L = T + 84
    takes = synthetic_takes(md, 8, L, seed=1)
    cnn = np.concatenate(synthetic_cnn_feat(8, L))
    cfg = Config('subject_03')
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda:0')  # noqa: E731
    res = {}
    for fs in (15, 5, 1):
        model = lib.Model(md, cfg.jkp, cfg.jkd, cfg.a_ref, cfg.a_scale, cfg.torque_lim, getattr(cfg, 'b_diffw', np.ones(md.nbody - 1)), cfg.reward_weights,
                          frame_skip=fs, device=0)
        rows, offs, lbs = [], [0], []
        for q in takes:
            r, lb = model.expert_features(q)
            rows.append(r.cpu().numpy()); offs.append(offs[-1] + q.shape[0]); lbs.append(lb)
        model.upload_experts(np.concatenate(rows), np.array(offs), np.array(lbs), cnn)
        for hid in ((300, 300), (32, 16)):
            w = helpers.policy_weights(model.S + 128, hid[0], hid[1], model.nu, seed=1)
            wd = {k: cu(v.ravel() if k == 'log_std' else v) for k, v in w.items()}
            out = {}
            ts = []
            for r in range(args.reps + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                model.rollout(wd, args.envs, T, episode_len=T, iteration=r, out=out)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[(fs, hid)] = min(ts[1:])
            print('frame_skip %2d hidden %s: %.2f ms per rollout, %.1f us per env step' % (fs, hid, res[(fs, hid)], 1e3 * res[(fs, hid)] / T),
                  flush=True)
        model.close()
    mlp = res[(15, (300, 300))] - res[(15, (32, 16))]
    sub = (res[(15, (32, 16))] - res[(1, (32, 16))]) / 14
    rest = res[(1, (32, 16))] - sub
    tot = res[(15, (300, 300))]
    print('split of %.1f ms: policy MLP %.1f (%.0f%%), 15 sub-steps %.1f (%.0f%%), obs/reward/reset/rest %.1f (%.0f%%)'
          % (tot, mlp, 100 * mlp / tot, 15 * sub, 1500 * sub / tot, rest, 100 * rest / tot))


if __name__ == '__main__':
    main()
