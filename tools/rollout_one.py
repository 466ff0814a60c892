"""One fused rollout at config-2 geometry (for `ncu -k regex:rollout_kernel_t4`):
   python tools/rollout_one.py [--envs 4096] [--horizon 10] [--hidden 300 300] [--reps 2]"""
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
    ap.add_argument('--horizon', type=int, default=10)
    ap.add_argument('--hidden', type=int, nargs=2, default=[300, 300])
    ap.add_argument('--reps', type=int, default=2)
    ap.add_argument('--no-next', action='store_true')
    args = ap.parse_args()
    md = load_builtin()
    T = args.horizon
    L = T + 84
    takes = synthetic_takes(md, 8, L, seed=1)
    cnn = np.concatenate(synthetic_cnn_feat(8, L))
    cfg = Config('subject_03')
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda:0')  # noqa: E731
    model = lib.Model(md, cfg.jkp, cfg.jkd, cfg.a_ref, cfg.a_scale, cfg.torque_lim, getattr(cfg, 'b_diffw', np.ones(md.nbody - 1)),
                      cfg.reward_weights, frame_skip=15, device=0)
    rows, offs, lbs = [], [0], []
    for q in takes:
        r, lb = model.expert_features(q)
        rows.append(r.cpu().numpy()); offs.append(offs[-1] + q.shape[0]); lbs.append(lb)
    model.upload_experts(np.concatenate(rows), np.array(offs), np.array(lbs), cnn)
    w = helpers.policy_weights(model.S + 128, args.hidden[0], args.hidden[1], model.nu, seed=1)
    wd = {k: cu(v.ravel() if k == 'log_std' else v) for k, v in w.items()}
    out = {}
    for r in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.rollout(wd, args.envs, T, episode_len=200, iteration=r, out=out, want_next=not args.no_next)
        e1.record()
        torch.cuda.synchronize()
        print('rollout %d x %d: %.2f ms (%.1f us per env step of the CTA)' % (args.envs, T, e0.elapsed_time(e1), 1e3 * e0.elapsed_time(e1) / T), flush=True)
        L = lib.load()
        if hasattr(L, 'egp_debug_t4_clk'):      # library built with EGP_NVCC_EXTRA=-DEGP_T4_CLK
            import ctypes
            buf = (ctypes.c_ulonglong * 24)()
            L.egp_debug_t4_clk(buf)
            nsub = max(1, buf[4])
            print('  warp 0 cycles per sub-step: backward<PD> %.0f, forward<torque+kin> %.0f, backward<FD> %.0f, forward<accel+Euler> %.0f (%d sub-steps)'
                  % (buf[0] / nsub, buf[1] / nsub, buf[2] / nsub, buf[3] / nsub, nsub), flush=True)
            print('  forward<0> marks per sub-step (cycles since previous mark): ' + ' '.join('%d:%.0f' % (k, buf[8 + k] / nsub) for k in range(16) if buf[8 + k]), flush=True)


if __name__ == '__main__':
    main()
