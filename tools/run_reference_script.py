#!/usr/bin/env python
"""Run the reference's UNMODIFIED training script (ego_pose/ego_mimic.py or ego_forecast.py) on the fused path:

   python tools/run_reference_script.py /path/to/ego_mimic.py [--iters 2] [--envs-batch 20000] [--workdir DIR] [-- script args]

What it does: creates a scratch working directory with a synthetic dataset (egopose_b200.dataset) and a config yml derived
from the committed constants of config/<task>/subject_03.yml (``max_iter_num`` / ``min_batch_size`` / ``save_model_interval``
overridden so that a short run also exercises the checkpoint branch), puts egopose_b200/compat first on sys.path so the
script's own imports (`from utils import *`, `from core.policy_gaussian import ...`, `from ego_pose.core.agent_ego import
AgentEgo`, ...) resolve to the B200-native implementations, and executes the script file as __main__ with cwd = the scratch
directory.  The script source is never copied or edited.
"""
import argparse
import json
import os
import runpy
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def prepare(workdir, task, cfg_id, iters, batch, ckpt_every, episode_len=None, write_data=True):
    sys.path.insert(0, ROOT)
    import yaml
    from egopose_b200.config import Config
    src = json.load(open(os.path.join(ROOT, 'egopose_b200', 'assets', '%s_subject_03.cfg.json' % task)))
    src.update(max_iter_num=iters, min_batch_size=batch, save_model_interval=ckpt_every)
    if task == 'egoforecast':       # warm start from the ego-mimic checkpoint of the same scratch directory (ego_forecast.py:60-69)
        src.update(ego_mimic_cfg=cfg_id, ego_mimic_iter=iters)
    if episode_len:
        src['env_episode_len'] = episode_len
    os.makedirs(os.path.join(workdir, 'config', task), exist_ok=True)
    yaml.safe_dump(src, open(os.path.join(workdir, 'config', task, cfg_id + '.yml'), 'w'))
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        cfg = Config(cfg_id, task=task)
        if write_data:
            from egopose_b200.dataset import write_synthetic_dataset
            write_synthetic_dataset(workdir, cfg)
        else:       # CPU-only preparation (tests): the take lists are enough to reach the first CUDA requirement
            from egopose_b200.dataset import write_meta_yml
            write_meta_yml(os.path.join(workdir, 'datasets', 'meta', '%s.yml' % cfg.meta_id), ['synth_00'], ['synth_01'])
    finally:
        os.chdir(cwd)
    return cfg


def run(script, workdir, cfg_id, extra_args):
    compat = os.path.join(ROOT, 'egopose_b200', 'compat')
    for p in (ROOT, compat):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [compat, ROOT]
    for m in [k for k in sys.modules if k.split('.')[0] in ('utils', 'core', 'models', 'agents', 'envs', 'ego_pose')]:
        del sys.modules[m]
    os.chdir(workdir)
    sys.argv = [script, '--cfg', cfg_id] + list(extra_args)
    runpy.run_path(script, run_name='__main__')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('script')
    ap.add_argument('--iters', type=int, default=2)
    ap.add_argument('--batch', type=int, default=20000, help='cfg.min_batch_size')
    ap.add_argument('--episode-len', type=int, default=None)
    ap.add_argument('--workdir', default=None)
    ap.add_argument('rest', nargs='*')
    args = ap.parse_args()
    task = 'egoforecast' if 'forecast' in os.path.basename(args.script) else 'egomimic'
    workdir = args.workdir or tempfile.mkdtemp(prefix='egp_dropin_')
    cfg_id = 'dropin_01'
    prepare(workdir, task, cfg_id, args.iters, args.batch, ckpt_every=1, episode_len=args.episode_len)
    print('[run_reference_script] %s --cfg %s in %s' % (args.script, cfg_id, workdir), flush=True)
    run(os.path.abspath(args.script), workdir, cfg_id, args.rest)
    for dp, _, files in os.walk(os.path.join(workdir, 'results')):
        for f in files:
            print('[run_reference_script] wrote', os.path.relpath(os.path.join(dp, f), workdir))


if __name__ == '__main__':
    main()
