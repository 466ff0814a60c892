#!/usr/bin/env python
"""Compile the reference's MJCF asset and yml hyper-parameters into the JSON constants shipped in
egopose_b200/assets (run in the build container, where /root/reference exists).

  python tools/compile_model.py [--ref /root/reference]
"""
import argparse
import json
import os
import sys

import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egopose_b200.mjcf import compile_mjcf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--ref', default='/root/reference')
args = ap.parse_args()
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'egopose_b200', 'assets')
os.makedirs(out_dir, exist_ok=True)

m = compile_mjcf(os.path.join(args.ref, 'assets/mujoco_models/humanoid_1205_v1.xml'))
m.to_json(os.path.join(out_dir, 'humanoid_1205_v1.model.json'))
print('model: nq=%d nv=%d nu=%d nbody=%d mass=%.4f' % (m.nq, m.nv, m.nu, m.nbody, m.total_mass()))

for task in ('egomimic', 'egoforecast'):
    for cfg_id in ('subject_03', 'cross_01'):
        src = os.path.join(args.ref, 'config', task, cfg_id + '.yml')
        cfg = yaml.safe_load(open(src))
        json.dump(cfg, open(os.path.join(out_dir, '%s_%s.cfg.json' % (task, cfg_id)), 'w'), indent=1)
        print('cfg:', task, cfg_id, len(cfg.get('joint_params', [])), 'joints')
