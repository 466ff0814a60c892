"""Host build of the rollout kernel's tree sweeps (egopose_b200/csrc/tree.cuh compiled by g++ into
tests/hostsim/libhost_sweeps.so) vs the C oracle: the block articulated-body arithmetic the CUDA kernel executes is
checked on the CPU, one environment at a time.  Tolerances as in tests/test_gpu_physics.py (north star: 1e-4)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import cphys
import helpers

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'hostsim', 'host_sweeps.cpp')
SO = os.path.join(HERE, 'hostsim', 'libhost_sweeps.so')
HDR = os.path.join(os.path.dirname(HERE), 'egopose_b200', 'csrc', 'tree.cuh')


@pytest.fixture(scope='module')
def hs():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
        subprocess.check_call([cxx, '-O2', '-std=c++17', '-fPIC', '-shared', '-o', SO, SRC])
    from egopose_b200 import lib
    from egopose_b200.mjcf import load_builtin
    L = C.CDLL(SO)
    desc, keep = lib.desc_from_cfg_dict(load_builtin(), helpers.cfg_dict())
    ok = L.hs_init(C.byref(desc))
    assert ok == 1, 'the block sweeps must support the humanoid model (t4_ok)'
    L._keep = (desc, keep)
    return L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_block_sweeps_env_step_vs_oracle(hs, oracle):
    """reset-forward + 15 stable-PD sub-steps from random states: qpos / qvel / first torque / bias / head height"""
    n = 24
    q, v = helpers.rand_states(n, seed=5, vel=0.5)
    act = np.random.RandomState(2).randn(n, 52) * 0.3
    takes = cphys.synthetic_takes(oracle.md, 1, 40, seed=2)
    oracle.make_expert(takes)
    lim = np.array(oracle._keep['torque_lim'])
    for i in range(n):
        qi, vi = q[i].copy(), v[i].copy()
        tq, hz, bias = np.zeros(52), np.zeros(1), np.zeros(58)
        hs.hs_env_step(_p(qi), _p(vi), _p(np.ascontiguousarray(act[i])), _p(tq), _p(hz), _p(bias))
        env = cphys.EoEnv()
        oracle.L.eo_env_set_state(C.byref(oracle.model), C.byref(env), cphys._p(q[i].copy()), cphys._p(v[i].copy()))
        assert helpers.relerr(bias, np.array(env.d.qfrc_bias[:58])) < 1e-11
        ctrl = np.array(oracle._keep['a_ref']) + act[i] * np.array(oracle._keep['a_scale'])
        t0 = np.clip(oracle.compute_torque(env.d, ctrl), -lim, lim)
        assert helpers.relerr(tq, t0) < 1e-8
        env.take = 0
        oracle.cfg.fix_head_lb = -100.0
        oracle.env_step(env, act[i])
        assert helpers.relerr(qi, np.array(env.d.qpos[:59])) < 1e-7
        assert helpers.relerr(vi, np.array(env.d.qvel[:58])) < 1e-6
        assert abs(hz[0] - np.array(env.d.xpos).reshape(-1, 3)[oracle.md['body_names'].index('Head'), 2]) < 1e-8


def test_block_sweeps_vs_reference_golden(hs, golden):
    """first env.step of both golden episodes (reference HumanoidEnv.step / compute_torque on the restated physics)"""
    g = golden('env_traj')
    lim = np.array([jp[5] for jp in helpers.cfg_dict()['joint_params']], dtype=np.float64)
    for ei in range(2):
        qi, vi = g['ep%d.qpos' % ei][0].copy(), g['ep%d.qvel' % ei][0].copy()
        tq, hz = np.zeros(52), np.zeros(1)
        hs.hs_env_step(_p(qi), _p(vi), _p(np.ascontiguousarray(g['ep%d.action' % ei][0])), _p(tq), _p(hz), None)
        assert helpers.relerr(tq, np.clip(g['ep%d.torque0' % ei][0], -lim, lim)) < 1e-7
        assert helpers.relerr(qi, g['ep%d.qpos' % ei][1]) < 1e-7
        assert helpers.relerr(vi, g['ep%d.qvel' % ei][1]) < 1e-6
        assert abs(hz[0] - g['ep%d.head_z' % ei][0]) < 1e-8
