"""ctypes binding of libegopose_b200.so (include/egopose_b200.h).

There is NO CPU fallback: importing this module is cheap, but every compute entry point requires the
CUDA library built in-tree (python -m egopose_b200.build) and a CUDA device; failures raise EgpError.
PyTorch is used only to own device memory / streams (tensor.data_ptr()).
"""
import ctypes as C
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libegopose_b200.so')

X = dict(QPOS=0, QVEL=59, RLINV_LOCAL=117, RANGV=120, RQ_RMH=123, EE_POS=127, BQUAT=142, BANGVEL=226, STRIDE=292)
NEE = 5
LOG = dict(NUM_STEPS=0, NUM_EPISODES=1, TOTAL_REWARD=2, TOTAL_C_REWARD=3, MIN_C_REWARD=4, MAX_C_REWARD=5, C_INFO=6,
           MIN_EPISODE_REWARD=11, MAX_EPISODE_REWARD=12, NUM_NAN_RESETS=13, NUM_FAILSAFE_RESETS=14, SIZE=16)
EE_NAMES = ['LeftFoot', 'RightFoot', 'LeftHand', 'RightHand', 'Head']       # humanoid_v1.py:100

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p


class EgpError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [('nq', C.c_int32), ('nv', C.c_int32), ('nu', C.c_int32), ('nbody', C.c_int32),
                ('timestep', C.c_double), ('gravity', C.c_double * 3),
                ('body_parent', _ip), ('body_dofadr', _ip), ('body_dofnum', _ip), ('body_qposadr', _ip),
                ('body_pos', _dp), ('body_mass', _dp), ('body_ipos', _dp), ('body_inertia', _dp),
                ('dof_body', _ip), ('dof_parent', _ip),
                ('dof_armature', _dp), ('dof_axis', _dp), ('dof_anchor', _dp),
                ('ee_body', C.c_int32 * NEE), ('head_body', C.c_int32), ('frame_skip', C.c_int32),
                ('jkp', _dp), ('jkd', _dp), ('a_ref', _dp), ('a_scale', _dp), ('torque_lim', _dp), ('b_diffw', _dp),
                ('w_p', C.c_double), ('w_v', C.c_double), ('w_e', C.c_double), ('w_rp', C.c_double),
                ('w_rv', C.c_double), ('k_p', C.c_double), ('k_v', C.c_double), ('k_e', C.c_double),
                ('k_rh', C.c_double), ('k_rq', C.c_double), ('k_rl', C.c_double), ('k_ra', C.c_double),
                ('v_ord', C.c_int32), ('decay', C.c_int32)]


class RolloutCfg(C.Structure):
    _fields_ = [('n_env', C.c_int32), ('horizon', C.c_int32), ('episode_len', C.c_int32), ('fr_margin', C.c_int32),
                ('end_reward', C.c_double), ('fix_head_lb', C.c_double), ('noise_rate', C.c_double),
                ('mean_action', C.c_int32), ('zf_clip', C.c_double), ('seed', C.c_uint64), ('iteration', C.c_uint64),
                ('max_resets', C.c_int32), ('eval_mode', C.c_int32)]


class PolicyWeights(C.Structure):
    _fields_ = [('in_dim', C.c_int32), ('h1', C.c_int32), ('h2', C.c_int32), ('out_dim', C.c_int32),
                ('d_W1', _vp), ('d_b1', _vp), ('d_W2', _vp), ('d_b2', _vp), ('d_W3', _vp), ('d_b3', _vp),
                ('d_log_std', _vp)]


class RolloutIn(C.Structure):
    _fields_ = [('d_eps', _vp), ('d_reset_take', _vp), ('d_reset_start', _vp), ('d_mean_flag', _vp),
                ('d_zf_mean', _vp), ('d_zf_std', _vp), ('d_ctx', _vp), ('d_win_off', _vp), ('ctx_dim', C.c_int32),
                ('ctx_mode', C.c_int32), ('ctx_T', C.c_int32), ('d_snet_W', _vp), ('d_snet_b', _vp), ('d_snet_state', _vp),
                ('snet_hdim', C.c_int32), ('d_fix_len', _vp), ('d_state_pred', _vp), ('value_net', _vp), ('d_vctx', _vp), ('d_value_stat', _vp),
                ('d_init_qpos', _vp), ('d_init_qvel', _vp)]


class TrajOut(C.Structure):
    _fields_ = [('d_states', _vp), ('d_actions', _vp), ('d_masks', _vp), ('d_next_states', _vp), ('d_rewards', _vp),
                ('d_exps', _vp), ('d_v_metas', _vp), ('d_c_info', _vp), ('d_raw_obs', _vp), ('d_final_qpos', _vp),
                ('d_final_qvel', _vp), ('d_logger', _vp), ('d_values', _vp), ('d_qpos_traj', _vp), ('d_qvel_traj', _vp)]


class MlpNet(C.Structure):
    _fields_ = [('in_dim', C.c_int32), ('h1', C.c_int32), ('h2', C.c_int32), ('out_dim', C.c_int32),
                ('d_W1', _vp), ('d_b1', _vp), ('d_W2', _vp), ('d_b2', _vp), ('d_W3', _vp), ('d_b3', _vp),
                ('d_gW1', _vp), ('d_gb1', _vp), ('d_gW2', _vp), ('d_gb2', _vp), ('d_gW3', _vp), ('d_gb3', _vp),
                ('d_dx', _vp), ('dx_cols', C.c_int32)]


class MlpLoss(C.Structure):
    _fields_ = [('kind', C.c_int32), ('d_actions', _vp), ('d_log_std', _vp), ('d_adv', _vp), ('d_stats', _vp),
                ('d_logp0', _vp), ('d_exps', _vp), ('clip_eps', C.c_double), ('inv_count', C.c_double),
                ('d_dlogstd', _vp), ('d_returns', _vp), ('inv_n', C.c_double), ('d_loss', _vp), ('init_logp0', C.c_int32)]


# every symbol include/egopose_b200.h declares: name -> (restype, argtypes)
_i64 = C.c_int64
_d = C.c_double
_int = C.c_int
SYMBOLS = {
    'egp_last_error_string': (C.c_char_p, []),
    'egp_version': (_int, []),
    'egp_model_create': (_int, [C.POINTER(ModelDesc), _int, C.POINTER(_vp)]),
    'egp_model_destroy': (None, [_vp]),
    'egp_expert_upload': (_int, [_vp, _int, _ip, _dp, _dp, _dp, _int]),
    'egp_expert_features_f64': (_int, [_vp, _int, _vp, _vp, _vp, _vp]),
    'egp_expert_features_ex_f64': (_int, [_vp, _int, _vp, _vp, _vp, _vp, _vp]),
    'egp_forward_debug_f64': (_int, [_vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'egp_env_step_debug_f64': (_int, [_vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'egp_rollout_f64': (_int, [_vp, C.POINTER(PolicyWeights), C.POINTER(RolloutCfg), C.POINTER(RolloutIn),
                               C.POINTER(TrajOut), _vp]),
    'egp_model_set_joint_limits': (_int, [_vp, _vp, _vp, _vp, _vp]),
    'egp_cons_cap_hits': (_i64, [_int]),
    'egp_cons_passes': (_int, [_vp, _int]),
    'egp_model_set_contacts': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _vp, _vp]),
    'egp_gae_work_bytes': (_i64, [_i64]),
    'egp_gae_f64': (_int, [_vp, _vp, _vp, _d, _d, _i64, _vp, _vp, _vp, _vp, _vp]),
    'egp_gae_set_onepass_min': (_i64, [_i64]),
    'egp_oz_mlp_set_fused_slicing': (_int, [_int]),
    'egp_standardize_f64': (_int, [_vp, _i64, _vp, _vp]),
    'egp_gauss_logp_f64': (_int, [_vp, _vp, _vp, _i64, _int, _vp, _vp]),
    'egp_ppo_loss_grad_f64': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _i64, _int, _vp, _vp, _vp, _vp]),
    'egp_value_loss_grad_f64': (_int, [_vp, _vp, _d, _i64, _vp, _vp, _vp]),
    'egp_bias_relu_f64': (_int, [_vp, _vp, _i64, _int, _vp]),
    'egp_relu_bwd_f64': (_int, [_vp, _vp, _i64, _int, _vp]),
    'egp_gather_rows_f64': (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    'egp_relu_bwd_colsum_f64': (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    'egp_colsum_f64': (_int, [_vp, _i64, _int, _vp, _vp]),
    'egp_col_moments_f64': (_int, [_vp, _i64, _int, _vp, _vp, _vp]),
    'egp_build_input_f64': (_int, [_vp, _vp, _vp, _vp, _i64, _int, _vp, _vp]),
    'egp_sumsq_f64': (_int, [_vp, _i64, _vp, _vp]),
    'egp_adam_step_f64': (_int, [_vp, _vp, _vp, _vp, _i64, _d, _d, _d, _d, _i64, _d, _vp, _vp]),
    'egp_oz_radix_bits': (_int, []),
    'egp_oz_slice_rows_f64': (_int, [_vp, _i64, _int, _i64, _int, _vp, _int, _vp, _vp, _vp]),
    'egp_oz_colmax_f64': (_int, [_vp, _i64, _int, _i64, _vp, _vp]),
    'egp_oz_slice_cols_t_f64': (_int, [_vp, _i64, _int, _i64, _int, _vp, _vp, _i64, _vp, _int, _vp]),
    'egp_oz_gemm_work_bytes': (_i64, [_i64, _int, _i64, _int]),
    'egp_oz_gemm_f64': (_int, [_vp, _vp, _i64, _vp, _vp, _int, _i64, _int, _vp, _int, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    'egp_oz_gemm_max_f64': (_int, [_vp, _vp, _i64, _vp, _vp, _int, _i64, _int, _vp, _int, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _vp]),
    'egp_oz_slice_both_f64': (_int, [_vp, _i64, _int, _i64, _int, _vp, _vp, _vp, _int, _vp, _vp, _i64, _vp, _int, _vp]),
    'egp_oz_mlp_chunk_rows': (_i64, []),
    'egp_oz_mlp_work_bytes': (_i64, [_int, _int, _int, _int, _i64, _int]),
    'egp_oz_mlp_xcache_bytes': (_i64, [_int, _i64, _i64, _int]),
    'egp_lstm_wfrag_elems': (_i64, [_int]),
    'egp_lstm_pack_whh_f64': (_int, [_vp, _int, _vp, _vp, _vp]),
    'egp_lstm_seq_fwd_f64': (_int, [_vp, _vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp]),
    'egp_lstm_seq_bwd_f64': (_int, [_vp, _vp, _vp, _vp, _int, _i64, _int, _vp, _vp, _vp]),
    'egp_comm_handle_bytes': (_i64, []),
    'egp_comm_create': (_int, [_int, _int, _int, _i64, _vp, _vp]),
    'egp_comm_connect': (_int, [_vp, _vp]),
    'egp_comm_connect_local': (_int, [_vp, _vp]),
    'egp_comm_set_timeout_cycles': (_i64, [_i64]),
    'egp_comm_src': (_vp, [_vp]),
    'egp_comm_out': (_vp, [_vp]),
    'egp_allreduce_grads_f64': (_int, [_vp, _i64, _vp]),
    'egp_comm_error': (_int, [_vp]),
    'egp_comm_destroy': (None, [_vp]),
    'egp_oz_mlp_step_f64': (_int, [C.POINTER(MlpNet), _vp, _i64, _i64, C.POINTER(MlpLoss), _vp, _int, _i64, _vp, _int, _vp, _i64,
                                   _vp]),
}

_lib = None
launches = 0        # kernels launched through this binding (bench.py's gpu_launches)


def load():
    """Load the C-ABI library; raises EgpError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EgpError('libegopose_b200.so is missing - run `python -m egopose_b200.build` '
                           '(there is no CPU fallback)')
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what=''):
    if rc != 0:
        raise EgpError('%s failed (%d): %s' % (what, rc, load().egp_last_error_string().decode()))


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    if t is None:
        return None
    if not t.is_cuda:
        raise EgpError('expected a CUDA tensor (no CPU fallback)')
    if not t.is_contiguous():
        raise EgpError('expected a contiguous tensor')
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _np_d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _np_i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def reward_weights(ws):
    """defaults of ego_pose/core/reward_function.py:8-12"""
    ws = ws or {}
    out = {}
    for name, dv in (('w_p', 0.5), ('w_v', 0.1), ('w_e', 0.2), ('w_rp', 0.1), ('w_rv', 0.1), ('k_p', 2), ('k_v', 0.005),
                     ('k_e', 20), ('k_rh', 300), ('k_rq', 300), ('k_rl', 5.0), ('k_ra', 0.5)):
        out[name] = float(ws.get(name, dv))
    out['v_ord'] = int(ws.get('v_ord', 2))
    out['decay'] = int(bool(ws.get('decay', False)))
    return out


def build_desc(md, jkp, jkd, a_ref, a_scale, torque_lim, b_diffw, reward_ws=None, frame_skip=15):
    """EgpModelDesc (include/egopose_b200.h) of a compiled MJCF model + cfg constants; returns (desc, arrays kept alive)"""
    keep = {}
    d = ModelDesc()
    d.nq, d.nv, d.nu, d.nbody, d.timestep = md.nq, md.nv, md.nu, md.nbody, md.timestep
    d.gravity[:] = md.gravity
    for name in ('body_parent', 'body_dofadr', 'body_dofnum', 'body_qposadr', 'dof_body', 'dof_parent'):
        keep[name] = _np_i(getattr(md, name))
        setattr(d, name, keep[name].ctypes.data_as(_ip))
    for name in ('body_pos', 'body_mass', 'body_ipos', 'body_inertia', 'dof_armature', 'dof_axis', 'dof_anchor'):
        keep[name] = _np_d(getattr(md, name))
        setattr(d, name, keep[name].ctypes.data_as(_dp))
    d.ee_body[:] = [md.body_names.index(n) for n in EE_NAMES]
    d.head_body = md.body_names.index('Head')
    d.frame_skip = frame_skip
    for name, val in (('jkp', jkp), ('jkd', jkd), ('a_ref', a_ref), ('a_scale', a_scale), ('torque_lim', torque_lim),
                      ('b_diffw', b_diffw)):
        keep[name] = _np_d(val)
        setattr(d, name, keep[name].ctypes.data_as(_dp))
    for k, v in reward_weights(reward_ws).items():
        setattr(d, k, v)
    return d, keep


def desc_from_cfg_dict(md, cfg, frame_skip=15):
    """build_desc from the yml dict of config/egomimic/*.yml (egomimic_config.py:105-122 semantics)"""
    jp = list(zip(*cfg['joint_params']))
    mult = cfg.get('jkp_multiplier', 1.0)
    jkp = np.array(jp[1], dtype=np.float64) * mult
    jkd = np.array(jp[2], dtype=np.float64) * cfg.get('jkd_multiplier', mult)
    a_ref = np.deg2rad(np.array(jp[3], dtype=np.float64))
    b_diffw = np.array(list(zip(*cfg['body_params']))[1], dtype=np.float64)
    return build_desc(md, jkp, jkd, a_ref, np.array(jp[4], dtype=np.float64), np.array(jp[5], dtype=np.float64), b_diffw,
                      cfg.get('reward_weights'), frame_skip)


class Model:
    """Owns an EgpModel handle: compiled MJCF constants + cfg constants + expert tables on one device."""

    def __init__(self, md, jkp, jkd, a_ref, a_scale, torque_lim, b_diffw, reward_ws=None, frame_skip=15, device=0):
        import torch
        if not torch.cuda.is_available():
            raise EgpError('a CUDA device is required (no CPU fallback)')
        self.lib = load()
        self.md = md
        self.device = int(device)
        self.nq, self.nv, self.nu, self.nbody = md.nq, md.nv, md.nu, md.nbody
        self.S = self.nq - 2 + self.nv
        self.dt = md.timestep * frame_skip
        d, keep = build_desc(md, jkp, jkd, a_ref, a_scale, torque_lim, b_diffw, reward_ws, frame_skip)
        self._keep = keep
        self.desc = d
        h = _vp()
        torch.cuda.set_device(self.device)
        check(self.lib.egp_model_create(C.byref(d), self.device, C.byref(h)), 'egp_model_create')
        self.handle = h
        self.n_takes = 0
        self.ctx_dim = 0
        self.take_off = None
        self.head_lb = None

    @classmethod
    def from_cfg_dict(cls, md, cfg, device=0):
        """cfg: the yml dict of config/egomimic/*.yml (egomimic_config.py:105-122 semantics)"""
        jp = list(zip(*cfg['joint_params']))
        mult = cfg.get('jkp_multiplier', 1.0)
        jkp = np.array(jp[1], dtype=np.float64) * mult
        jkd = np.array(jp[2], dtype=np.float64) * cfg.get('jkd_multiplier', mult)
        a_ref = np.deg2rad(np.array(jp[3], dtype=np.float64))
        b_diffw = np.array(list(zip(*cfg['body_params']))[1], dtype=np.float64)
        return cls(md, jkp, jkd, a_ref, np.array(jp[4], dtype=np.float64), np.array(jp[5], dtype=np.float64), b_diffw,
                   cfg.get('reward_weights'), device=device)

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.egp_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- expert tables ---------------------------------------------------------------------
    def upload_experts(self, rows, take_off, head_lb, ctx=None):
        rows, take_off, head_lb = _np_d(rows), _np_i(take_off), _np_d(head_lb)
        if rows.shape != (take_off[-1], X['STRIDE']):
            raise EgpError('expert rows must be [total_frames, %d]' % X['STRIDE'])
        cp, cd = None, 0
        if ctx is not None:
            ctx = _np_d(ctx)
            cp, cd = ctx.ctypes.data_as(_dp), ctx.shape[1]
        check(self.lib.egp_expert_upload(self.handle, len(head_lb), take_off.ctypes.data_as(_ip),
                                         rows.ctypes.data_as(_dp), head_lb.ctypes.data_as(_dp), cp, cd),
              'egp_expert_upload')
        self.n_takes, self.ctx_dim, self.take_off, self.head_lb = len(head_lb), cd, take_off, head_lb
        self.rows_host = rows           # host copy of the packed expert rows (evaluate.expert_obs_table)

    def expert_features(self, qpos, extras=False):
        """gen_expert.get_expert for one take on the GPU: qpos [L, nq] -> (rows [L, 292] tensor, head_height_lb);
        extras=True also returns [L, 21] = head_pos 3 | com 3 | ee_wpos 15 (the dict keys that are not rollout inputs)"""
        global launches
        import torch
        q = torch.as_tensor(np.ascontiguousarray(qpos), dtype=torch.float64, device='cuda:%d' % self.device)
        L = q.shape[0]
        rows = torch.empty((L, X['STRIDE']), dtype=torch.float64, device=q.device)
        hz = torch.empty(L, dtype=torch.float64, device=q.device)
        ex = torch.empty((L, 21), dtype=torch.float64, device=q.device) if extras else None
        check(self.lib.egp_expert_features_ex_f64(self.handle, L, ptr(q), ptr(rows), ptr(hz), ptr(ex), stream_ptr()),
              'egp_expert_features_ex_f64')
        launches += 1
        if extras:
            return rows, float(hz.min().item()), ex
        return rows, float(hz.min().item())

    # ---- joint limits -------------------------------------------------------------------------
    def dof_ranges(self):
        """[nv][2] radians: the XML's hinge ranges per dof (free root: 0, 0 = none)"""
        md = self.md
        rng = np.zeros((self.nv, 2))
        free = md.body_dofnum[0] == 6
        for j, r in enumerate(md.jnt_range):
            if free and j == 0:
                continue
            rng[j + 5 if free else j] = r
        return rng

    def invweight0(self):
        """mjModel.dof_invweight0: diag(M^-1) at qpos0 (mjcf.inverse_weights)"""
        from .mjcf import inverse_weights
        return inverse_weights(self.md)[0]

    def set_joint_limits(self, on=True, solref=None, solimp=None):
        """sim.step() with the XML's joint ranges as MuJoCo soft constraints (include/egopose_b200.h:
        egp_model_set_joint_limits); off = smooth dynamics, the default"""
        if not on:
            check(self.lib.egp_model_set_joint_limits(self.handle, None, None, None, None), 'egp_model_set_joint_limits')
            self.joint_limits = False
            return
        keep = [np.ascontiguousarray(self.dof_ranges()), np.ascontiguousarray(self.invweight0())]
        sr = np.ascontiguousarray(solref, dtype=np.float64) if solref is not None else None
        si = np.ascontiguousarray(solimp, dtype=np.float64) if solimp is not None else None
        hp = lambda a: a.ctypes.data if a is not None else None      # noqa: E731
        check(self.lib.egp_model_set_joint_limits(self.handle, hp(keep[0]), hp(keep[1]), hp(sr), hp(si)),
              'egp_model_set_joint_limits')
        self.joint_limits = True

    def set_contacts(self, on=True, margin=0.001, friction=1.0, solref=None, solimp=None):
        """sim.step() with the floor (plane z = 0) against the body geoms, MuJoCo's soft contacts with pyramidal cones
        (include/egopose_b200.h: egp_model_set_contacts); off = smooth dynamics, the default.  margin / friction default
        to the XML's values (geom margin 0.001, floor friction 1)"""
        if not on:
            check(self.lib.egp_model_set_contacts(self.handle, None, None, None, None, None, 0.0, 1.0, None, None),
                  'egp_model_set_contacts')
            self.contacts = False
            return
        from .mjcf import inverse_weights
        md = self.md
        if not md.geom_type:
            raise EgpError('the model description carries no collision geoms (recompile it with tools/compile_model.py)')
        keep = [_np_i(md.geom_type), _np_d(md.geom_size), _np_d(md.geom_p0), _np_d(md.geom_p1),
                np.ascontiguousarray(inverse_weights(md)[1])]
        sr = np.ascontiguousarray(solref, dtype=np.float64) if solref is not None else None
        si = np.ascontiguousarray(solimp, dtype=np.float64) if solimp is not None else None
        hp = lambda a: a.ctypes.data if a is not None else None      # noqa: E731
        check(self.lib.egp_model_set_contacts(self.handle, hp(keep[0]), hp(keep[1]), hp(keep[2]), hp(keep[3]), hp(keep[4]),
                                              float(margin), float(friction), hp(sr), hp(si)), 'egp_model_set_contacts')
        self.contacts = True

    # ---- debug / parity -----------------------------------------------------------------------
    def forward_debug(self, qpos, qvel, ctrl=None):
        global launches
        import torch
        n = qpos.shape[0]
        bias = torch.empty((n, self.nv), dtype=torch.float64, device=qpos.device)
        xpos = torch.empty((n, self.nbody, 3), dtype=torch.float64, device=qpos.device)
        qacc = torch.empty((n, self.nv), dtype=torch.float64, device=qpos.device)
        check(self.lib.egp_forward_debug_f64(self.handle, n, ptr(qpos), ptr(qvel), ptr(ctrl), ptr(bias), ptr(xpos),
                                             ptr(qacc), stream_ptr()), 'egp_forward_debug_f64')
        launches += 1
        return bias, xpos, qacc

    def env_step_debug(self, qpos, qvel, action):
        """in-place env.step on freshly reset states; returns (obs, head_z, torque of sub-step 0)"""
        global launches
        import torch
        n = qpos.shape[0]
        obs = torch.empty((n, self.S), dtype=torch.float64, device=qpos.device)
        hz = torch.empty(n, dtype=torch.float64, device=qpos.device)
        tq = torch.empty((n, self.nu), dtype=torch.float64, device=qpos.device)
        check(self.lib.egp_env_step_debug_f64(self.handle, n, ptr(qpos), ptr(qvel), ptr(action), ptr(obs), ptr(hz),
                                              ptr(tq), stream_ptr()), 'egp_env_step_debug_f64')
        launches += 1
        return obs, hz, tq

    # ---- rollout ------------------------------------------------------------------------------
    def rollout(self, weights, n_env, horizon, episode_len, fr_margin=10, end_reward=0.0, fix_head_lb=None,
                noise_rate=1.0, mean_action=False, zf_mean=None, zf_std=None, zf_clip=5.0, seed=1, iteration=0,
                eps=None, reset_take=None, reset_start=None, mean_flag=None, want_next=True, want_raw=True, out=None,
                ctx=None, win_off=None, ctx_const=False, snet=None, eval_mode=False, fix_len=None, state_pred=None,
                want_traj=False, init_qpos=None, init_qvel=None, value_weights=None, vctx=None, value_stat=None):
        """weights: dict with W1,b1,W2,b2,W3,b3,log_std CUDA float64 tensors (torch [out,in] layout).
        Returns a dict of CUDA tensors in TrajBatchEgo layout (+ logger, c_info, raw_obs, final state)."""
        global launches
        import torch
        dev = weights['W1'].device
        N, S, nu = n_env * horizon, self.S, self.nu
        f64 = dict(dtype=torch.float64, device=dev)
        if out is None:
            out = {}
        def buf(name, shape, dtype=torch.float64):
            t = out.get(name)
            if t is None or tuple(t.shape) != tuple(shape):
                t = torch.empty(shape, dtype=dtype, device=dev)
                out[name] = t
            return t
        o = TrajOut()
        o.d_states = ptr(buf('states', (N, S)))
        o.d_actions = ptr(buf('actions', (N, nu)))
        o.d_masks = ptr(buf('masks', (N,)))
        o.d_next_states = ptr(buf('next_states', (N, S))) if want_next else None
        o.d_rewards = ptr(buf('rewards', (N,)))
        o.d_exps = ptr(buf('exps', (N,)))
        o.d_v_metas = ptr(buf('v_metas', (N, 2), torch.int32))
        o.d_c_info = ptr(buf('c_info', (N, 5)))
        o.d_raw_obs = ptr(buf('raw_obs', (N, S))) if want_raw else None
        o.d_final_qpos = ptr(buf('final_qpos', (n_env, self.nq)))
        o.d_final_qvel = ptr(buf('final_qvel', (n_env, self.nv)))
        o.d_logger = ptr(buf('logger', (LOG['SIZE'],)))
        if want_traj:               # env.data.qpos / qvel before every step (evaluation roll-outs)
            o.d_qpos_traj = ptr(buf('qpos_traj', (N, self.nq)))
            o.d_qvel_traj = ptr(buf('qvel_traj', (N, self.nv)))
        cfg = RolloutCfg()
        cfg.n_env, cfg.horizon, cfg.episode_len, cfg.fr_margin = n_env, horizon, episode_len, fr_margin
        cfg.end_reward = float(end_reward)
        cfg.fix_head_lb = float('nan') if fix_head_lb is None else float(fix_head_lb)
        cfg.noise_rate, cfg.mean_action, cfg.zf_clip = float(noise_rate), int(bool(mean_action)), float(zf_clip)
        cfg.seed, cfg.iteration = int(seed), int(iteration)
        cfg.max_resets = int(reset_take.shape[1]) if reset_take is not None else 0
        cfg.eval_mode = int(eval_mode)          # 0 training, 1 'naivefs', 2 'valuefs'
        inp = RolloutIn()
        inp.d_eps, inp.d_reset_take, inp.d_reset_start = ptr(eps), ptr(reset_take), ptr(reset_start)
        inp.d_mean_flag, inp.d_zf_mean, inp.d_zf_std = ptr(mean_flag), ptr(zf_mean), ptr(zf_std)
        if value_weights is not None:   # Value(MLP) evaluated before every step ('valuefs'): W1,b1,W2,b2,W3,b3 like ``weights``
            vw = PolicyWeights()
            vw.in_dim, vw.h1 = value_weights['W1'].shape[1], value_weights['W1'].shape[0]
            vw.h2, vw.out_dim = value_weights['W2'].shape[0], value_weights['W3'].shape[0]
            vw.d_W1, vw.d_b1, vw.d_W2 = ptr(value_weights['W1']), ptr(value_weights['b1']), ptr(value_weights['W2'])
            vw.d_b2, vw.d_W3, vw.d_b3 = ptr(value_weights['b2']), ptr(value_weights['W3']), ptr(value_weights['b3'])
            self._keep['vw'] = vw
            inp.value_net = C.cast(C.pointer(vw), _vp)
            inp.d_vctx = ptr(vctx)
            if value_stat is not None:  # [n_env, 2] (n, mean) running statistic, read and written back
                if tuple(value_stat.shape) != (n_env, 2):
                    raise EgpError('value_stat must be [n_env, 2]')
                inp.d_value_stat = ptr(value_stat)
            o.d_values = ptr(buf('values', (N,)))
            launches += 3
        if init_qpos is not None:   # [n_env, nq] / [n_env, nv] state set right after the first reset
            if tuple(init_qpos.shape) != (n_env, self.nq) or tuple(init_qvel.shape) != (n_env, self.nv):
                raise EgpError('init_qpos / init_qvel must be [n_env, nq] / [n_env, nv]')
            inp.d_init_qpos, inp.d_init_qvel = ptr(init_qpos), ptr(init_qvel)
        if fix_len is not None:     # per-environment episode length, int32 [n_env]
            inp.d_fix_len = ptr(fix_len)
        if state_pred is not None:  # [total_frames, S] predicted observations (eval_mode state replacement)
            if tuple(state_pred.shape) != (int(self.take_off[-1]), S):
                raise EgpError('state_pred must be [total_frames = %d, %d]' % (int(self.take_off[-1]), S))
            inp.d_state_pred = ptr(state_pred)
        if ctx is not None:         # per-rollout context table (window-indexed when win_off is given)
            inp.d_ctx, inp.ctx_dim = ptr(ctx), ctx.shape[1]
            inp.d_win_off, inp.ctx_mode, inp.ctx_T = ptr(win_off), int(win_off is not None), int(episode_len)
            if ctx_const:           # one row per (take, start) window, constant over the episode (VideoForecastNet v_out)
                inp.ctx_mode, inp.ctx_T = 2, 1
        if snet is not None:        # (W [4H, S + H], b [4H], H): state LSTM stepped inside the kernel
            sW, sb, sH = snet
            inp.d_snet_W, inp.d_snet_b, inp.snet_hdim = ptr(sW), ptr(sb), int(sH)
            inp.d_snet_state = ptr(buf('snet_state', ((n_env + 31) // 32 * 2 * int(sH) * 32,)))
            launches += 1
        pw = PolicyWeights()
        pw.in_dim, pw.h1 = weights['W1'].shape[1], weights['W1'].shape[0]
        pw.h2, pw.out_dim = weights['W2'].shape[0], weights['W3'].shape[0]
        pw.d_W1, pw.d_b1, pw.d_W2, pw.d_b2 = ptr(weights['W1']), ptr(weights['b1']), ptr(weights['W2']), ptr(weights['b2'])
        pw.d_W3, pw.d_b3, pw.d_log_std = ptr(weights['W3']), ptr(weights['b3']), ptr(weights['log_std'])
        check(self.lib.egp_rollout_f64(self.handle, C.byref(pw), C.byref(cfg), C.byref(inp), C.byref(o), stream_ptr()),
              'egp_rollout_f64')
        launches += 5               # 3 weight packs + rollout kernel + logger merge
        return out

    def build_input(self, states, v_metas, masks, horizon):
        global launches
        import torch
        n = states.shape[0]
        x = torch.empty((n, self.ctx_dim + self.S), dtype=torch.float64, device=states.device)
        check(self.lib.egp_build_input_f64(self.handle, ptr(states), ptr(v_metas), ptr(masks), n, horizon, ptr(x),
                                           stream_ptr()), 'egp_build_input_f64')
        launches += 1
        return x


# ---- update-phase kernels (thin wrappers; all tensors CUDA float64 contiguous) ------------------
def gae(rewards, masks, values, gamma, tau, work=None, out=None):
    """K4 (core/common.py:5-21): returns (adv_raw, returns, stats[3] = n, mean, M2); ``out`` may carry the three
    preallocated output tensors"""
    global launches
    import torch
    lib = load()
    n = rewards.numel()
    if out is not None:
        adv, ret, stats = out
    else:
        adv = torch.empty(n, dtype=torch.float64, device=rewards.device)
        ret = torch.empty(n, dtype=torch.float64, device=rewards.device)
        stats = torch.empty(3, dtype=torch.float64, device=rewards.device)
    nbytes = lib.egp_gae_work_bytes(n)
    if work is None or work.numel() < nbytes:
        work = torch.empty(nbytes, dtype=torch.uint8, device=rewards.device)
    check(lib.egp_gae_f64(ptr(rewards), ptr(masks), ptr(values), gamma, tau, n, ptr(adv), ptr(ret), ptr(stats),
                          ptr(work), stream_ptr()), 'egp_gae_f64')
    launches += 1 if n >= lib.egp_gae_set_onepass_min(-1) else 3
    return adv, ret, stats


def gae_set_onepass_min(n):
    """batch size from which egp_gae_f64 takes the one-pass scan (returns the previous value; n < 0 only queries)"""
    return int(load().egp_gae_set_onepass_min(int(n)))


def standardize_(x, stats):
    global launches
    check(load().egp_standardize_f64(ptr(x), x.numel(), ptr(stats), stream_ptr()), 'egp_standardize_f64')
    launches += 1
    return x


def gauss_logp(mu, actions, log_std, out=None):
    global launches
    import torch
    n, a = mu.shape
    if out is None:
        out = torch.empty(n, dtype=torch.float64, device=mu.device)
    check(load().egp_gauss_logp_f64(ptr(mu), ptr(actions), ptr(log_std), n, a, ptr(out), stream_ptr()),
          'egp_gauss_logp_f64')
    launches += 1
    return out


def ppo_loss_grad(mu, actions, log_std, adv, stats, logp0, exps, clip_eps, inv_count, dmu, dlogstd, loss):
    global launches
    n, a = mu.shape
    check(load().egp_ppo_loss_grad_f64(ptr(mu), ptr(actions), ptr(log_std), ptr(adv), ptr(stats), ptr(logp0), ptr(exps),
                                       clip_eps, inv_count, n, a, ptr(dmu), ptr(dlogstd), ptr(loss), stream_ptr()),
          'egp_ppo_loss_grad_f64')
    launches += 1


def value_loss_grad(v, ret, inv_n, dv, loss):
    global launches
    check(load().egp_value_loss_grad_f64(ptr(v), ptr(ret), inv_n, v.numel(), ptr(dv), ptr(loss), stream_ptr()),
          'egp_value_loss_grad_f64')
    launches += 1


def bias_relu_(y, b):
    global launches
    check(load().egp_bias_relu_f64(ptr(y), ptr(b), y.shape[0], y.shape[1], stream_ptr()), 'egp_bias_relu_f64')
    launches += 1
    return y


def relu_bwd_(dy, y):
    global launches
    check(load().egp_relu_bwd_f64(ptr(dy), ptr(y), y.shape[0], y.shape[1], stream_ptr()), 'egp_relu_bwd_f64')
    launches += 1
    return dy


def gather_rows(src, perm, out):
    """out[i] = src[perm[i]] for [n] or [n, dim] float64 tensors; perm int64"""
    global launches
    dim = 1 if src.dim() == 1 else src.shape[1]
    check(load().egp_gather_rows_f64(ptr(src), ptr(perm), perm.numel(), dim, ptr(out), stream_ptr()), 'egp_gather_rows_f64')
    launches += 1
    return out


def relu_bwd_colsum_(dy, y, out):
    global launches
    check(load().egp_relu_bwd_colsum_f64(ptr(dy), ptr(y), y.shape[0], y.shape[1], ptr(out), stream_ptr()),
          'egp_relu_bwd_colsum_f64')
    launches += 1
    return dy


def colsum(x, out):
    global launches
    check(load().egp_colsum_f64(ptr(x), x.shape[0], x.shape[1], ptr(out), stream_ptr()), 'egp_colsum_f64')
    launches += 1
    return out


def col_moments(x, shift=None):
    global launches
    import torch
    out = torch.empty(2 * x.shape[1], dtype=torch.float64, device=x.device)
    check(load().egp_col_moments_f64(ptr(x), x.shape[0], x.shape[1], ptr(shift), ptr(out), stream_ptr()),
          'egp_col_moments_f64')
    launches += 1
    return out


def sumsq(g, out):
    global launches
    check(load().egp_sumsq_f64(ptr(g), g.numel(), ptr(out), stream_ptr()), 'egp_sumsq_f64')
    launches += 1
    return out


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, max_norm=0.0, norm2=None):
    global launches
    check(load().egp_adam_step_f64(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, beta1, beta2, eps, step,
                                   float(max_norm or 0.0), ptr(norm2), stream_ptr()), 'egp_adam_step_f64')
    launches += 1


# ---- float64 dense layers on the int8 tensor cores (csrc/ozaki.cu) ----------------------------------------------
def _pad16(k):
    return (int(k) + 15) // 16 * 16


def oz_slice_rows(x, n_slices, colmax=None, out=None):
    """x [M, K] float64 (row stride >= K) -> (int8 slices [S, M, Kp], int32 exponents [M]); scale constant along K.
    ``colmax`` ([K] float64, zeroed by the caller) additionally receives the column abs-max of x."""
    global launches
    import torch
    M, K = x.shape
    if x.stride(1) != 1:
        raise EgpError('oz_slice_rows: x must be row-major')
    kp = _pad16(K)
    if out is None:
        out = (torch.empty((n_slices, M, kp), dtype=torch.int8, device=x.device),
               torch.empty((M,), dtype=torch.int32, device=x.device))
    sl, ex = out
    check(load().egp_oz_slice_rows_f64(ptr(x), M, K, x.stride(0), n_slices, ptr(sl), kp, ptr(ex), ptr(colmax), stream_ptr()),
          'egp_oz_slice_rows_f64')
    launches += 1
    return sl, ex


def oz_colmax(x, colmax=None):
    global launches
    import torch
    if colmax is None:
        colmax = torch.zeros(x.shape[1], dtype=torch.float64, device=x.device)
    check(load().egp_oz_colmax_f64(ptr(x), x.shape[0], x.shape[1], x.stride(0), ptr(colmax), stream_ptr()), 'egp_oz_colmax_f64')
    launches += 1
    return colmax


def oz_slice_colsT(x, n_slices, colmax, out=None, ones_row=False):
    """x [N, F] float64 -> (int8 slices [S, F, Np] transposed, int32 exponents [F]); scale constant along the rows.
    ``ones_row`` appends the virtual feature F = 1.0 for every sample (bias gradient through the same GEMM)."""
    global launches
    import torch
    N, F = x.shape
    if x.stride(1) != 1:
        raise EgpError('oz_slice_colsT: x must be row-major')
    npad = _pad16(N)
    ft = F + (1 if ones_row else 0)
    if out is None:
        out = (torch.empty((n_slices, ft, npad), dtype=torch.int8, device=x.device),
               torch.empty((ft,), dtype=torch.int32, device=x.device))
    sl, ex = out
    check(load().egp_oz_slice_cols_t_f64(ptr(x), N, F, x.stride(0), n_slices, ptr(colmax), ptr(sl), npad, ptr(ex),
                                         int(bool(ones_row)), stream_ptr()), 'egp_oz_slice_cols_t_f64')
    launches += 2
    return sl, ex


_oz_work = {}


def oz_slice_both(x, n_slices, rowmax, colmax, ones_row=False):
    """x [N, F] float64 -> ((row slices [S, N, Kp32], exps [N]), (transposed slices [S, F (+1), Np], exps [F (+1)])) from ONE read,
    given rowmax ([N] int32/uint32 high words of the row abs-maxima) and colmax ([F] float64 bit patterns)"""
    global launches
    import torch
    N, F = x.shape
    if x.stride(1) != 1:
        raise EgpError('oz_slice_both: x must be row-major')
    kp = (F + 31) // 32 * 32
    npad = _pad16(N)
    ft = F + (1 if ones_row else 0)
    r = (torch.empty((n_slices, N, kp), dtype=torch.int8, device=x.device), torch.empty((N,), dtype=torch.int32, device=x.device))
    t = (torch.empty((n_slices, ft, npad), dtype=torch.int8, device=x.device), torch.empty((ft,), dtype=torch.int32, device=x.device))
    check(load().egp_oz_slice_both_f64(ptr(x), N, F, x.stride(0), n_slices, ptr(rowmax), ptr(colmax), ptr(r[0]), kp, ptr(r[1]),
                                       ptr(t[0]), npad, ptr(t[1]), int(bool(ones_row)), stream_ptr()), 'egp_oz_slice_both_f64')
    launches += 2
    return r, t


def oz_gemm(a, ea, b, eb, bias=None, relu=False, out=None, mask=None, rowmax=None, colmax=None):
    """C [M, N] = A B^T (+ bias, relu; * (mask > 0)) from row-scaled slices a [S, M, Kp], b [S, N, Kp] and exponents
    ea [M], eb [N]; rowmax ([M] int32, zeroed) / colmax ([N] float64, zeroed) optionally receive the abs-maxima of C"""
    global launches
    import torch
    S, M, kp = a.shape
    N = b.shape[1]
    if b.shape[0] != S or b.shape[2] != kp:
        raise EgpError('oz_gemm: slice shapes do not match: %s vs %s' % (tuple(a.shape), tuple(b.shape)))
    if out is None:
        out = torch.empty((M, N), dtype=torch.float64, device=a.device)
    lib = load()
    need = lib.egp_oz_gemm_work_bytes(M, N, kp, S)
    work = None
    if need:
        work = _oz_work.get(a.device)
        if work is None or work.numel() < need:
            work = torch.empty(need, dtype=torch.uint8, device=a.device)
            _oz_work[a.device] = work
    check(lib.egp_oz_gemm_max_f64(ptr(a), ptr(ea), M, ptr(b), ptr(eb), N, kp, S, ptr(bias), int(bool(relu)), ptr(mask),
                                  mask.stride(0) if mask is not None else 0, ptr(out), out.stride(0), ptr(rowmax), ptr(colmax),
                                  ptr(work), need, stream_ptr()), 'egp_oz_gemm_max_f64')
    launches += 2 if need else 1
    return out


# ---- chunked MLP forward / loss / backward on the int8 tensor cores (csrc/oz_mlp.cu) -------------------------------
def _raw(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class OzMlp:
    """Workspace + input-slice cache of egp_oz_mlp_step_f64 for one (in, h1, h2, out) shape.  ``weights`` / ``grads`` are
    6-tuples of contiguous float64 CUDA tensors (W1, b1, W2, b2, W3, b3) in torch nn.Linear layout."""

    def __init__(self, in_dim, h1, h2, out_dim, n_slices=6, chunk_rows=None, device=None):
        import torch
        lib = load()
        self.dims = (int(in_dim), int(h1), int(h2), int(out_dim))
        self.S = int(n_slices)
        self.chunk = int(chunk_rows or lib.egp_oz_mlp_chunk_rows())
        self.device = device
        nbytes = lib.egp_oz_mlp_work_bytes(*self.dims, self.chunk, self.S)
        self.work = torch.empty(nbytes, dtype=torch.uint8, device=device)

    @staticmethod
    def launch_count(n_chunks, bwd, slice_x, fill_cache, dx=False, head1=False):
        """kernels launched by one egp_oz_mlp_step_f64 call (csrc/oz_mlp.cu): weight slicing + per chunk the x slicing
        (unless cached), 3 GEMMs + 2 row / 2 transposed slicings forward, loss, 5 GEMMs + 3 reductions + slicings backward"""
        prep = 3 + (6 if bwd else 0) + (3 if dx else 0)
        x = (3 if (bwd or fill_cache) else 1) if slice_x else 0
        fused = bwd and load().egp_oz_mlp_set_fused_slicing(-1) != 0
        if fused:
            # one-read two-orientation slicing (exponent kernel + slicer per intermediate), no column-maximum passes:
            # 8 GEMMs, 3 x 2 slice_both, dy rows + 2 x 2 transposed, loss, 3 reductions | scalar head: 5 GEMMs, 2 x 2 + 2, 4 head / loss, 2
            per_chunk = x + (17 if head1 else 23) + (1 if dx else 0)
            if head1:
                prep -= 1 + 3
            return prep + n_chunks * per_chunk
        per_chunk = x + (9 + 1 + 17 if bwd else 5) + (1 if dx else 0)
        if head1:       # scalar head streamed in float64: no slices of the last hidden layer, head forward / backward + reduce
            prep -= 1 + (3 if bwd else 0)
            per_chunk -= (3 + 4) if bwd else 1
        return prep + n_chunks * per_chunk

    def new_cache(self, n):
        import torch
        nbytes = load().egp_oz_mlp_xcache_bytes(self.dims[0], n, self.chunk, self.S)
        return dict(buf=torch.empty(nbytes, dtype=torch.uint8, device=self.device), n=int(n), valid=False)

    def step(self, weights, x, grads=None, loss=None, y=None, cache=None, dx=None):
        """loss: None (forward only, returns y), or dict(kind='ppo', actions, log_std, adv, stats, logp0, exps, clip_eps,
        inv_count, dlogstd, loss) / dict(kind='value', returns, inv_n, loss); dx ([n, c], optional) receives dL/dx[:, :c]"""
        global launches
        import torch
        n = x.shape[0]
        if x.stride(1) != 1 or x.shape[1] != self.dims[0] or x.dtype != torch.float64:
            raise EgpError('OzMlp.step: x must be a row-major float64 [n, %d] tensor' % self.dims[0])
        net = MlpNet(*self.dims, *[_raw(w) for w in weights], *([_raw(g) for g in grads] if grads is not None else [None] * 6),
                     _raw(dx), int(dx.shape[1]) if dx is not None else 0)
        if dx is not None and (dx.shape[0] != n or not dx.is_contiguous() or loss is None):
            raise EgpError('OzMlp.step: dx must be a contiguous [n, dx_cols] tensor of a backward pass')
        ls = MlpLoss()
        ls.kind = 0
        if loss is not None:
            if loss['kind'] == 'ppo':
                ls.kind = 1
                ls.d_actions, ls.d_log_std, ls.d_adv, ls.d_stats = (_raw(loss[k]) for k in ('actions', 'log_std', 'adv', 'stats'))
                ls.d_logp0, ls.d_exps = _raw(loss['logp0']), _raw(loss['exps'])
                ls.clip_eps, ls.inv_count = float(loss['clip_eps']), float(loss['inv_count'])
                ls.d_dlogstd = _raw(loss.get('dlogstd'))
                ls.init_logp0 = int(bool(loss.get('init_logp0', False)))
            else:
                ls.kind = 2
                ls.d_returns, ls.inv_n = _raw(loss['returns']), float(loss['inv_n'])
            ls.d_loss = _raw(loss['loss'])
        elif y is None:
            y = torch.empty((n, self.dims[3]), dtype=torch.float64, device=x.device)
        state = 0
        if cache is not None:
            if cache['n'] != n:
                raise EgpError('OzMlp.step: input cache was sized for %d rows, got %d' % (cache['n'], n))
            state = 2 if cache['valid'] else 1
        check(load().egp_oz_mlp_step_f64(C.byref(net), ptr(x[:1]) if not x.is_contiguous() else ptr(x), x.stride(0), n, C.byref(ls),
                                         _raw(y), self.S, self.chunk, _raw(cache['buf']) if cache is not None else None, state,
                                         _raw(self.work), self.work.numel(), stream_ptr()), 'egp_oz_mlp_step_f64')
        launches += self.launch_count((n + self.chunk - 1) // self.chunk, loss is not None, state != 2, state == 1, dx is not None,
                                      head1=self.dims[3] == 1 and self.dims[2] <= 512)
        if cache is not None:
            cache['valid'] = True
        return y


# ---- fused LSTM sequence recurrence (csrc/lstm.cu) -----------------------------------------------
LSTM_FUSED_H = (64, 128)


def lstm_pack_whh(weight_hh):
    """torch weight_hh [4H, H] -> (forward fragments, backward fragments) for lstm_seq_fwd / lstm_seq_bwd"""
    global launches
    import torch
    H = weight_hh.shape[1]
    n = load().egp_lstm_wfrag_elems(H)
    wf = torch.empty(n, dtype=torch.float64, device=weight_hh.device)
    wb = torch.empty(n, dtype=torch.float64, device=weight_hh.device)
    check(load().egp_lstm_pack_whh_f64(ptr(weight_hh.contiguous()), H, ptr(wf), ptr(wb), stream_ptr()), 'egp_lstm_pack_whh_f64')
    launches += 2
    return wf, wb


def lstm_seq_fwd(xi, off, L, B, H, wf):
    """xi [Np, 4H] packed input projections, off int64 [L + 1] on the device -> (h [Np, H], gates [Np, 4H], c [Np, H])"""
    global launches
    import torch
    Np = xi.shape[0]
    h = torch.zeros((Np, H), dtype=torch.float64, device=xi.device)
    gates = torch.empty((Np, 4 * H), dtype=torch.float64, device=xi.device)
    c = torch.empty((Np, H), dtype=torch.float64, device=xi.device)
    check(load().egp_lstm_seq_fwd_f64(ptr(xi), ptr(off), L, B, H, ptr(wf), ptr(h), ptr(gates), ptr(c), stream_ptr()),
          'egp_lstm_seq_fwd_f64')
    launches += 1
    return h, gates, c


def lstm_seq_bwd(dh, gates, c, off, L, B, H, wb):
    global launches
    import torch
    dxi = torch.zeros_like(gates)
    check(load().egp_lstm_seq_bwd_f64(ptr(dh), ptr(gates), ptr(c), ptr(off), L, B, H, ptr(wb), ptr(dxi), stream_ptr()),
          'egp_lstm_seq_bwd_f64')
    launches += 1
    return dxi


# ---- gradient exchange over NVLink peer memory (csrc/p2p.cu) -------------------------------------
class _DevMem:
    """raw device pointer as a __cuda_array_interface__ object (float64 vector) so that torch can wrap it"""

    def __init__(self, ptr_, n, owner):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': '<f8', 'data': (int(ptr_), False), 'version': 2}
        self._owner = owner


class PeerComm:
    """One exchange block per rank, mapped by every peer of the node (include/egopose_b200.h: egp_comm_*).  ``src`` is
    the tensor to write the local gradient into, ``allreduce()`` leaves the rank-ordered sum in ``out`` (one kernel
    launch, no NCCL).  The IPC handles travel through ``torch.distributed`` (any backend) once, at construction."""

    @classmethod
    def local_group(cls, n, devices):
        """the communicators of len(devices) ranks that all live in this process (rank r on devices[r])"""
        import torch
        L = load()
        world = len(devices)
        hb = int(L.egp_comm_handle_bytes())
        comms = []
        for r, dev in enumerate(devices):
            c = cls.__new__(cls)
            c.lib, c.n, c.rank, c.world = L, int(n), r, world
            c.device = torch.device(dev)
            h = C.c_void_p()
            check(L.egp_comm_create(r, world, c.device.index or 0, c.n, C.byref(h), C.create_string_buffer(hb)), 'egp_comm_create')
            c.handle = h
            comms.append(c)
        arr = (C.c_void_p * world)(*[c.handle for c in comms])
        for c in comms:
            check(L.egp_comm_connect_local(c.handle, arr), 'egp_comm_connect_local')
            c.src = torch.as_tensor(_DevMem(L.egp_comm_src(c.handle), c.n, c), device=c.device)
            c.out = torch.as_tensor(_DevMem(L.egp_comm_out(c.handle), c.n, c), device=c.device)
        return comms

    def __init__(self, n, device, dist):
        import torch
        self.lib = load()
        self.n = int(n)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        hb = int(self.lib.egp_comm_handle_bytes())
        mine = C.create_string_buffer(hb)
        h = C.c_void_p()
        dev = torch.device(device)
        check(self.lib.egp_comm_create(self.rank, self.world, dev.index if dev.index is not None else torch.cuda.current_device(),
                                       self.n, C.byref(h), mine), 'egp_comm_create')
        self.handle = h
        blobs = [None] * self.world
        dist.all_gather_object(blobs, bytes(mine.raw))
        allh = C.create_string_buffer(b''.join(blobs), hb * self.world)
        check(self.lib.egp_comm_connect(self.handle, allh), 'egp_comm_connect')
        self.src = torch.as_tensor(_DevMem(self.lib.egp_comm_src(self.handle), self.n, self), device=dev)
        self.out = torch.as_tensor(_DevMem(self.lib.egp_comm_out(self.handle), self.n, self), device=dev)

    def allreduce(self, n=None):
        global launches
        check(self.lib.egp_allreduce_grads_f64(self.handle, int(n if n is not None else self.n), stream_ptr()),
              'egp_allreduce_grads_f64')
        launches += 1
        return self.out

    def error(self):
        return int(self.lib.egp_comm_error(self.handle))

    def close(self):
        if getattr(self, 'handle', None):
            self.src = self.out = None
            self.lib.egp_comm_destroy(self.handle)
            self.handle = None
