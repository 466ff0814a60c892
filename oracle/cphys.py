"""TEST INFRASTRUCTURE ONLY - ctypes front-end of oracle/egopose_oracle.c (CPU float64 restatement).

Never imported by egopose_b200/.  Loads the model / hyper-parameter constants from the JSON files in
egopose_b200/assets (plain data compiled from the reference's XML / yml by tools/compile_model.py).
"""
import ctypes as C
import json
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ASSETS = os.path.join(os.path.dirname(HERE), 'egopose_b200', 'assets')
MAXB, MAXV, NEE = 32, 64, 5
X = dict(QPOS=0, QVEL=59, RLINV_LOCAL=117, RANGV=120, RQ_RMH=123, EE_POS=127, BQUAT=142, BANGVEL=226, STRIDE=292)
EE_NAMES = ['LeftFoot', 'RightFoot', 'LeftHand', 'RightHand', 'Head']   # humanoid_v1.py:100

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class EoModel(C.Structure):
    _fields_ = [('nq', C.c_int), ('nv', C.c_int), ('nu', C.c_int), ('nbody', C.c_int),
                ('timestep', C.c_double), ('gravity', C.c_double * 3),
                ('body_parent', _ip), ('body_dofadr', _ip), ('body_dofnum', _ip), ('body_qposadr', _ip),
                ('body_pos', _dp), ('body_mass', _dp), ('body_ipos', _dp), ('body_inertia', _dp),
                ('dof_body', _ip), ('dof_parent', _ip),
                ('dof_armature', _dp), ('dof_axis', _dp), ('dof_anchor', _dp),
                ('ee_body', C.c_int * NEE), ('head_body', C.c_int),
                ('dof_range', _dp), ('dof_invweight0', _dp), ('solref', C.c_double * 2), ('solimp', C.c_double * 5),
                ('geom_type', _ip), ('geom_size', _dp), ('geom_p0', _dp), ('geom_p1', _dp), ('body_invweight0', _dp),
                ('contact_margin', C.c_double), ('contact_mu', C.c_double)]


class EoData(C.Structure):
    _fields_ = [('qpos', C.c_double * (MAXV + 1)), ('qvel', C.c_double * MAXV), ('ctrl', C.c_double * MAXV),
                ('qacc', C.c_double * MAXV),
                ('xpos', C.c_double * 3 * MAXB), ('xquat', C.c_double * 4 * MAXB), ('xipos', C.c_double * 3 * MAXB),
                ('qM', C.c_double * (MAXV * MAXV)), ('qfrc_bias', C.c_double * MAXV),
                ('cdof', C.c_double * 6 * MAXV), ('subtree_com', C.c_double * 3), ('n_efc', C.c_int), ('solver_iter', C.c_int)]


class EoCfg(C.Structure):
    _fields_ = [('frame_skip', C.c_int), ('episode_len', C.c_int), ('fr_margin', C.c_int),
                ('jkp', _dp), ('jkd', _dp), ('a_ref', _dp), ('a_scale', _dp), ('torque_lim', _dp), ('b_diffw', _dp),
                ('w_p', C.c_double), ('w_v', C.c_double), ('w_e', C.c_double), ('w_rp', C.c_double),
                ('w_rv', C.c_double), ('k_p', C.c_double), ('k_v', C.c_double), ('k_e', C.c_double),
                ('k_rh', C.c_double), ('k_rq', C.c_double), ('k_rl', C.c_double), ('k_ra', C.c_double),
                ('v_ord', C.c_int), ('decay', C.c_int), ('end_reward', C.c_double), ('fix_head_lb', C.c_double)]


class EoExpert(C.Structure):
    _fields_ = [('n_takes', C.c_int), ('take_off', _ip), ('rows', _dp), ('head_height_lb', _dp),
                ('ctx', _dp), ('ctx_dim', C.c_int)]


class EoPolicy(C.Structure):
    _fields_ = [('in_dim', C.c_int), ('h1', C.c_int), ('h2', C.c_int), ('out_dim', C.c_int),
                ('W1', _dp), ('b1', _dp), ('W2', _dp), ('b2', _dp), ('W3', _dp), ('b3', _dp), ('log_std', _dp)]


class EoEnv(C.Structure):
    _fields_ = [('d', EoData), ('cur_t', C.c_int), ('take', C.c_int), ('start_ind', C.c_int),
                ('prev_qpos', C.c_double * (MAXV + 1)), ('prev_qvel', C.c_double * MAXV),
                ('bquat', C.c_double * (4 * MAXB)), ('prev_bquat', C.c_double * (4 * MAXB))]


def build(force=False):
    so = os.path.join(HERE, 'libegopose_oracle.so')
    src = [os.path.join(HERE, f) for f in ('egopose_oracle.c', 'egopose_oracle.h')]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        cc = '/usr/bin/gcc' if os.path.exists('/usr/bin/gcc') else 'gcc'
        subprocess.check_call(['make', '-C', HERE, 'CC=' + cc, '-B', '-s'])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.eo_reward.restype = C.c_double
        _lib.eo_rollout.restype = C.c_int
        _lib.eo_chol_solve.restype = C.c_int
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(_dp if a.dtype == np.float64 else _ip)


def default_cfg_dict(task='egomimic', cfg_id='subject_03'):
    return json.load(open(os.path.join(ASSETS, '%s_%s.cfg.json' % (task, cfg_id))))


class Oracle:
    """Model + cfg bound to the C oracle.  ``cfg`` is the yml dict (config/egomimic/*.yml)."""

    def __init__(self, cfg=None, model_json=None, episode_len=None, fix_head_lb=None, joint_limits=False, contacts=False):
        self.L = lib()
        md = json.load(open(model_json or os.path.join(ASSETS, 'humanoid_1205_v1.model.json')))
        self.md = md
        self.nq, self.nv, self.nu, self.nbody = md['nq'], md['nv'], md['nu'], md['nbody']
        self.S = self.nq - 2 + self.nv
        self._keep = k = {}
        for name in ('body_parent', 'body_dofadr', 'body_dofnum', 'body_qposadr', 'dof_body', 'dof_parent'):
            k[name] = _i(md[name])
        for name in ('body_pos', 'body_mass', 'body_ipos', 'body_inertia', 'dof_armature', 'dof_axis', 'dof_anchor'):
            k[name] = _d(md[name])
        m = EoModel()
        m.nq, m.nv, m.nu, m.nbody, m.timestep = md['nq'], md['nv'], md['nu'], md['nbody'], md['timestep']
        m.gravity[:] = md['gravity']
        for name in k:
            setattr(m, name, _p(k[name]))
        m.ee_body[:] = [md['body_names'].index(n) for n in EE_NAMES]
        m.head_body = md['body_names'].index('Head')
        self.model = m
        cfg = cfg if cfg is not None else default_cfg_dict()
        self.cfg_dict = cfg
        jp = list(zip(*cfg['joint_params']))
        mult = cfg.get('jkp_multiplier', 1.0)
        k['jkp'] = _d(jp[1]) * mult                                  # egomimic_config.py:108-116
        k['jkd'] = _d(jp[2]) * cfg.get('jkd_multiplier', mult)
        k['a_ref'] = np.deg2rad(_d(jp[3]))
        k['a_scale'] = _d(jp[4])
        k['torque_lim'] = _d(jp[5])
        k['b_diffw'] = _d(list(zip(*cfg['body_params']))[1])         # egomimic_config.py:119-122
        ws = cfg.get('reward_weights', {}) or {}
        c = EoCfg()
        c.frame_skip = 15                                            # humanoid_v1.py:16
        c.episode_len = episode_len if episode_len is not None else cfg.get('env_episode_len', 200)
        c.fr_margin = cfg.get('fr_margin', 10)
        for name in ('jkp', 'jkd', 'a_ref', 'a_scale', 'torque_lim', 'b_diffw'):
            setattr(c, name, _p(k[name]))
        # defaults of reward_function.py:8-12
        for name, dv in (('w_p', 0.5), ('w_v', 0.1), ('w_e', 0.2), ('w_rp', 0.1), ('w_rv', 0.1), ('k_p', 2), ('k_v', 0.005),
                         ('k_e', 20), ('k_rh', 300), ('k_rq', 300), ('k_rl', 5.0), ('k_ra', 0.5)):
            setattr(c, name, float(ws.get(name, dv)))
        c.v_ord = int(ws.get('v_ord', 2))
        c.decay = int(bool(ws.get('decay', False)))
        c.end_reward = 0.0
        c.fix_head_lb = float('nan') if fix_head_lb is None else float(fix_head_lb)
        self.cfg = c
        self.dt = md['timestep'] * 15
        if joint_limits:
            self.enable_joint_limits()
        if contacts:
            self.enable_contacts()

    def dof_ranges(self):
        """[nv][2] (radians): the hinge ranges of the XML per dof; the free root has none (0, 0)"""
        rng = np.zeros((self.nv, 2))
        free = self.md['body_dofnum'][0] == 6
        for j, r in enumerate(self.md['jnt_range']):
            if free and j == 0:
                continue
            rng[j + 5 if free else j] = r
        return rng

    def invweight0(self):
        """mjModel.dof_invweight0 of the hinges: diag(M^-1) at qpos0"""
        return np.ascontiguousarray(np.diag(self.minv0()[0]))

    def minv0(self):
        """M^-1 at qpos0 and the data it was computed from (smooth forward)"""
        d = self.new_data(self.md['qpos0'], np.zeros(self.nv))
        smooth = EoModel.from_buffer_copy(self.model)       # a copy: pointer fields read from a Structure alias its memory
        smooth.dof_range, smooth.geom_type = None, None
        self.L.eo_forward(C.byref(smooth), C.byref(d))
        return np.linalg.inv(self.qM(d)), d

    def body_invweight0(self):
        """mjModel.body_invweight0: mean diagonal of J M^-1 J^T at qpos0 for the translational and the rotational Jacobian
        of every body's centre of mass (engine_setconst.c: set0)"""
        Mi, d = self.minv0()
        cdof = np.array(d.cdof).reshape(MAXV, 6)[:self.nv]
        xipos = np.array(d.xipos).reshape(MAXB, 3)[:self.nbody]
        out = np.zeros((self.nbody, 2))
        for b in range(self.nbody):
            J = np.zeros((6, self.nv))
            i = self.md['body_dofadr'][b] + self.md['body_dofnum'][b] - 1
            while i >= 0:
                J[3:, i] = cdof[i, :3]                                      # rotational
                J[:3, i] = cdof[i, 3:] + np.cross(cdof[i, :3], xipos[b])    # velocity of the com: v_O + w x c
                i = self.md['dof_parent'][i]
            A = J @ Mi @ J.T
            out[b] = [np.trace(A[:3, :3]) / 3.0, np.trace(A[3:, 3:]) / 3.0]
        return out

    def enable_contacts(self, margin=0.001, mu=1.0, solref=(0.02, 1.0), solimp=(0.9, 0.95, 0.001, 0.5, 2.0)):
        """floor contacts with MuJoCo's defaults / the XML's values (margin 0.001 on every geom, floor friction 1)"""
        k = self._keep
        k['body_invweight0'] = np.ascontiguousarray(self.body_invweight0())
        k['geom_type'] = _i(self.md['geom_type'])
        for name in ('geom_size', 'geom_p0', 'geom_p1'):
            k[name] = _d(self.md[name])
        self.model.solref[:] = solref
        self.model.solimp[:] = solimp
        self.model.contact_margin, self.model.contact_mu = margin, mu
        for name in ('geom_size', 'geom_p0', 'geom_p1', 'body_invweight0'):
            setattr(self.model, name, _p(k[name]))
        self.model.geom_type = _p(k['geom_type'])

    def enable_joint_limits(self, solref=(0.02, 1.0), solimp=(0.9, 0.95, 0.001, 0.5, 2.0)):
        """MuJoCo's defaults (the XML sets none)"""
        k = self._keep
        k['dof_invweight0'] = self.invweight0()
        k['dof_range'] = np.ascontiguousarray(self.dof_ranges())
        self.model.solref[:] = solref
        self.model.solimp[:] = solimp
        self.model.dof_invweight0 = _p(k['dof_invweight0'])
        self.model.dof_range = _p(k['dof_range'])

    # ---- physics ---------------------------------------------------------------------------
    def new_data(self, qpos, qvel, ctrl=None):
        d = EoData()
        d.qpos[:self.nq] = list(qpos)
        d.qvel[:self.nv] = list(qvel)
        if ctrl is not None:
            d.ctrl[:self.nu] = list(ctrl)
        return d

    def forward(self, d):
        self.L.eo_forward(C.byref(self.model), C.byref(d))

    def step(self, d):
        self.L.eo_step(C.byref(self.model), C.byref(d))

    def qM(self, d):
        return np.array(d.qM[:self.nv * self.nv]).reshape(self.nv, self.nv)

    def kinematics(self, qpos):
        """-> xpos[nb,3], xquat[nb,4], xipos[nb,3], axis_w[nv,3], anchor_w[nv,3]"""
        d = self.new_data(qpos, np.zeros(self.nv))
        ax = np.zeros((MAXV, 3))
        an = np.zeros((MAXV, 3))
        self.L.eo_kinematics(C.byref(self.model), C.byref(d), _p(ax), _p(an))
        nb = self.nbody
        return (np.array(d.xpos).reshape(MAXB, 3)[:nb].copy(), np.array(d.xquat).reshape(MAXB, 4)[:nb].copy(),
                np.array(d.xipos).reshape(MAXB, 3)[:nb].copy(), ax[:self.nv], an[:self.nv])

    def compute_torque(self, d, ctrl):
        out = np.zeros(self.nu)
        ctrl = _d(ctrl)
        self.L.eo_compute_torque(C.byref(self.model), C.byref(self.cfg), C.byref(d), _p(ctrl), _p(out))
        return out

    # ---- expert tables ---------------------------------------------------------------------
    def expert_features(self, qpos):
        """gen_expert.get_expert for one take -> (rows [L,292], head_height_lb)"""
        qpos = _d(qpos)
        rows = np.zeros((qpos.shape[0], X['STRIDE']))
        hz = C.c_double()
        self.L.eo_expert_features(C.byref(self.model), qpos.shape[0], _p(qpos), C.c_double(self.dt), _p(rows), C.byref(hz))
        return rows, hz.value

    def expert_extras(self, qpos):
        """the expert-dict keys that are not rollout inputs (gen_expert.py:44,49-50): head_pos = get_body_com('Head')
        (body frame origin, humanoid_v1.py / mujoco_env.py get_body_com), com = data.subtree_com[0] (whole-body centre
        of mass, mujoco_env.py get_com), ee_wpos = get_ee_pos(None) (humanoid_v1.py:98-111): [L,3], [L,3], [L,15]"""
        mass = np.asarray(self.md['body_mass'], dtype=np.float64)
        ee = [self.md['body_names'].index(n) for n in EE_NAMES]
        head = self.md['body_names'].index('Head')
        hp, com, eew = [], [], []
        for q in np.asarray(qpos, dtype=np.float64):
            xpos, _, xipos, _, _ = self.kinematics(q)
            hp.append(xpos[head])
            com.append((mass[:, None] * xipos).sum(0) / mass.sum())
            eew.append(np.concatenate([xpos[b] for b in ee]))          # transform None: world positions, root not subtracted
        return np.array(hp), np.array(com), np.array(eew)

    def make_expert(self, takes_qpos, ctx=None):
        rows, lbs, off = [], [], [0]
        for q in takes_qpos:
            r, lb = self.expert_features(q)
            rows.append(r)
            lbs.append(lb)
            off.append(off[-1] + r.shape[0])
        return self.pack_expert(np.concatenate(rows), off, lbs, ctx)

    def pack_expert(self, rows, take_off, head_lb, ctx=None):
        k = self._keep
        k['x_rows'], k['x_off'], k['x_lb'] = _d(rows), _i(take_off), _d(head_lb)
        x = EoExpert()
        x.n_takes = len(head_lb)
        x.take_off, x.rows, x.head_height_lb = _p(k['x_off']), _p(k['x_rows']), _p(k['x_lb'])
        if ctx is not None:
            k['x_ctx'] = _d(ctx)
            x.ctx, x.ctx_dim = _p(k['x_ctx']), k['x_ctx'].shape[1]
        else:
            x.ctx, x.ctx_dim = None, 0
        self.expert = x
        return x

    # ---- env ---------------------------------------------------------------------------------
    def env_reset(self, env, take, start):
        self.L.eo_env_reset(C.byref(self.model), C.byref(self.cfg), C.byref(self.expert), C.byref(env), take, start)

    def env_step(self, env, action):
        a = _d(action)
        fail, end = C.c_int(), C.c_int()
        self.L.eo_env_step(C.byref(self.model), C.byref(self.cfg), C.byref(self.expert), C.byref(env), _p(a),
                           C.byref(fail), C.byref(end))
        return bool(fail.value), bool(end.value)

    def env_set_state(self, env, qpos, qvel):
        """MujocoEnv.set_state (envs/common/mujoco_env.py:95-101): overwrite qpos / qvel, sim.forward()"""
        q, v = _d(qpos), _d(qvel)
        self.L.eo_env_set_state(C.byref(self.model), C.byref(env), _p(q), _p(v))

    def policy_mean(self, policy, x):
        """PolicyGaussian action mean of one input row (core/policy_gaussian.py:19-24)"""
        x = _d(x)
        mean, scratch = np.zeros(policy.out_dim), np.zeros(policy.h1 + policy.h2)
        self.L.eo_policy_mean(C.byref(policy), _p(x), _p(mean), _p(scratch))
        return mean

    def env_obs(self, env):
        out = np.zeros(self.S)
        self.L.eo_env_obs(C.byref(self.model), C.byref(env), _p(out))
        return out

    def env_reward(self, env, end):
        info = np.zeros(5)
        r = self.L.eo_reward(C.byref(self.model), C.byref(self.cfg), C.byref(self.expert), C.byref(env), int(end), _p(info))
        return r, info

    # ---- rollout -----------------------------------------------------------------------------
    def make_policy(self, W1, b1, W2, b2, W3, b3, log_std):
        k = self._keep
        arrs = [_d(a) for a in (W1, b1, W2, b2, W3, b3, np.ravel(log_std))]
        k['pol'] = arrs
        p = EoPolicy()
        p.in_dim, p.h1, p.h2, p.out_dim = arrs[0].shape[1], arrs[0].shape[0], arrs[2].shape[0], arrs[4].shape[0]
        p.W1, p.b1, p.W2, p.b2, p.W3, p.b3, p.log_std = [_p(a) for a in arrs]
        return p

    def rollout(self, policy, n_env, T, reset_take, reset_start, eps, mean_flag=None, zf_mean=None, zf_std=None,
                zf_clip=5.0, n_threads=1, want_next=True):
        N, S, nu = n_env * T, self.S, self.nu
        reset_take, reset_start = _i(reset_take), _i(reset_start)
        eps = _d(eps)
        out = dict(states=np.zeros((N, S)), actions=np.zeros((N, nu)), rewards=np.zeros(N), masks=np.zeros(N),
                   next_states=np.zeros((N, S)) if want_next else None, exps=np.zeros(N),
                   v_metas=np.zeros((N, 2), dtype=np.int32), c_info=np.zeros((N, 5)), raw_obs=np.zeros((N, S)),
                   final_qpos=np.zeros((n_env, self.nq)), final_qvel=np.zeros((n_env, self.nv)))
        mf = None if mean_flag is None else np.ascontiguousarray(mean_flag, dtype=np.uint8)
        zm = None if zf_mean is None else _d(zf_mean)
        zs = None if zf_std is None else _d(zf_std)
        rc = self.L.eo_rollout(
            C.byref(self.model), C.byref(self.cfg), C.byref(self.expert), C.byref(policy), n_env, T,
            reset_take.shape[1], _p(reset_take), _p(reset_start), _p(eps),
            None if mf is None else mf.ctypes.data_as(C.POINTER(C.c_ubyte)), _p(zm), _p(zs), C.c_double(zf_clip),
            _p(out['states']), _p(out['actions']), _p(out['rewards']), _p(out['masks']), _p(out['next_states']),
            _p(out['exps']), _p(out['v_metas']), _p(out['c_info']), _p(out['raw_obs']), _p(out['final_qpos']),
            _p(out['final_qvel']), int(n_threads))
        if rc != 0:
            raise RuntimeError('eo_rollout failed: %d' % rc)
        return out


def synthetic_takes(md, n_takes, length, seed=1, dt=1.0 / 30):
    """Seeded smooth expert qpos trajectories inside the joint ranges (SURVEY 8d 'synthetic inputs'):
    root xy random walk (<= 1 m/s), z = 0.90 + 0.02 sin, yaw(t) * small tilt, per-joint sinusoids
    (0.2-1 Hz, amplitude 1/4 range centred in range), hands zeroed (gen_expert.py:38-39)."""
    rng = np.random.RandomState(seed)
    nq = md['nq']
    rngs = np.array(md['jnt_range'][1:])
    names = md['body_names']
    qadr = md['body_qposadr']
    takes = []
    t = np.arange(length) * dt
    for _ in range(n_takes):
        q = np.zeros((length, nq))
        v = rng.uniform(-1, 1, size=2) * 0.5
        ph = rng.uniform(0, 2 * math.pi, size=4)
        q[:, 0] = v[0] * t + 0.1 * np.sin(0.5 * t + ph[0])
        q[:, 1] = v[1] * t + 0.1 * np.sin(0.4 * t + ph[1])
        q[:, 2] = 0.90 + 0.02 * np.sin(2.0 * t + ph[2])
        yaw = rng.uniform(-math.pi, math.pi) + 0.3 * np.sin(0.3 * t + ph[3])
        tilt = 0.05 * np.sin(1.1 * t + ph[0])
        # yaw about z then small tilt about x: q = qz * qx
        q[:, 3] = np.cos(yaw / 2) * np.cos(tilt / 2)
        q[:, 4] = np.cos(yaw / 2) * np.sin(tilt / 2)
        q[:, 5] = np.sin(yaw / 2) * np.sin(tilt / 2)
        q[:, 6] = np.sin(yaw / 2) * np.cos(tilt / 2)
        freq = rng.uniform(0.2, 1.0, size=nq - 7) * 2 * math.pi
        phase = rng.uniform(0, 2 * math.pi, size=nq - 7)
        mid = 0.5 * (rngs[:, 0] + rngs[:, 1])
        amp = 0.25 * 0.5 * (rngs[:, 1] - rngs[:, 0])
        q[:, 7:] = mid + amp * np.sin(freq * t[:, None] + phase)
        for hand in ('LeftHand', 'RightHand'):
            a = qadr[names.index(hand)]
            q[:, a:a + 3] = 0.0
        takes.append(q)
    return takes
