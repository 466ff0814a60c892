"""TEST INFRASTRUCTURE ONLY (build container).  Fake ``mujoco_py`` + ``gym`` modules exposing exactly the
surface the reference env touches, backed by the C restatement in egopose_oracle.c, so
that the UNMODIFIED ego_pose/envs/humanoid_v1.py + ego_pose/core/reward_function.py can be executed to
produce golden vectors for the env-logic half of the oracle (tests/golden/make_golden.py).
"""
import ctypes as C
import sys
import types

import numpy as np

from . import cphys, refimport


class _Opt:
    pass


class ShimModel:
    def __init__(self, orc):
        md = orc.md
        self._orc = orc
        self.nq, self.nv, self.nu = md['nq'], md['nv'], md['nu']
        self.opt = _Opt()
        self.opt.timestep = md['timestep']
        self.stat = _Opt()
        self.stat.extent = 3.0
        self.actuator_names = tuple(md['actuator_names'])
        self.actuator_ctrlrange = np.zeros((self.nu, 2))
        self.body_names = ('world',) + tuple(md['body_names'])
        self._body_name2id = {n: i for i, n in enumerate(self.body_names)}
        # joints: free joint = 1 joint; hinges 1 each
        jntadr, jntnum, qposadr = [-1], [0], []
        j = 0
        for b in range(md['nbody']):
            n = 1 if md['body_dofnum'][b] == 6 else md['body_dofnum'][b]
            jntadr.append(j)
            jntnum.append(n)
            if md['body_dofnum'][b] == 6:
                qposadr.append(md['body_qposadr'][b])
            else:
                qposadr.extend(md['body_qposadr'][b] + k for k in range(n))
            j += n
        self.body_jntadr = np.array(jntadr)
        self.body_jntnum = np.array(jntnum)
        self.jnt_qposadr = np.array(qposadr)
        self.jnt_stiffness = np.zeros(j)
        self.dof_damping = np.zeros(self.nv)


class ShimData:
    def __init__(self, orc):
        self._orc = orc
        self._d = cphys.EoData()
        nq, nv, nu, nb = orc.nq, orc.nv, orc.nu, orc.nbody
        d = self._d
        base = C.addressof(d)

        def view(field, count):
            off = getattr(cphys.EoData, field).offset
            return np.ctypeslib.as_array((C.c_double * count).from_address(base + off))
        self.qpos, self.qvel, self.ctrl = view('qpos', nq), view('qvel', nv), view('ctrl', nu)
        self.qfrc_bias = view('qfrc_bias', nv)
        self.qM = view('qM', nv * nv)           # dense here; fake mj_fullM just copies it
        self._xpos = view('xpos', 3 * cphys.MAXB).reshape(cphys.MAXB, 3)
        self._com = view('subtree_com', 3)
        self._nb = nb

    @property
    def body_xpos(self):
        return np.vstack([np.zeros((1, 3)), self._xpos[:self._nb]])

    @property
    def subtree_com(self):
        return self._com[None, :].copy()

    def get_body_xpos(self, name):
        return self._xpos[self._orc.md['body_names'].index(name)].copy()


class MjSimState:
    def __init__(self, time, qpos, qvel, act, udd_state):
        self.time, self.qpos, self.qvel, self.act, self.udd_state = time, qpos, qvel, act, udd_state


class MjSim:
    def __init__(self, model):
        self.model = model
        self._orc = model._orc
        self.data = ShimData(self._orc)
        self.reset()

    def reset(self):
        self.data.qpos[:] = self._orc.md['qpos0']
        self.data.qvel[:] = 0
        self.data.ctrl[:] = 0
        self._time = 0.0

    def get_state(self):
        return MjSimState(self._time, self.data.qpos.copy(), self.data.qvel.copy(), None, {})

    def set_state(self, s):
        self._time = s.time
        self.data.qpos[:] = s.qpos
        self.data.qvel[:] = s.qvel

    def forward(self):
        self._orc.forward(self.data._d)

    def step(self):
        self._orc.step(self.data._d)
        self._time += self.model.opt.timestep


def install(orc):
    """Install fake mujoco_py / gym modules backed by ``orc`` (a cphys.Oracle); returns nothing."""
    refimport.install()
    mj = sys.modules['mujoco_py']
    mj.load_model_from_path = lambda path: ShimModel(orc)
    mj.MjSim = MjSim
    mj.MjSimState = MjSimState
    fn = sys.modules['mujoco_py.functions']

    def mj_fullM(model, dst, qM):
        dst[:] = qM
    fn.mj_fullM = mj_fullM
    mj.functions = fn
    # names envs/common/mjviewer.py:3-8 imports at module level (viewer classes are never instantiated)
    cymj = types.SimpleNamespace(MjRenderContextWindow=object)
    sys.modules['mujoco_py.builder'].cymj = cymj
    sys.modules['mujoco_py.generated'].const = types.SimpleNamespace()
    sys.modules['mujoco_py.utils'].rec_copy = lambda x: x
    sys.modules['mujoco_py.utils'].rec_assign = lambda a, b: None
    gym = sys.modules['gym']

    class Box:
        def __init__(self, low, high, dtype=np.float32):
            self.low, self.high = low, high
            self.shape = np.asarray(low).shape
    spaces = sys.modules['gym.spaces']
    spaces.Box = Box
    gym.spaces = spaces
    seeding = types.SimpleNamespace(np_random=lambda seed=None: (np.random.RandomState(seed), seed))
    sys.modules['gym.utils'].seeding = seeding
    gym.utils = sys.modules['gym.utils']
    gym.error = types.SimpleNamespace()
