"""MJCF-subset compiler: humanoid XML -> flat model constants for the rollout kernels.

Host-side replacement for what ``mujoco_py.load_model_from_path`` gives the reference
(envs/common/mujoco_env.py:22) restricted to what the hot path reads
(ego_pose/envs/humanoid_v1.py:27,58-59,105,116,134; utils/tools.py:55-68).

Handled subset (everything humanoid_1205_v1.xml uses, assets/mujoco_models/humanoid_1205_v1.xml:2-192):
  * <compiler coordinate="global" angle="degree" inertiafromgeom="true">
  * <default><joint armature= .../></default>, <option timestep=>
  * nested <body pos=> with one <geom type=sphere|capsule|box> each (density 1000)
  * one free joint on the root body, 1..3 hinge joints on the others
  * <actuator><motor joint= gear=>

Semantics followed (MuJoCo 2.x compile step): with global coordinates every
body frame is world-aligned at qpos0, local offsets are differences of global positions; a body's
inertial properties are those of its single geom at density 1000.
"""
from __future__ import annotations

import json
import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field, asdict

import numpy as np

DENSITY = 1000.0
GRAVITY = (0.0, 0.0, -9.81)


def _vec(s, n=None):
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise ValueError('expected %d numbers, got %r' % (n, s))
    return v


def _frame_from_z(z):
    """Rotation whose third column is the unit vector ``z`` (capsule axis)."""
    z = z / np.linalg.norm(z)
    helper = np.array([1.0, 0.0, 0.0]) if abs(z[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
    x = np.cross(helper, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return np.stack([x, y, z], axis=1)


def geom_inertial(geom):
    """(mass, global centre, 3x3 inertia about the centre in world axes) of one geom."""
    gtype = geom.get('type', 'sphere')
    size = _vec(geom.get('size'))
    if gtype == 'sphere':
        r = size[0]
        mass = DENSITY * 4.0 / 3.0 * math.pi * r ** 3
        centre = _vec(geom.get('pos', '0 0 0'), 3)
        inertia = np.eye(3) * (0.4 * mass * r * r)
    elif gtype == 'capsule':
        ft = _vec(geom.get('fromto'), 6)
        p0, p1 = ft[:3], ft[3:]
        r = size[0]
        h = float(np.linalg.norm(p1 - p0))          # cylinder length
        m_cyl = DENSITY * math.pi * r * r * h
        m_sph = DENSITY * 4.0 / 3.0 * math.pi * r ** 3
        mass = m_cyl + m_sph
        i_sph = 0.4 * m_sph * r * r
        i_perp = m_cyl * (3.0 * r * r + h * h) / 12.0 + i_sph + m_sph * h * (3.0 * r + 2.0 * h) / 8.0
        i_axial = 0.5 * m_cyl * r * r + i_sph
        rot = _frame_from_z(p1 - p0)
        inertia = rot @ np.diag([i_perp, i_perp, i_axial]) @ rot.T
        centre = 0.5 * (p0 + p1)
    elif gtype == 'box':
        sx, sy, sz = size[:3]
        mass = DENSITY * 8.0 * sx * sy * sz
        centre = _vec(geom.get('pos', '0 0 0'), 3)
        inertia = np.diag([mass * (sy * sy + sz * sz) / 3.0,
                           mass * (sx * sx + sz * sz) / 3.0,
                           mass * (sx * sx + sy * sy) / 3.0])
        quat = _vec(geom.get('quat', '1 0 0 0'), 4)
        if not np.allclose(np.abs(quat), [1, 0, 0, 0]):
            raise NotImplementedError('rotated box geoms')
    else:
        raise NotImplementedError('geom type %s' % gtype)
    return mass, centre, inertia


@dataclass
class ModelDesc:
    """Flat constants; all lists are row-major python lists so the object JSON round-trips."""
    nq: int = 0
    nv: int = 0
    nu: int = 0
    nbody: int = 0                      # moving bodies (world excluded)
    timestep: float = 0.0
    gravity: list = field(default_factory=lambda: list(GRAVITY))
    body_names: list = field(default_factory=list)
    body_parent: list = field(default_factory=list)     # index into bodies, -1 = world
    body_pos: list = field(default_factory=list)        # [nbody][3] offset from parent frame origin
    body_mass: list = field(default_factory=list)
    body_ipos: list = field(default_factory=list)       # [nbody][3] COM in body frame
    body_inertia: list = field(default_factory=list)    # [nbody][6] xx yy zz xy xz yz about COM, body axes
    body_dofadr: list = field(default_factory=list)     # first dof of the body
    body_dofnum: list = field(default_factory=list)
    body_qposadr: list = field(default_factory=list)    # first qpos of the body's joints
    dof_body: list = field(default_factory=list)
    dof_parent: list = field(default_factory=list)      # parent dof, -1 for dof 0
    dof_armature: list = field(default_factory=list)
    dof_axis: list = field(default_factory=list)        # [nv][3] hinge axis in body frame (free joint: unit axes)
    dof_anchor: list = field(default_factory=list)      # [nv][3] joint anchor in body frame
    joint_names: list = field(default_factory=list)     # one per joint (free joint counts once)
    jnt_range: list = field(default_factory=list)       # [njnt][2] radians (free joint: 0 0)
    actuator_names: list = field(default_factory=list)
    actuator_dof: list = field(default_factory=list)
    actuator_gear: list = field(default_factory=list)
    qpos0: list = field(default_factory=list)
    # collision geoms (one per body): 0 sphere | 1 capsule | 2 box; size = radius | radius | half extents;
    # p0 = centre (sphere, box) or first end point (capsule), p1 = second end point, both in the body frame
    geom_type: list = field(default_factory=list)
    geom_size: list = field(default_factory=list)
    geom_p0: list = field(default_factory=list)
    geom_p1: list = field(default_factory=list)

    # -- helpers mirroring what the reference reads from mujoco_py -------------------------------
    def body_qposaddr(self):
        """utils/tools.py:55-68 get_body_qposaddr -> {body: (start, end)}"""
        out = {}
        for b, name in enumerate(self.body_names):
            start = self.body_qposadr[b]
            n = 7 if self.body_dofnum[b] == 6 else self.body_dofnum[b]
            out[name] = (start, start + n)
        return out

    def total_mass(self):
        return float(sum(self.body_mass))

    def to_json(self, path):
        with open(path, 'w') as f:
            json.dump(asdict(self), f, indent=1)

    @staticmethod
    def from_json(path):
        with open(path) as f:
            return ModelDesc(**json.load(f))


def compile_mjcf(path) -> ModelDesc:
    root = ET.parse(path).getroot()
    comp = root.find('compiler')
    if comp is None or comp.get('coordinate') != 'global' or comp.get('angle', 'degree') != 'degree' \
            or comp.get('inertiafromgeom') != 'true':
        raise NotImplementedError('only coordinate=global / angle=degree / inertiafromgeom=true models are supported')
    jdef = root.find('default/joint')
    def_armature = float(jdef.get('armature', 0.0)) if jdef is not None else 0.0
    m = ModelDesc()
    m.timestep = float(root.find('option').get('timestep'))

    world = root.find('worldbody')
    gpos = []           # global body positions
    qadr = 0

    def visit(elem, parent):
        nonlocal qadr
        b = len(m.body_names)
        pos = _vec(elem.get('pos'), 3)
        gpos.append(pos)
        m.body_names.append(elem.get('name'))
        m.body_parent.append(parent)
        m.body_pos.append(list(pos - (gpos[parent] if parent >= 0 else 0.0)))
        geoms = elem.findall('geom')
        if len(geoms) != 1:
            raise NotImplementedError('exactly one geom per body expected (%s)' % elem.get('name'))
        mass, centre, inertia = geom_inertial(geoms[0])
        gtype = geoms[0].get('type', 'sphere')
        gsize = list(_vec(geoms[0].get('size'))) + [0.0, 0.0]
        m.geom_type.append({'sphere': 0, 'capsule': 1, 'box': 2}[gtype])
        m.geom_size.append(gsize[:3])
        if gtype == 'capsule':
            ft = _vec(geoms[0].get('fromto'), 6)
            m.geom_p0.append(list(ft[:3] - pos))
            m.geom_p1.append(list(ft[3:] - pos))
        else:
            m.geom_p0.append(list(_vec(geoms[0].get('pos', '0 0 0'), 3) - pos))
            m.geom_p1.append([0.0, 0.0, 0.0])
        m.body_mass.append(mass)
        m.body_ipos.append(list(centre - pos))
        m.body_inertia.append([inertia[0, 0], inertia[1, 1], inertia[2, 2], inertia[0, 1], inertia[0, 2], inertia[1, 2]])
        joints = elem.findall('joint')
        m.body_dofadr.append(len(m.dof_body))
        m.body_qposadr.append(qadr)
        pdof = -1 if parent < 0 else m.body_dofadr[parent] + m.body_dofnum[parent] - 1
        if joints and joints[0].get('type') == 'free':
            if len(joints) != 1 or parent >= 0:
                raise NotImplementedError('free joint only as the single joint of the root body')
            m.body_dofnum.append(6)
            m.joint_names.append(joints[0].get('name'))
            m.jnt_range.append([0.0, 0.0])
            arm = float(joints[0].get('armature', def_armature))
            for k in range(6):
                m.dof_body.append(b)
                m.dof_parent.append(pdof)
                pdof = len(m.dof_body) - 1
                m.dof_armature.append(arm)
                axis = [0.0, 0.0, 0.0]
                axis[k % 3] = 1.0
                m.dof_axis.append(axis)
                m.dof_anchor.append([0.0, 0.0, 0.0])
            m.qpos0.extend(list(pos) + [1.0, 0.0, 0.0, 0.0])
            qadr += 7
        else:
            m.body_dofnum.append(len(joints))
            for j in joints:
                if j.get('type', 'hinge') != 'hinge':
                    raise NotImplementedError('joint type %s' % j.get('type'))
                m.joint_names.append(j.get('name'))
                rng = _vec(j.get('range', '0 0'), 2)
                m.jnt_range.append(list(np.deg2rad(rng)))
                m.dof_body.append(b)
                m.dof_parent.append(pdof)
                pdof = len(m.dof_body) - 1
                m.dof_armature.append(float(j.get('armature', def_armature)))
                axis = _vec(j.get('axis'), 3)
                m.dof_axis.append(list(axis / np.linalg.norm(axis)))
                m.dof_anchor.append(list(_vec(j.get('pos'), 3) - pos))
                m.qpos0.append(0.0)
                qadr += 1
        for child in elem.findall('body'):
            visit(child, b)

    for top in world.findall('body'):
        visit(top, -1)
    m.nbody = len(m.body_names)
    m.nv = len(m.dof_body)
    m.nq = qadr
    # actuators: motors in XML order; joint name -> dof index
    hinge_names = m.joint_names[1:] if m.body_dofnum[0] == 6 else m.joint_names
    first_hinge_dof = 6 if m.body_dofnum[0] == 6 else 0
    name2dof = {n: first_hinge_dof + i for i, n in enumerate(hinge_names)}
    act = root.find('actuator')
    for mot in (act.findall('motor') if act is not None else []):
        m.actuator_names.append(mot.get('name'))
        m.actuator_dof.append(name2dof[mot.get('joint')])
        m.actuator_gear.append(float(mot.get('gear', 1.0)))
    m.nu = len(m.actuator_names)
    return m


def inverse_weights(md):
    """(dof_invweight0 [nv], body_invweight0 [nbody][2]) as MuJoCo's compiler derives them (engine_setconst.c: set0): the
    diagonal of M^-1 at qpos0 per dof, and per body the mean diagonal of J M^-1 J^T for the translational / rotational
    Jacobian of its centre of mass.  At qpos0 every body frame of a coordinate="global" model is the world frame, so the
    mass matrix is a plain sum over bodies of J_b^T diag(m 1, I_b) J_b plus the armature."""
    nb, nv = md.nbody, md.nv
    xpos = np.zeros((nb, 3))
    for b in range(nb):
        p = md.body_parent[b]
        xpos[b] = np.asarray(md.body_pos[b]) + (xpos[p] if p >= 0 else 0.0)
    root_free = md.body_dofnum[0] == 6
    if root_free:
        xpos += np.asarray(md.qpos0[:3]) - xpos[0]
    com = xpos + np.asarray(md.body_ipos)
    Js = []
    M = np.diag(np.asarray(md.dof_armature, dtype=np.float64))
    for b in range(nb):
        J = np.zeros((6, nv))                           # rows: v_com (3), omega (3)
        i = md.body_dofadr[b] + md.body_dofnum[b] - 1
        while i >= 0:
            jb = md.dof_body[i]
            k = i - md.body_dofadr[jb]
            if md.body_dofnum[jb] == 6 and k < 3:
                J[k, i] = 1.0
            else:
                ax = np.asarray(md.dof_axis[i], dtype=np.float64)
                anc = xpos[jb] + np.asarray(md.dof_anchor[i])
                J[3:, i] = ax
                J[:3, i] = np.cross(ax, com[b] - anc)
            i = md.dof_parent[i]
        xx, yy, zz, xy, xz, yz = md.body_inertia[b]
        I6 = np.zeros((6, 6))
        I6[:3, :3] = np.eye(3) * md.body_mass[b]
        I6[3:, 3:] = [[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]]
        M += J.T @ I6 @ J
        Js.append(J)
    Mi = np.linalg.inv(M)
    body_iw = np.zeros((nb, 2))
    for b in range(nb):
        A = Js[b] @ Mi @ Js[b].T
        body_iw[b] = [np.trace(A[:3, :3]) / 3.0, np.trace(A[3:, 3:]) / 3.0]
    return np.ascontiguousarray(np.diag(Mi)), body_iw


def load_builtin(name='humanoid_1205_v1') -> ModelDesc:
    """Constants compiled from the reference asset by tools/compile_model.py (shipped as JSON so the
    GPU box, which has no /root/reference, does not need the XML)."""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    return ModelDesc.from_json(os.path.join(here, 'assets', name + '.model.json'))
