"""TEST INFRASTRUCTURE ONLY - never imported by the product path.

Import harness for the *unmodified* reference sources at /root/reference (build container only; the
GPU box has no /root/reference).  The reference star-imports GUI / simulator packages that are not
installed (utils/__init__.py:1-7, utils/tools.py:6-11, envs/common/mujoco_env.py:3-8); they are
replaced by empty stub modules carrying a real ``__spec__`` so the PPO half, the reward function and
the math helpers can be imported and executed to produce golden vectors (tests/golden/make_golden.py).
"""
import importlib.machinery
import os
import sys
import types

REF = os.environ.get('EGOPOSE_REFERENCE', '/root/reference')

_STUBS = ['gym', 'gym.envs', 'gym.envs.mujoco', 'gym.envs.mujoco.mujoco_env', 'gym.utils', 'gym.spaces',
          'OpenGL', 'OpenGL.GL', 'glfw', 'tensorflow', 'mujoco_py', 'mujoco_py.builder', 'mujoco_py.generated',
          'mujoco_py.utils', 'mujoco_py.functions', 'pyautogui', 'imageio', 'scipy.misc', 'cv2', 'PIL', 'PIL.Image']


def available():
    return os.path.isdir(os.path.join(REF, 'agents'))


def install():
    """Put the reference on sys.path with stubbed third-party GUI/sim modules. Idempotent."""
    if not available():
        raise RuntimeError('reference sources not found at %s' % REF)
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
        mod.__path__ = []
        sys.modules[name] = mod
        if '.' in name:
            parent, child = name.rsplit('.', 1)
            setattr(sys.modules[parent], child, mod)
    sys.modules['gym.envs.mujoco.mujoco_env'].MujocoEnv = object
    sys.modules['gym'].error = types.SimpleNamespace()
    sys.modules['gym.utils'].seeding = types.SimpleNamespace()
    sys.modules['OpenGL'].GL = sys.modules['OpenGL.GL']
    if REF not in sys.path:
        sys.path.insert(0, REF)
