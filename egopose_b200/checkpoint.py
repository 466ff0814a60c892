"""Checkpoint format of ego_pose/ego_mimic.py:133-139 (save) and :57-65 (resume): a pickle of

    {'policy_dict', 'policy_vs_dict', 'value_dict', 'value_vs_dict', 'running_state'}

with CPU state-dicts (reference key names) and the ZFilter instance; optionally ("incl. optimizer / RNG
state, missing in the reference") three more keys that the reference's resume code (:57-65) never looks at, so such a file
still loads there: ``optimizer_policy`` / ``optimizer_value`` = torch.optim state dicts on the CPU (Adam m, v, step of the
flat buffers) and ``rng`` = {'seed', 'iteration'} of the rollout's counter-based Philox streams plus the host numpy / torch
generator states.  The pickle stream names the filter
classes ``utils.zfilter.ZFilter`` / ``utils.zfilter.RunningStat`` exactly like the reference does, so files
move between the two code bases in both directions; on load those names resolve to egopose_b200.zfilter
without requiring the reference (or the compat shim) on sys.path.
"""
import pickle

import torch

from . import zfilter

_ALIASES = {('utils.zfilter', 'ZFilter'): zfilter.ZFilter, ('utils.zfilter', 'RunningStat'): zfilter.RunningStat}


class _alias_modules:
    """while dumping, make ``utils.zfilter`` resolve to the reference-named filter classes so that pickle
    records exactly the global names the reference's own checkpoints carry"""

    def __enter__(self):
        import sys
        import types
        self.saved = {k: sys.modules.get(k) for k in ('utils', 'utils.zfilter')}
        mod = types.ModuleType('utils.zfilter')
        mod.ZFilter, mod.RunningStat = _RefNamedZFilter, _RefNamedRunningStat
        pkg = self.saved['utils'] or types.ModuleType('utils')
        self.had_attr = getattr(pkg, 'zfilter', None)
        pkg.zfilter = mod
        sys.modules['utils'], sys.modules['utils.zfilter'] = pkg, mod

    def __exit__(self, *exc):
        import sys
        for k, v in self.saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        if self.saved['utils'] is not None:
            if self.had_attr is None:
                try:
                    delattr(self.saved['utils'], 'zfilter')
                except AttributeError:
                    pass
            else:
                self.saved['utils'].zfilter = self.had_attr
        return False


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _ALIASES:
            return _ALIASES[(module, name)]
        return super().find_class(module, name)


def _cpu_state(net):
    return {k: v.detach().cpu().clone() for k, v in net.state_dict().items()} if net is not None else {}


def _cpu_tree(x):
    if torch.is_tensor(x):
        return x.detach().cpu().clone()
    if isinstance(x, dict):
        return {k: _cpu_tree(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_cpu_tree(v) for v in x)
    return x


def save_checkpoint(path, policy_net, policy_vs_net, value_net, value_vs_net, running_state, optimizer_policy=None,
                    optimizer_value=None, agent=None):
    """ego_mimic.py:133-139 (+ optional optimizer / RNG state under keys the reference ignores).  ``agent``: its
    env seed and rollout counter define every Philox stream of the fused sampler."""
    cp = {'policy_dict': _cpu_state(policy_net), 'policy_vs_dict': _cpu_state(policy_vs_net),
          'value_dict': _cpu_state(value_net), 'value_vs_dict': _cpu_state(value_vs_net),
          'running_state': _to_reference_filter(running_state)}
    if optimizer_policy is not None:
        cp['optimizer_policy'] = _cpu_tree(optimizer_policy.state_dict())
    if optimizer_value is not None:
        cp['optimizer_value'] = _cpu_tree(optimizer_value.state_dict())
    if agent is not None:
        import numpy as np
        cp['rng'] = {'seed': int(agent.env._seed), 'iteration': int(agent.iteration), 'numpy': np.random.get_state(),
                     'torch': torch.get_rng_state()}
    with open(path, 'wb') as f, _alias_modules():
        pickle.dump(cp, f)


def restore_training_state(cp, optimizer_policy=None, optimizer_value=None, agent=None):
    """the optional part of a checkpoint: Adam moments / step into the caller's optimizers (the fused update adopts them
    when it builds its flat buffers, agent._FlatNet), Philox seed / iteration and host generator states"""
    for opt, key in ((optimizer_policy, 'optimizer_policy'), (optimizer_value, 'optimizer_value')):
        if opt is not None and cp.get(key) is not None:
            opt.load_state_dict(cp[key])
    if agent is not None and cp.get('rng') is not None:
        import numpy as np
        agent.env._seed = cp['rng']['seed']
        agent.iteration = cp['rng']['iteration']
        np.random.set_state(cp['rng']['numpy'])
        torch.set_rng_state(cp['rng']['torch'])
        if getattr(agent, '_nets', None) is not None:       # flat buffers already exist: re-adopt the optimizer state
            for flat in (agent._pf, agent._vf):
                for p, a, b in zip(flat.params, flat.offsets[:-1], flat.offsets[1:]):
                    st = flat.optimizer.state.get(p, {})
                    if 'exp_avg' in st and st['exp_avg'].data_ptr() != flat.m[a:b].data_ptr():
                        flat.m[a:b].copy_(st['exp_avg'].reshape(-1)); flat.v[a:b].copy_(st['exp_avg_sq'].reshape(-1))
                        flat.step = int(st['step'])
                        st['exp_avg'], st['exp_avg_sq'] = flat.m[a:b].view(p.shape), flat.v[a:b].view(p.shape)


def load_checkpoint(path, policy_net=None, policy_vs_net=None, value_net=None, value_vs_net=None):
    """ego_mimic.py:57-65; returns (checkpoint dict, running_state)"""
    with open(path, 'rb') as f:
        cp = _Unpickler(f).load()
    for net, key in ((policy_net, 'policy_dict'), (policy_vs_net, 'policy_vs_dict'), (value_net, 'value_dict'),
                     (value_vs_net, 'value_vs_dict')):
        if net is not None and cp.get(key):
            with torch.no_grad():
                for k, v in net.state_dict().items():
                    v.copy_(cp[key][k])         # in place: keeps the flat-buffer aliasing of the fused update
    return cp, _from_any_filter(cp.get('running_state'))


def _to_reference_filter(rs):
    """instances of a class whose pickled name is utils.zfilter.ZFilter (works with or without the shim)"""
    if rs is None:
        return None
    z = _RefNamedZFilter.__new__(_RefNamedZFilter)
    z.__dict__.update(rs.__dict__)
    r = _RefNamedRunningStat.__new__(_RefNamedRunningStat)
    r.__dict__.update(rs.rs.__dict__)
    z.rs = r
    return z


def _from_any_filter(rs):
    if rs is None:
        return None
    z = zfilter.ZFilter(rs.rs._M.shape, demean=rs.demean, destd=rs.destd, clip=rs.clip)
    z.rs._n, z.rs._M, z.rs._S = rs.rs._n, rs.rs._M.copy(), rs.rs._S.copy()
    return z


class _RefNamedZFilter(zfilter.ZFilter):
    pass


class _RefNamedRunningStat(zfilter.RunningStat):
    pass


# the pickle stream records __module__ / __qualname__ of the class: name them like the reference's
_RefNamedZFilter.__module__, _RefNamedZFilter.__qualname__, _RefNamedZFilter.__name__ = 'utils.zfilter', 'ZFilter', 'ZFilter'
_RefNamedRunningStat.__module__, _RefNamedRunningStat.__qualname__, _RefNamedRunningStat.__name__ = 'utils.zfilter', 'RunningStat', 'RunningStat'
