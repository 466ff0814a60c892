"""Checkpoint format of ego_pose/ego_mimic.py:133-139 (save) and :57-65 (resume): a pickle of

    {'policy_dict', 'policy_vs_dict', 'value_dict', 'value_vs_dict', 'running_state'}

with CPU state-dicts (reference key names) and the ZFilter instance.  The pickle stream names the filter
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


def save_checkpoint(path, policy_net, policy_vs_net, value_net, value_vs_net, running_state):
    """ego_mimic.py:133-139"""
    cp = {'policy_dict': _cpu_state(policy_net), 'policy_vs_dict': _cpu_state(policy_vs_net),
          'value_dict': _cpu_state(value_net), 'value_vs_dict': _cpu_state(value_vs_net),
          'running_state': _to_reference_filter(running_state)}
    with open(path, 'wb') as f, _alias_modules():
        pickle.dump(cp, f)


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
