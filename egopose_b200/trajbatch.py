"""core/trajbatch.py:4-16 + ego_pose/core/trajbatch_ego.py:5-9 layout contract, device-resident.

Attributes ``states, actions, masks, next_states, rewards, exps, v_metas`` are numpy arrays exactly as the
reference's (row-major, episode-contiguous, env-major); they are materialised from the device tensors on
first access so that the fused update path, which reads ``.dev`` directly, pays no host round trip."""
import numpy as np

_FIELDS = ('states', 'actions', 'masks', 'next_states', 'rewards', 'exps', 'v_metas')


class TrajBatch:
    fields = _FIELDS[:6]

    def __init__(self, dev=None, horizon=None, host=None):
        self.dev = dev or {}            # name -> CUDA tensor
        self.horizon = horizon          # rows per environment (None for reference-format batches)
        self._host = dict(host or {})

    @classmethod
    def from_numpy(cls, **arrays):
        return cls(host=arrays)

    def __getattr__(self, name):
        if name in _FIELDS:
            if name not in self._host:
                if name not in self.dev or self.dev[name] is None:
                    raise AttributeError(name)
                a = self.dev[name].cpu().numpy()
                if name in ('masks', 'exps'):
                    a = a.astype(np.int64)          # reference stores python ints (agents/agent.py:60-61)
                self._host[name] = a
            return self._host[name]
        raise AttributeError(name)

    def to_host(self):
        for f in self.fields:
            if f in self.dev and self.dev[f] is not None:
                getattr(self, f)
        return self

    def __len__(self):
        if 'rewards' in self.dev:
            return int(self.dev['rewards'].shape[0])
        return int(self._host['rewards'].shape[0])


class TrajBatchEgo(TrajBatch):
    fields = _FIELDS
