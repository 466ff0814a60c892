"""core/trajbatch.py:4-16 + ego_pose/core/trajbatch_ego.py:5-9 layout contract, device-resident.

Attributes ``states, actions, masks, next_states, rewards, exps, v_metas`` are numpy arrays exactly as the
reference's (row-major, episode-contiguous, env-major); they are materialised from the device tensors on
first access so that the fused update path, which reads ``.dev`` directly, pays no host round trip."""
import numpy as np

_FIELDS = ('states', 'actions', 'masks', 'next_states', 'rewards', 'exps', 'v_metas')


class TrajBatch:
    fields = _FIELDS[:6]

    def __init__(self, dev=None, horizon=None, host=None, pinned=None):
        self.dev = dev or {}            # name -> CUDA tensor
        self.horizon = horizon          # rows per environment (None for reference-format batches)
        self._host = dict(host or {})
        self._pinned = dict(pinned or {})   # name -> pinned CPU tensor backing the numpy view in _host

    @classmethod
    def from_numpy(cls, **arrays):
        return cls(host=arrays)

    def __getattr__(self, name):
        if name in _FIELDS:
            if name not in self._host:
                if name not in self.dev or self.dev[name] is None:
                    raise AttributeError('%s was not recorded by this rollout%s' % (
                        name, ' (set agent.keep_next_states = True or sample(to_host=True))' if name == 'next_states' else ''))
                a = self.dev[name].cpu().numpy()
                if name in ('masks', 'exps'):
                    a = a.astype(np.int64)          # reference stores python ints (agents/agent.py:60-61)
                self._host[name] = a
            return self._host[name]
        raise AttributeError(name)

    def to_host(self, pool=None):
        """materialise every field on the host.  With ``pool`` (dict name -> pinned CPU tensor, reused across
        iterations) the copies are asynchronous DMA into page-locked memory with one synchronise at the end;
        the numpy attributes are views of those buffers (valid until the next sample())."""
        import torch
        if pool is None:
            for f in self.fields:
                if f in self.dev and self.dev[f] is not None:
                    getattr(self, f)
            return self
        for f in self.fields:
            t = self.dev.get(f)
            if t is None:
                continue
            want = torch.int64 if f in ('masks', 'exps') else t.dtype
            buf = pool.get(f)
            if buf is None or buf.shape != t.shape or buf.dtype != want:
                buf = torch.empty(t.shape, dtype=want, pin_memory=True)
                pool[f] = buf
            if want != t.dtype:
                t = t.to(want)              # device-side cast, then DMA
            buf.copy_(t, non_blocking=True)
            self._pinned[f] = buf
        torch.cuda.current_stream().synchronize()
        for f, buf in self._pinned.items():
            self._host[f] = buf.numpy()
        return self

    def pinned(self, name):
        """the pinned tensor behind attribute ``name`` if it is one of ours and still the exposed array"""
        buf = self._pinned.get(name)
        if buf is not None and name in self._host and self._host[name].ctypes.data == buf.data_ptr():
            return buf
        return None

    def __len__(self):
        if 'rewards' in self.dev:
            return int(self.dev['rewards'].shape[0])
        return int(self._host['rewards'].shape[0])


class TrajBatchEgo(TrajBatch):
    fields = _FIELDS
