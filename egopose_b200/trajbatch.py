"""core/trajbatch.py:4-16 + ego_pose/core/trajbatch_ego.py:5-9 layout contract, device-resident.

Attributes ``states, actions, masks, next_states, rewards, exps, v_metas`` are numpy arrays exactly as the
reference's (row-major, episode-contiguous, env-major); they are materialised from the device tensors on
first access so that the fused update path, which reads ``.dev`` directly, pays no host round trip."""
import numpy as np

_FIELDS = ('states', 'actions', 'masks', 'next_states', 'rewards', 'exps', 'v_metas')
_CHUNK_BYTES = 128 << 20


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class TrajBatch:
    fields = _FIELDS[:6]

    def __init__(self, dev=None, horizon=None, host=None, pinned=None):
        self.dev = dev or {}            # name -> CUDA tensor
        self.horizon = horizon          # rows per environment (None for reference-format batches)
        self._host = dict(host or {})
        self._pinned = dict(pinned or {})   # name -> pinned CPU tensor backing the numpy view in _host
        self._events = {}               # name -> CUDA event of a device-to-host copy still in flight (to_host(stream=...))
        self._chunks = {}               # name -> [(row0, row1, event)] of that copy

    @classmethod
    def from_numpy(cls, **arrays):
        return cls(host=arrays)

    def __getattr__(self, name):
        if name in _FIELDS:
            ev = self.__dict__.get('_events', {}).pop(name, None)
            if ev is not None:
                ev.synchronize()        # the copy behind this array was issued asynchronously: wait for it (only it)
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

    def to_host(self, pool=None, stream=None):
        """materialise every field on the host.  With ``pool`` (dict name -> pinned CPU tensor, reused across
        iterations) the copies are asynchronous DMA into page-locked memory with one synchronise at the end;
        the numpy attributes are views of those buffers (valid until the next sample()).
        With ``stream`` (a CUDA side stream) nothing is synchronised here: the copies run on that stream in the order the
        update consumes the fields (``next_states``, which it never reads, last), every field gets its own event, a numpy
        attribute waits for its event on first access, and ``host_event(name)`` lets a consumer on another stream wait on
        the device instead (the re-upload of field k then overlaps the download of field k + 1: PCIe is full duplex)."""
        import torch
        if pool is None:
            for f in self.fields:
                if f in self.dev and self.dev[f] is not None:
                    getattr(self, f)
            return self
        order = [f for f in ('v_metas', 'states', 'actions', 'rewards', 'masks', 'exps', 'next_states') if f in self.fields]
        if stream is not None:
            stream.wait_stream(torch.cuda.current_stream())     # the rollout, and earlier uploads out of these pinned buffers
        with torch.cuda.stream(stream) if stream is not None else _null():
            for f in order:
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
                self._pinned[f] = buf
                if stream is None:
                    buf.copy_(t, non_blocking=True)
                    continue
                # large fields go down in row chunks with an event each, so that a consumer can start uploading the head of
                # the array while its tail is still on the way
                rows = int(t.shape[0])
                per = max(1, rows // max(1, (t.numel() * t.element_size()) // _CHUNK_BYTES)) if rows else 1
                chunks = []
                for r0 in range(0, max(rows, 1), per):
                    r1 = min(rows, r0 + per)
                    buf[r0:r1].copy_(t[r0:r1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    chunks.append((r0, r1, ev))
                self._chunks[f] = chunks
                self._events[f] = chunks[-1][2]
        if stream is None:
            torch.cuda.current_stream().synchronize()
        for f, buf in self._pinned.items():
            self._host[f] = buf.numpy()
        return self

    def host_chunks(self, name):
        """[(row0, row1, event)] of the pending download behind ``name`` (empty when complete or never streamed)"""
        return self._chunks.get(name, []) if name in self._events else []

    def host_event(self, name):
        """CUDA event of the still pending device-to-host copy behind ``name`` (None when the array is complete)"""
        return self._events.get(name)

    def wait_host(self):
        """block until every asynchronous device-to-host copy of this batch has landed"""
        for ev in list(self._events.values()):
            ev.synchronize()
        self._events.clear()
        return self

    def host_only(self):
        """the same batch as a reference-format object: numpy arrays only (views of the pinned buffers, pending copies and
        their events carried over), no device tensors - what a caller that built the batch from host data would pass"""
        b = type(self)(host=dict(self._host), horizon=self.horizon, pinned=dict(self._pinned))
        b._events = self._events            # shared: whoever waits first clears the entry for both
        b._chunks = self._chunks
        for f in self.fields:
            if f not in b._host and self.dev.get(f) is not None:
                b._host[f] = getattr(self, f)
        return b

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
