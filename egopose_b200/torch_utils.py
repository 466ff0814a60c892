"""utils/torch.py:1-158 helpers the training scripts call (context managers, lr setter, flat-param tools)."""
import numpy as np
import torch

tensor = torch.tensor
DoubleTensor = torch.DoubleTensor
FloatTensor = torch.FloatTensor
LongTensor = torch.LongTensor
ones = torch.ones
zeros = torch.zeros


def _dev(x):
    return x.device if hasattr(x, 'device') else next(x.parameters()).device


class _Restore:
    def __enter__(self):
        pass


class to_cpu(_Restore):
    """utils/torch.py:13-27: inside the context the models' tensors live on the CPU (checkpointing, ego_mimic.py:134).
    Implemented the way the PyTorch of the reference's era behaved - every parameter / buffer keeps its identity and only
    its ``.data`` is swapped, the original tensors come back on exit - so the flat parameter buffers of the fused agents
    (and the caller's optimizer) stay aliased.  The fused sampler reads the policy parameters where they are (GPU), so
    callers no longer need this around sample()."""

    def __init__(self, *models):
        self.models = [x for x in models if x is not None]
        self.saved = []
        for m in self.models:
            for t in list(m.parameters()) + list(m.buffers()):
                self.saved.append((t, t.data))
                t.data = t.data.to(torch.device('cpu'))

    def __exit__(self, *args):
        for t, d in self.saved:
            t.data = d
        return False


class to_device(_Restore):
    """utils/torch.py:30-44.  ego_mimic.py:66 calls it as a plain statement to place the nets for good (the object is
    dropped without leaving the context); as a context manager it moves them back on exit."""

    def __init__(self, device, *models):
        self.models = [x for x in models if x is not None]
        self.prev = [_dev(x) if list(x.parameters()) else device for x in self.models]
        for x in self.models:
            x.to(device)

    def __exit__(self, *args):
        for x, d in zip(self.models, self.prev):
            x.to(d)
        return False


class to_test(_Restore):
    def __init__(self, *models):
        self.models = [x for x in models if x is not None]
        self.prev = [x.training for x in self.models]
        for x in self.models:
            x.train(False)

    def __exit__(self, *args):
        for x, m in zip(self.models, self.prev):
            x.train(m)
        return False


class to_train(to_test):
    def __init__(self, *models):
        self.models = [x for x in models if x is not None]
        self.prev = [x.training for x in self.models]
        for x in self.models:
            x.train(True)


def batch_to(dst, *args):
    return [x.to(dst) if x is not None else None for x in args]


def set_optimizer_lr(optimizer, lr):
    for g in optimizer.param_groups:
        g['lr'] = lr


def get_flat_params_from(model):
    return torch.cat([p.data.view(-1) for p in model.parameters()])


def set_flat_params_to(model, flat_params):
    i = 0
    for p in model.parameters():
        n = p.numel()
        p.data.copy_(flat_params[i:i + n].view(p.size()))
        i += n


def filter_state_dict(state_dict, filter_keys):
    """utils/torch.py:153-158 (forecast warm start drops net.affine_layers.0 when the input width changes)"""
    for key in list(state_dict.keys()):
        for f_key in filter_keys:
            if f_key in key:
                del state_dict[key]
                break
