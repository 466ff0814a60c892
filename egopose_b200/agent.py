"""Host mirrors of the reference agents on the PPO hot path, same constructor keywords and methods:

  Agent      agents/agent.py:8-122       sample(min_batch_size) -> (TrajBatch, LoggerRL), set_noise_rate, hooks
  AgentPG    agents/agent_pg.py:7-57     update_params(batch) -> seconds, update_value / update_policy
  AgentPPO   agents/agent_ppo.py:6-65    clipped surrogate, grad-norm clip, full-batch epochs
  AgentEgo   ego_pose/core/agent_ego.py  video-context nets, TrajBatchEgo, v_metas

What changes underneath (B200-first):
  * sample() launches ONE fused rollout kernel for E independent environments x T steps
    (egp_rollout_f64) instead of forking `num_threads` Python workers; `num_threads` is accepted and
    ignored (document: the parallelism is the environment batch, `num_envs`).
  * update_params() keeps the trajbatch on the device, runs the GAE scan / loss / clip+Adam kernels and
    dense layers with explicit backward GEMMs; parameters and Adam moments live in flat buffers that the
    caller's nn.Module parameters and torch.optim.Adam state alias, so state_dict()/checkpoints keep the
    reference's format (ego_mimic.py:133-139) and set_optimizer_lr (ego_mimic.py:96) keeps working.
  * with torch.distributed initialised (one process per GPU) environments are sharded across ranks and
    each PPO epoch does one all-reduce of the flat [value | policy] gradient.
There is no CPU fallback: every numeric step is a kernel of libegopose_b200.so or a cuBLAS GEMM.
"""
import math
import os
import time

import numpy as np
import torch

from . import lib
from .logger_rl import LoggerRL
from .nets import FrameContext, VideoForecastNet, VideoStateNet, trunk_ok
from .trajbatch import TrajBatch, TrajBatchEgo


from . import dist_utils


def _dist():
    return dist_utils.group()


class _FlatNet:
    """Flat parameter / gradient / Adam-moment storage for one optimizer; the module parameters and the
    torch.optim.Adam state become views of it."""

    def __init__(self, named, optimizer, device, grad=None):
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        self.optimizer = optimizer
        sizes = [p.numel() for p in self.params]
        self.offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        n = int(self.offsets[-1])
        f64 = dict(dtype=torch.float64, device=device)
        self.flat = torch.empty(n, **f64)
        self.grad = torch.zeros(n, **f64) if grad is None else grad     # may be a slice of a buffer shared with another net
        self.m = torch.zeros(n, **f64)
        self.v = torch.zeros(n, **f64)
        self.norm2 = torch.zeros(1, **f64)
        self.step = 0
        for p, a, b in zip(self.params, self.offsets[:-1], self.offsets[1:]):
            self.flat[a:b].copy_(p.data.reshape(-1).to(**f64))
            p.data = self.flat[a:b].view(p.shape)
        if optimizer is not None:
            for p, a, b in zip(self.params, self.offsets[:-1], self.offsets[1:]):
                st = optimizer.state[p]
                if 'exp_avg' in st:         # resume from an existing Adam state
                    self.m[a:b].copy_(st['exp_avg'].reshape(-1))
                    self.v[a:b].copy_(st['exp_avg_sq'].reshape(-1))
                    self.step = int(st['step'])
                st['exp_avg'] = self.m[a:b].view(p.shape)
                st['exp_avg_sq'] = self.v[a:b].view(p.shape)
                st['step'] = torch.tensor(float(self.step))

    def realias(self, named=None):
        """Re-establish parameter <-> flat-buffer aliasing after the caller moved the modules (a stock module.to(cpu) /
        .to(cuda) round trip around checkpointing, ego_mimic.py:134): current PyTorch then REPLACES the Parameter objects
        (CPU and CUDA tensors are not shallow-copy compatible), older versions rebind p.data.  ``named`` is the current
        (name, parameter) list of the modules; the module tensors are the source of truth at that point (the caller may
        also have loaded a state dict).  The caller's optimizer is re-pointed to the new objects, its state follows."""
        moved = False
        if named is not None:
            cur = dict(named)
            for i, (n, old) in enumerate(zip(self.names, self.params)):
                new = cur.get(n)
                if new is None or new is old:
                    continue
                self.params[i] = new
                moved = True
                if self.optimizer is not None:
                    for g in self.optimizer.param_groups:
                        g['params'] = [new if q is old else q for q in g['params']]
                    if old in self.optimizer.state:
                        self.optimizer.state[new] = self.optimizer.state.pop(old)
        for p, a, b in zip(self.params, self.offsets[:-1], self.offsets[1:]):
            if p.data.data_ptr() != self.flat[a:b].data_ptr() or p.data.device != self.flat.device:
                self.flat[a:b].copy_(p.data.reshape(-1).to(self.flat.device, torch.float64))
                p.data = self.flat[a:b].view(p.shape)
                moved = True
        return moved

    def view(self, buf, name):
        i = self.names.index(name)
        return buf[self.offsets[i]:self.offsets[i + 1]].view(self.params[i].shape)

    def hyper(self):
        g = self.optimizer.param_groups[0]
        b1, b2 = g.get('betas', (0.9, 0.999))
        return g['lr'], b1, b2, g.get('eps', 1e-8)

    def adam(self, max_norm=0.0):
        lr, b1, b2, eps = self.hyper()
        self.step += 1
        if max_norm and max_norm > 0:
            lib.sumsq(self.grad, self.norm2)
        lib.adam_step(self.flat, self.grad, self.m, self.v, lr, b1, b2, eps, self.step, max_norm or 0.0, self.norm2)
        for p in self.params:
            self.optimizer.state[p]['step'].fill_(self.step)


class _Trunk:
    """Two-hidden-layer relu trunk + linear head: explicit forward / backward with reusable activation
    buffers (models/mlp.py + policy_gaussian.py:19-24 / critic.py:15-18).  GEMMs are cuBLAS (torch.mm with
    out=), bias+relu, relu-backward and bias gradients are kernels of this library."""

    def __init__(self, flat, head):
        self.f = flat
        self.names = ['net.affine_layers.0', 'net.affine_layers.1', head]
        self.buf = {}
        self.cap = {}

    def W(self, i):
        return self.f.view(self.f.flat, self.names[i] + '.weight')

    def b(self, i):
        return self.f.view(self.f.flat, self.names[i] + '.bias')

    def gW(self, i):
        return self.f.view(self.f.grad, self.names[i] + '.weight')

    def gb(self, i):
        return self.f.view(self.f.grad, self.names[i] + '.bias')

    def weights(self):
        return (self.W(0), self.b(0), self.W(1), self.b(1), self.W(2), self.b(2))

    def grads(self):
        return (self.gW(0), self.gb(0), self.gW(1), self.gb(1), self.gW(2), self.gb(2))

    def dims(self):
        return (self.W(0).shape[1], self.W(0).shape[0], self.W(1).shape[0], self.W(2).shape[0])

    def _buf(self, key, shape, like):
        """reusable activation buffer: grows to the largest row count seen, hands out a [:n] view"""
        t = self.cap.get(key)
        if t is None or t.shape[0] < shape[0] or tuple(t.shape[1:]) != tuple(shape[1:]):
            t = torch.empty(shape, dtype=torch.float64, device=like.device)
            self.cap[key] = t
        v = t[:shape[0]]
        self.buf[key] = v
        return v

    def forward(self, x):
        n = x.shape[0]
        W0, W1, W2 = self.W(0), self.W(1), self.W(2)
        h1 = self._buf('h1', (n, W0.shape[0]), x)
        h2 = self._buf('h2', (n, W1.shape[0]), x)
        y = self._buf('y', (n, W2.shape[0]), x)
        torch.mm(x, W0.t(), out=h1)
        lib.bias_relu_(h1, self.b(0))
        torch.mm(h1, W1.t(), out=h2)
        lib.bias_relu_(h2, self.b(1))
        torch.addmm(self.b(2), h2, W2.t(), out=y)
        self.x = x
        return y

    def backward(self, dy, dx_cols=0):
        """writes the six gradient views of the flat buffer; with dx_cols > 0 also returns dL/dx[:, :dx_cols]
        (the gradient reaching the video-context columns of the input)"""
        x, h1, h2 = self.x, self.buf['h1'], self.buf['h2']
        n = x.shape[0]
        torch.mm(dy.t(), h2, out=self.gW(2))
        lib.colsum(dy, self.gb(2))
        dh2 = self._buf('dh2', h2.shape, x)
        torch.mm(dy, self.W(2), out=dh2)
        lib.relu_bwd_colsum_(dh2, h2, self.gb(1))
        torch.mm(dh2.t(), h1, out=self.gW(1))
        dh1 = self._buf('dh1', h1.shape, x)
        torch.mm(dh2, self.W(1), out=dh1)
        lib.relu_bwd_colsum_(dh1, h1, self.gb(0))
        torch.mm(dh1.t(), x, out=self.gW(0))
        if dx_cols > 0:
            dx = self._buf('dx', (n, dx_cols), x)
            torch.mm(dh1, self.W(0)[:, :dx_cols], out=dx)
            return dx
        return None


class _NetInput:
    """Dense-layer input of one net.  Constant tensor (identity / per-frame table), or cat(train_context(), states)
    rebuilt on every forward when a VideoStateNet produces the context (trans_policy / trans_value,
    agent_ego.py:28-32): x() hands the trunk a contiguous buffer, backward() feeds dL/dctx to torch autograd so the
    (Bi)LSTM parameters receive gradients, which are then copied into the flat gradient buffer."""

    def __init__(self, x_const=None, vs=None, states=None, flat=None):
        self.x_const, self.vs, self.flat = x_const, vs, flat
        self.ctx = None
        if vs is not None:
            self.states = states
            self.forecast = isinstance(vs, VideoForecastNet)
            if self.forecast:           # the whole trunk input is produced by the net: cat(v_out, s_net(states))
                self.cd = vs.out_dim
                self.xbuf = torch.empty((states.shape[0], self.cd), dtype=states.dtype, device=states.device)
            else:
                self.cd = vs.v_hdim
                self.xbuf = torch.empty((states.shape[0], self.cd + states.shape[1]), dtype=states.dtype, device=states.device)
                self.xbuf[:, self.cd:].copy_(states)
            self.vs_params = [(n, p) for n, p in vs.named_parameters() if p.requires_grad]

    @property
    def learned(self):
        return self.vs is not None

    def x(self, grad=True):
        if self.vs is None:
            return self.x_const
        with torch.set_grad_enabled(grad):
            self.ctx = self.vs.train_context(self.states) if self.forecast else self.vs.train_context()
        self.xbuf[:, :self.cd].copy_(self.ctx.detach())
        return self.xbuf

    def backward(self, dctx):
        for _, p in self.vs_params:
            p.grad = None
        self.ctx.backward(dctx)
        for n, p in self.vs_params:
            self.flat.view(self.flat.grad, 'vs.' + n).copy_(p.grad)
            p.grad = None
        self.ctx = None


class Agent:
    def __init__(self, env, policy_net, value_net, dtype, device, custom_reward=None, mean_action=False, render=False,
                 running_state=None, num_threads=1, num_envs=None, horizon=None):
        self.env = env
        self.policy_net = policy_net
        self.value_net = value_net
        self.dtype = dtype
        self.device = device
        self.custom_reward = custom_reward      # the reward is the kernel's quat_v3 (reward_function.py:4-60)
        self.mean_action = mean_action
        self.running_state = running_state
        self.render = render
        self.num_threads = num_threads          # accepted for API parity; parallelism = num_envs
        self.num_envs = num_envs
        self.horizon = horizon
        self.noise_rate = 1.0
        self.traj_cls = TrajBatch
        self.logger_cls = LoggerRL
        self.sample_modules = [policy_net]
        self.update_modules = [policy_net, value_net]
        self.iteration = 0
        self._out = {}
        self._host_pool = {}        # pinned host staging buffers reused by sample(to_host=True)
        self.keep_next_states = False
        if dtype != torch.float64:
            raise lib.EgpError('the fused path computes in float64 like the reference (ego_mimic.py:31-32)')
        if render:
            raise lib.EgpError('rendering is out of scope of the fused path')

    # hooks kept for API parity (agents/agent.py:78-85,113-122)
    def pre_episode(self):
        return

    def pre_sample(self):
        return

    def push_memory(self, memory, state, action, mask, next_state, reward, exp):
        raise lib.EgpError('trajectories are written by the rollout kernel; push_memory is not used')

    def trans_policy(self, states):
        return states

    def trans_value(self, states):
        return states

    def set_noise_rate(self, noise_rate):
        self.noise_rate = noise_rate

    def _policy_weights(self):
        p = self.policy_net
        if not trunk_ok(p.net):
            raise lib.EgpError('fused path needs the two-hidden-layer relu MLP trunk')
        L = p.net.affine_layers
        w = dict(W1=L[0].weight.data, b1=L[0].bias.data, W2=L[1].weight.data, b2=L[1].bias.data,
                 W3=p.action_mean.weight.data, b3=p.action_mean.bias.data, log_std=p.action_log_std.data.view(-1))
        for k, t in w.items():
            if not t.is_cuda or t.dtype != torch.float64:
                raise lib.EgpError('policy parameter %s must be a CUDA float64 tensor (move the nets to the GPU)' % k)
        return {k: t.contiguous() for k, t in w.items()}

    def _rollout_context(self):
        """(ctx table, win_off) handed to the rollout kernel, or (None, None) for the table uploaded with the experts"""
        return None, None

    def _rollout_extra(self):
        """extra keyword arguments of lib.Model.rollout (state LSTM, constant per-window context)"""
        return {}

    def _zf(self):
        rs = self.running_state
        if rs is None:
            return None, None, 0.0
        if rs.rs.n < 2:
            return None, None, rs.clip or 0.0
        dev = self.policy_net.action_mean.weight.device
        mean = torch.as_tensor(rs.rs.mean, dtype=torch.float64, device=dev)
        std = torch.as_tensor(rs.rs.std, dtype=torch.float64, device=dev)
        return mean, std, rs.clip or 0.0

    def plan(self, min_batch_size):
        """(E, T): T = horizon or cfg.env_episode_len, E = num_envs or ceil(min_batch / T)"""
        T = int(self.horizon or self.env.cfg.env_episode_len)
        E = int(self.num_envs or math.ceil(min_batch_size / T))
        return E, T

    def sample(self, min_batch_size, to_host=None, parity=None):
        """agents/agent.py:87-111.  ``parity`` may carry pre-drawn eps / reset_take / reset_start / mean_flag
        device tensors (SURVEY 7 'RNG parity'); otherwise noise and resets come from in-kernel Philox.

        The returned batch is DEVICE-RESIDENT and lazy (``to_host=None``, the default): ``batch.states`` etc. are numpy
        arrays exactly as the reference's, materialised on first access; ``update_params(batch)`` reads the device tensors
        directly, and re-uploads an array only if the caller touched it.  ``to_host=True`` copies every field eagerly into
        pinned host buffers (asynchronous DMA, one synchronise).  ``next_states`` - never read by the update
        (agent_ego.py:37-42) - is recorded only with ``to_host=True`` or ``agent.keep_next_states = True``.
        The batch aliases buffers that the NEXT sample() overwrites (both the device tensors and the pinned host copies):
        a batch kept across iterations must be copied by the caller; the reference returns independent arrays."""
        t_start = time.time()
        self.pre_sample()
        if getattr(self, '_d2h_done', None) is not None:
            torch.cuda.current_stream().wait_event(self._d2h_done)     # the last batch's downloads read the buffers this rollout overwrites
            self._d2h_done = None
        E, T = self.plan(min_batch_size)
        w = self._policy_weights()
        rs = self.running_state
        if rs is not None and rs.rs.n < 2:
            # first rollout: statistics from one warm-up rollout of the same size (frozen-stats deviation)
            wctx, wwin = self._rollout_context()
            warm = self.env.kernel.rollout(w, E, min(T, 8), self.env.cfg.env_episode_len, self.env.cfg.fr_margin,
                                           fix_head_lb=self.env.fix_head_lb, noise_rate=self.noise_rate,
                                           seed=self.env._seed, iteration=2 ** 40 + self.iteration, zf_clip=0.0,
                                           want_next=False, ctx=wctx, win_off=wwin, **self._rollout_extra())
            self._merge_obs(warm['raw_obs'])
        zm, zs, clip = self._zf()
        p = parity or {}
        ctx, win_off = self._rollout_context()
        out = self.env.kernel.rollout(
            w, E, T, self.env.cfg.env_episode_len, self.env.cfg.fr_margin, end_reward=self.env.end_reward,
            fix_head_lb=self.env.fix_head_lb, noise_rate=self.noise_rate, mean_action=self.mean_action,
            zf_mean=zm, zf_std=zs, zf_clip=clip, seed=self.env._seed, iteration=self.iteration,
            eps=p.get('eps'), reset_take=p.get('reset_take'), reset_start=p.get('reset_start'),
            mean_flag=p.get('mean_flag'), want_next=bool(to_host) or self.keep_next_states, want_raw=rs is not None, out=self._out,
            ctx=ctx, win_off=win_off, **self._rollout_extra())
        self.iteration += 1
        if rs is not None:
            self._merge_obs(out['raw_obs'])
        dev = {k: out.get(k) for k in self.traj_cls.fields}
        if not (to_host or self.keep_next_states):
            dev['next_states'] = None       # a buffer left from an earlier to_host=True rollout would be stale
        batch = self.traj_cls(dev=dev, horizon=T)
        lg = out['logger']
        lg = dist_utils.reduce_logger_(lg.clone(), (lib.LOG['MIN_C_REWARD'], lib.LOG['MIN_EPISODE_REWARD']),
                                       (lib.LOG['MAX_C_REWARD'], lib.LOG['MAX_EPISODE_REWARD']))
        logger = self.logger_cls.from_device(lg.cpu().numpy())     # D2H of 16 doubles: the rollout's sync point
        if to_host:
            # downloads on a side stream, one event per field (TrajBatch.to_host): update_params() re-uploads field k while
            # field k + 1 is still coming down, numpy readers wait for the array they touch; EGP_D2H_STREAM=0: one blocking pass
            if os.environ.get('EGP_D2H_STREAM', '1') != '0':
                if getattr(self, '_copy_stream', None) is None:
                    self._copy_stream = torch.cuda.Stream(device=lg.device)
                batch.to_host(self._host_pool, stream=self._copy_stream)
                self._d2h_done = torch.cuda.Event()
                self._d2h_done.record(self._copy_stream)
            else:
                batch.to_host(self._host_pool)
        logger.sample_time = time.time() - t_start
        return batch, logger

    def _merge_obs(self, raw):
        rs = self.running_state.rs
        shift = torch.as_tensor(rs.mean if rs.n > 0 else np.zeros(rs.shape), dtype=torch.float64, device=raw.device)
        mo = lib.col_moments(raw, shift)
        n_b = raw.shape[0]
        d = _dist()
        if d is not None:
            d.all_reduce(mo)
            n_b *= d.get_world_size()
        mo = mo.cpu().numpy()
        S = raw.shape[1]
        rs.merge_moments(n_b, mo[:S], mo[S:], shift.cpu().numpy())


class AgentPG(Agent):
    def __init__(self, gamma=0.99, tau=0.95, optimizer_policy=None, optimizer_value=None, opt_num_epochs=1,
                 value_opt_niter=1, gemm=None, oz_slices=None, **kwargs):
        super().__init__(**kwargs)
        self.gamma = gamma
        self.tau = tau
        self.optimizer_policy = optimizer_policy
        self.optimizer_value = optimizer_value
        self.opt_num_epochs = opt_num_epochs
        self.value_opt_niter = value_opt_niter
        self._nets = None
        self.last_info = {}
        # dense layers of the update: 'ozaki' = float64 on the int8 tensor cores (csrc/ozaki.cu, csrc/oz_mlp.cu),
        # 'cublas' = cuBLAS DGEMM.
        self.gemm = gemm or os.environ.get('EGP_GEMM', 'ozaki')
        self.oz_slices = int(oz_slices or os.environ.get('EGP_OZ_SLICES', 6))
        self.oz_chunk_waves = int(os.environ.get('EGP_OZ_CHUNK_WAVES', 8))
        if self.gemm not in ('ozaki', 'cublas'):
            raise lib.EgpError("gemm must be 'ozaki' or 'cublas'")
        self._ozs = {}
        self._xcaches = {}

    # ---- flat storage ------------------------------------------------------------------------------
    def _named_params(self):
        """(policy list, value list) of (name, parameter): the nets plus the video-context nets the optimizers also own
        (ego_mimic.py:68-69); the grad-norm clip spans policy_net and policy_vs_net jointly"""
        pol = [(n, p) for n, p in self.policy_net.named_parameters() if p.requires_grad]
        val = [(n, p) for n, p in self.value_net.named_parameters() if p.requires_grad]
        for lst, vs in ((pol, getattr(self, 'policy_vs_net', None)), (val, getattr(self, 'value_vs_net', None))):
            if isinstance(vs, (VideoStateNet, VideoForecastNet)):
                lst += [('vs.' + n, p) for n, p in vs.named_parameters() if p.requires_grad]
        return pol, val

    def _setup(self):
        if self._nets is not None:
            pol, val = self._named_params()
            self._pf.realias(pol)
            self._vf.realias(val)
            return
        for opt in (self.optimizer_policy, self.optimizer_value):
            if not isinstance(opt, torch.optim.Adam):
                raise lib.EgpError('the fused update implements torch.optim.Adam (ego_mimic.py:70-77)')
            g = opt.param_groups[0]
            if g.get('weight_decay', 0) or g.get('amsgrad', False):
                raise lib.EgpError('weight_decay / amsgrad are not supported by the fused Adam')
        dev = self.policy_net.action_mean.weight.device
        pol, val = self._named_params()
        if not (trunk_ok(self.policy_net.net) and trunk_ok(self.value_net.net)):
            raise lib.EgpError('fused path needs two-hidden-layer relu MLP trunks')
        # ONE flat [value | policy] gradient buffer: with several ranks every PPO epoch all-reduces it once
        nv = sum(p.numel() for _, p in val)
        npol = sum(p.numel() for _, p in pol)
        self._gflat = torch.zeros(nv + npol, dtype=torch.float64, device=dev)
        # With several ranks the flat gradient lives in this rank's peer-memory exchange block (csrc/p2p.cu): the
        # per-epoch sum is one kernel that reads the peers' blocks over NVLink, no NCCL call.  EGP_GRAD_EXCHANGE=nccl
        # keeps torch.distributed's all-reduce; the choice is collective (every rank must have mapped every peer).
        self._peer = None
        d = _dist()
        if d is not None and d.get_world_size() > 1 and os.environ.get('EGP_GRAD_EXCHANGE', 'p2p') != 'nccl':
            ok, peer = 1, None
            try:
                peer = lib.PeerComm(nv + npol, dev, d)
            except Exception as exc:        # noqa: BLE001  (no peer access / IPC refused: decided together below)
                ok, self._peer_error = 0, str(exc)
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            d.all_reduce(flag, op=d.ReduceOp.MIN)
            if int(flag.item()) == 1:
                self._peer = peer
                self._gflat = peer.src
                self._gflat.zero_()
            elif peer is not None:
                peer.close()
        self._pf = _FlatNet(pol, self.optimizer_policy, dev, grad=self._gflat[nv:])
        self._vf = _FlatNet(val, self.optimizer_value, dev, grad=self._gflat[:nv])
        self._pt = _Trunk(self._pf, 'action_mean')
        self._vt = _Trunk(self._vf, 'value_head')
        self._learn_std = 'action_log_std' in self._pf.names
        self._scal = torch.zeros(4, dtype=torch.float64, device=dev)
        self._nets = True

    # ---- pieces of update_params ---------------------------------------------------------------------
    def _device_batch(self, batch):
        dev = self.policy_net.action_mean.weight.device
        if getattr(batch, 'dev', None) and batch.dev.get('states') is not None and 'states' not in batch._host:
            b = batch.dev
            return b['states'], b['actions'], b['rewards'], b['masks'], b['exps'], b.get('v_metas'), batch.horizon
        # reference-format host batch (agents/agent_pg.py:43-47): async H2D straight from the pinned buffers
        # sample() handed out, else through a reusable pinned staging buffer
        def up(name, dtype=torch.float64):
            src = batch.pinned(name) if hasattr(batch, 'pinned') else None
            if src is None:
                a = np.ascontiguousarray(getattr(batch, name))
                stage = self._host_pool.get('up.' + name)
                if stage is None or stage.shape != a.shape or stage.dtype != torch.from_numpy(a).dtype:
                    stage = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
                    self._host_pool['up.' + name] = stage
                stage.numpy()[...] = a
                src = stage
            chunks = batch.host_chunks(name) if hasattr(batch, 'host_chunks') else []
            if chunks:
                # the download of this field is still in flight: wait for it on the device, chunk by chunk, and send every chunk
                # back up as soon as it has landed (the two directions of the link run at the same time)
                out = torch.empty(src.shape, dtype=src.dtype, device=dev)
                cur = torch.cuda.current_stream()
                for r0, r1, ev in chunks:
                    cur.wait_event(ev)
                    out[r0:r1].copy_(src[r0:r1], non_blocking=True)
                return out.to(dtype)
            return src.to(dev, non_blocking=True).to(dtype)
        has_vm = 'v_metas' in getattr(batch, '_host', {}) or (not hasattr(batch, '_host') and hasattr(batch, 'v_metas'))
        vm = up('v_metas', torch.int32) if has_vm else None
        return (up('states'), up('actions'), up('rewards'), up('masks'), up('exps'), vm, getattr(batch, 'horizon', None))

    def _inputs(self, states, v_metas, masks, horizon):
        """trans_policy / trans_value of the base agent: identity"""
        inp = _NetInput(x_const=states)
        return inp, inp

    # ---- int8-tensor-core dense layers --------------------------------------------------------------
    def _oz(self, trunk, inp):
        """OzMlp of a trunk when the ozaki backend is selected, else None"""
        if self.gemm != 'ozaki':
            return None
        dims = tuple(int(v) for v in trunk.dims())
        # rows per chunk: 8 waves of one 128-row tile per SM (per-kernel fixed costs amortised), less for small batches
        base = lib.load().egp_oz_mlp_chunk_rows()
        n = (inp.xbuf if inp.learned else inp.x_const).shape[0]
        chunk = int(min(self.oz_chunk_waves, -(-n // base)) * base)
        oz = self._ozs.get(dims + (chunk,))
        if oz is None:
            oz = lib.OzMlp(*dims, n_slices=self.oz_slices, chunk_rows=chunk, device=trunk.W(0).device)
            self._ozs[dims + (chunk,)] = oz
        return oz

    def _xcache(self, oz, x):
        """input-slice cache of one constant input tensor for the current update (shared by nets with the same input)"""
        key = (x.data_ptr(), x.shape[0], x.shape[1], oz.chunk, oz.S)
        ent = self._xcaches.get(key)
        if ent is None:
            # reuse an allocation of the same size from an earlier update
            for k in list(self._xcaches):
                if k[1:] == key[1:] and not self._xcaches[k].get('live'):
                    ent = self._xcaches.pop(k)
                    break
            if ent is None:
                ent = oz.new_cache(x.shape[0])
            ent['valid'] = False
            self._xcaches[key] = ent
        ent['live'] = True
        return ent

    def update_value(self, x, returns, inv_n, reuse_forward=False, cache=True, defer=False):
        """agents/agent_pg.py:19-26.  ``reuse_forward``: the activations of the value forward that produced the
        GAE inputs are still valid (no parameter step since), so the first epoch skips its forward GEMMs.
        ``x`` is a tensor or a _NetInput."""
        inp = x if isinstance(x, _NetInput) else _NetInput(x_const=x)
        oz = self._oz(self._vt, inp)
        for it in range(self.value_opt_niter):
            if oz is not None:
                xt = inp.x() if inp.learned else inp.x_const
                dx = self._vt._buf('dx', (xt.shape[0], inp.cd), xt) if inp.learned else None
                self._scal[0:1].zero_()
                oz.step(self._vt.weights(), xt, grads=self._vt.grads(), dx=dx,
                        cache=self._xcache(oz, xt) if (cache and not inp.learned) else None,
                        loss=dict(kind='value', returns=returns, inv_n=inv_n, loss=self._scal[0:1]))
                if inp.learned:
                    inp.backward(dx)
                if defer:               # the caller all-reduces the flat [value | policy] gradient once and steps both nets
                    continue
                d = _dist()
                if d is not None:
                    d.all_reduce(self._vf.grad)
                self._vf.adam(0.0)
                continue
            if reuse_forward and it == 0 and not inp.learned:
                v, xt = self._vt.buf['y'], self._vt.x
            else:
                xt = inp.x()
                v = self._vt.forward(xt)
            dv = self._vt._buf('dy', v.shape, xt)
            self._scal[0:1].zero_()
            lib.value_loss_grad(v.view(-1), returns, inv_n, dv.view(-1), self._scal[0:1])
            dx = self._vt.backward(dv, inp.cd if inp.learned else 0)
            if inp.learned:
                inp.backward(dx)
            if defer:
                continue
            d = _dist()
            if d is not None:
                d.all_reduce(self._vf.grad)
            self._vf.adam(0.0)

    def _reduce_and_step(self, max_norm):
        """one sum all-reduce of the flat [value | policy] gradient, then both optimizer steps (value first, agent_ppo.py:46-51;
        the nets share no parameters, so stepping the value net after the policy backward changes nothing)"""
        d = _dist()
        if self._peer is not None:
            self._gflat.copy_(self._peer.allreduce())     # the kernel has completed before the copy: no peer reads src any more
        elif d is not None:
            d.all_reduce(self._gflat)
        self._vf.adam(0.0)
        self._pf.adam(max_norm)

    def update_params(self, batch):
        t0 = time.time()
        self._setup()
        if self._peer is not None:
            # a rank that never arrived at a gradient exchange of the previous update (barrier timeout inside the kernel)
            # leaves garbage in the sum: fail loudly instead of training on it
            err = self._peer.error()
            if err:
                raise lib.EgpError('peer-memory gradient exchange: barrier %d timed out in the previous update' % err)
        states, actions, rewards, masks, exps, v_metas, horizon = self._device_batch(batch)
        xp, xv = self._inputs(states, v_metas, masks, horizon)
        # values + GAE (agent_pg.py:48-53, core/common.py:5-25)
        # input slices are valid for ONE update only: the rollout buffers (and the caching allocator's blocks) keep
        # their addresses across iterations, so a pointer match says nothing about the contents
        for ent in self._xcaches.values():
            ent['live'] = False
            ent['valid'] = False
        oz = self._oz(self._vt, xv)
        if oz is not None:
            xt = xv.x(grad=False) if xv.learned else xv.x_const
            values = oz.step(self._vt.weights(), xt, y=self._vt._buf('y', (xt.shape[0], 1), xt),
                             cache=None if xv.learned else self._xcache(oz, xt)).view(-1)
            self._value_fresh = False
        else:
            values = self._vt.forward(xv.x(grad=False)).view(-1)
            self._value_fresh = not xv.learned
        adv, returns, stats = lib.gae(rewards, masks, values.contiguous(), self.gamma, self.tau)
        n_local = states.shape[0]
        n_exp = exps.sum()
        n_global = float(n_local)
        if _dist() is not None:
            # global standardisation / denominators: Chan merge of (n, mean, M2) over ranks
            dist_utils.merge_moments_(stats)
            dist_utils.allreduce_sum_(n_exp)
            n_global = float(stats[0].item())
        self._stats = stats
        n_exp = float(n_exp.item())
        # a batch without sampled rows (mean actions only) has no surrogate term: zero policy gradient instead of
        # the reference's NaN mean over an empty selection (agent_ppo.py:45,64)
        self.update_policy(xp, xv, actions, returns, adv, exps, 1.0 / n_exp if n_exp > 0 else 0.0, 1.0 / n_global)
        torch.cuda.synchronize()
        return time.time() - t0

    def update_policy(self, xp, xv, actions, returns, adv, exps, inv_count, inv_n):
        raise NotImplementedError('AgentPG (A2C) policy step is not on the hot path; use AgentPPO / AgentEgo')


class AgentPPO(AgentPG):
    def __init__(self, clip_epsilon=0.2, opt_batch_size=64, use_mini_batch=False, policy_grad_clip=None, **kwargs):
        super().__init__(**kwargs)
        self.clip_epsilon = clip_epsilon
        self.opt_batch_size = opt_batch_size
        self.use_mini_batch = use_mini_batch
        self.policy_grad_clip = policy_grad_clip

    def _max_norm(self):
        if not self.policy_grad_clip:
            return 0.0
        if len(self.policy_grad_clip) != 1:
            raise lib.EgpError('one (params, max_norm) clip group is supported (ego_mimic.py:90)')
        return float(self.policy_grad_clip[0][1])

    def _policy_step(self, xp, actions, adv, logp0, exps, inv_count, log_std, max_norm, mu=None, cache=True, init_logp0=False,
                     defer=False):
        """ppo_loss forward+backward, gradient all-reduce, clip + Adam (agent_ppo.py:47-51 / :38-43)"""
        inp = xp if isinstance(xp, _NetInput) else _NetInput(x_const=xp)
        oz = self._oz(self._pt, inp)
        if oz is not None:
            xt = inp.x() if inp.learned else inp.x_const
            dx = self._pt._buf('dx', (xt.shape[0], inp.cd), xt) if inp.learned else None
            self._scal[1:2].zero_()
            dls = None
            if self._learn_std:
                dls = self._pf.view(self._pf.grad, 'action_log_std').view(-1)
                dls.zero_()
            oz.step(self._pt.weights(), xt, grads=self._pt.grads(), dx=dx,
                    cache=self._xcache(oz, xt) if (cache and not inp.learned) else None,
                    loss=dict(kind='ppo', actions=actions, log_std=log_std, adv=adv, stats=self._stats, logp0=logp0, exps=exps,
                              clip_eps=self.clip_epsilon, inv_count=inv_count, dlogstd=dls, loss=self._scal[1:2],
                              init_logp0=init_logp0))
            if inp.learned:
                inp.backward(dx)
            if defer:
                return self._scal[1:2].clone()
            if _dist() is not None:
                _dist().all_reduce(self._pf.grad)
            self._pf.adam(max_norm)
            return self._scal[1:2].clone()
        if mu is None or inp.learned:
            mu = self._pt.forward(inp.x())
        dmu = self._pt._buf('dy', mu.shape, mu)
        self._scal[1:2].zero_()
        dls = None
        if self._learn_std:
            dls = self._pf.view(self._pf.grad, 'action_log_std').view(-1)
            dls.zero_()
        lib.ppo_loss_grad(mu, actions, log_std, adv, self._stats, logp0, exps, self.clip_epsilon, inv_count, dmu,
                          dls, self._scal[1:2])
        dx = self._pt.backward(dmu, inp.cd if inp.learned else 0)
        if inp.learned:
            inp.backward(dx)
        if defer:
            return self._scal[1:2].clone()
        if _dist() is not None:
            _dist().all_reduce(self._pf.grad)
        self._pf.adam(max_norm)
        return self._scal[1:2].clone()

    def _update_policy_minibatch(self, xp, xv, actions, returns, adv, exps, logp0, log_std, max_norm):
        """agents/agent_ppo.py:24-43: per epoch a fresh np.random.shuffle permutation applied on top of the
        previous one, contiguous slices of opt_batch_size rows, value then policy step per slice with per-slice
        means.  The permutation is composed on the host and applied with one row-gather per array."""
        n = xp.shape[0]
        B = int(self.opt_batch_size)
        nb = int(math.ceil(n / B))
        dev = xp.device
        same_x = xv is xp
        g = {k: torch.empty_like(t) for k, t in (('xp', xp), ('ac', actions), ('ret', returns), ('adv', adv), ('lp', logp0),
                                                  ('ex', exps))}
        if not same_x:
            g['xv'] = torch.empty_like(xv)
        cur = np.arange(n)
        surr, vloss = [], []
        for _ in range(self.opt_num_epochs):
            perm = np.arange(n)
            np.random.shuffle(perm)                                      # :26-27 (global numpy RNG, like the reference)
            cur = cur[perm]
            pd = torch.from_numpy(cur).to(dev, non_blocking=True)
            for k, src in (('xp', xp), ('ac', actions), ('ret', returns), ('adv', adv), ('lp', logp0), ('ex', exps)):
                lib.gather_rows(src, pd, g[k])
            if not same_x:
                lib.gather_rows(xv, pd, g['xv'])
            gxv = g['xp'] if same_x else g['xv']
            # exps count of every slice with one device->host copy per epoch
            pad = nb * B - n
            ex = torch.cat([g['ex'], g['ex'].new_zeros(pad)]) if pad else g['ex']
            counts = ex.view(nb, B).sum(1)
            dist_utils.allreduce_sum_(counts)
            counts = counts.cpu().numpy()
            ws = dist_utils.world_size()
            for i in range(nb):
                lo, hi = i * B, min((i + 1) * B, n)
                defer = _dist() is not None and self.value_opt_niter == 1
                self.update_value(gxv[lo:hi], g['ret'][lo:hi], 1.0 / ((hi - lo) * ws), cache=False, defer=defer)
                vloss.append(self._scal[0:1].clone())
                surr.append(self._policy_step(g['xp'][lo:hi], g['ac'][lo:hi], g['adv'][lo:hi], g['lp'][lo:hi], g['ex'][lo:hi],
                                              (1.0 / float(counts[i])) if counts[i] > 0 else 0.0, log_std, max_norm, cache=False,
                                              defer=defer))
                if defer:
                    self._reduce_and_step(max_norm)
        self.last_info = dict(surr_loss=surr, value_loss=vloss)

    def update_policy(self, xp, xv, actions, returns, adv, exps, inv_count, inv_n):
        """agents/agent_ppo.py:16-51"""
        log_std = self.policy_net.action_log_std.data.view(-1)
        oz = self._oz(self._pt, xp)
        # Full-batch branch on the tensor-core back end: the first epoch's policy pass runs at the parameters that define
        # fixed_log_probs (:18-20), so its loss kernel WRITES them (ratio = exp(0) = 1 there, in the reference as well) and
        # the separate no-grad forward over the batch is skipped
        fuse0 = oz is not None and not self.use_mini_batch and not xp.learned and os.environ.get('EGP_FUSE_LOGP0', '1') == '1'
        mu = None
        if fuse0:
            logp0 = self._pt._buf('logp0', (actions.shape[0],), actions)
        else:
            if oz is not None:
                xt = xp.x(grad=False) if xp.learned else xp.x_const
                mu = oz.step(self._pt.weights(), xt, y=self._pt._buf('y', (xt.shape[0], self._pt.dims()[3]), xt),
                             cache=None if xp.learned else self._xcache(oz, xt))
            else:
                mu = self._pt.forward(xp.x(grad=False))
            logp0 = lib.gauss_logp(mu, actions, log_std)                # fixed_log_probs (:18-20)
        max_norm = self._max_norm()
        if self.use_mini_batch:
            if xp.learned or xv.learned:
                raise lib.EgpError('mini-batch PPO with a VideoStateNet context is not supported (the reference forces '
                                   'the full-batch branch for AgentEgo, agent_ego.py:11)')
            return self._update_policy_minibatch(xp.x_const, xv.x_const, actions, returns, adv, exps, logp0, log_std, max_norm)
        surr, vloss = [], []
        d = _dist()
        # The value chain (10 x [forward / loss / backward / Adam] of the value net) and the policy chain never exchange
        # data inside update_policy (advantages / returns are fixed before the epochs, agent_ppo.py:44-51), so they are
        # issued on two streams (EGP_UPDATE_STREAMS=1 restores one): slicing (HBM-bound) and GEMM (tensor-bound) kernels of
        # the two chains overlap at tile-wave tails.  Same arithmetic, same order within each chain.
        two = (os.environ.get('EGP_UPDATE_STREAMS', '2') == '2' and oz is not None and not xp.learned and not xv.learned)
        if two:
            cur = torch.cuda.current_stream()
            if getattr(self, '_vstream', None) is None:
                self._vstream = torch.cuda.Stream()
            self._vstream.wait_stream(cur)
        # several ranks: both backward passes of an epoch first, then ONE all-reduce of the flat [value | policy] gradient
        # (SURVEY 8e collective C1) and both optimizer steps; the two streams join at that point
        defer = d is not None and self.value_opt_niter == 1
        for ep in range(self.opt_num_epochs):
            if two:
                with torch.cuda.stream(self._vstream):
                    self.update_value(xv, returns, inv_n, defer=defer)
                    vloss.append(self._scal[0:1].clone())
            else:
                self.update_value(xv, returns, inv_n, reuse_forward=(ep == 0 and self._value_fresh), defer=defer)     # :46
                vloss.append(self._scal[0:1].clone())
            surr.append(self._policy_step(xp, actions, adv, logp0, exps, inv_count, log_std, max_norm,
                                          mu=mu if ep == 0 else None,       # epoch 0 reuses the fixed-log-prob forward
                                          init_logp0=fuse0 and ep == 0, defer=defer))
            if defer:
                if two:
                    torch.cuda.current_stream().wait_stream(self._vstream)
                self._reduce_and_step(max_norm)
                if two:
                    self._vstream.wait_stream(torch.cuda.current_stream())
        if two:
            torch.cuda.current_stream().wait_stream(self._vstream)
        self.last_info = dict(surr_loss=surr, value_loss=vloss)

    def losses(self):
        """per-epoch (surrogate, value) losses of the last update as python floats"""
        li = self.last_info
        d = _dist()
        out = {}
        for k in ('surr_loss', 'value_loss'):
            t = torch.cat(li[k]) if li.get(k) else torch.zeros(0)
            if d is not None and t.numel():
                d.all_reduce(t)
            out[k] = t.cpu().numpy()
        return out


class AgentEgo(AgentPPO):
    def __init__(self, policy_vs_net=None, value_vs_net=None, use_mini_batch=False, **kwargs):
        # the reference forces the full-batch branch (agent_ego.py:11) because its BiLSTM context packing needs
        # whole episodes; the per-frame context table has no such constraint, so BASELINE config 5's
        # mini-batch PPO can be switched on explicitly (use_mini_batch=True, opt_batch_size=...)
        super().__init__(use_mini_batch=use_mini_batch, **kwargs)
        self.traj_cls = TrajBatchEgo
        self.policy_vs_net = policy_vs_net
        self.value_vs_net = value_vs_net
        self.sample_modules.append(policy_vs_net)
        self.update_modules += [policy_vs_net, value_vs_net]
        for net in (policy_vs_net, value_vs_net):
            if net is not None and not isinstance(net, (FrameContext, VideoStateNet, VideoForecastNet)):
                raise lib.EgpError('video context nets must be nets.VideoStateNet / VideoForecastNet (lstm) or nets.FrameContext')

    def pre_sample(self):
        if self.policy_vs_net is not None:
            self.policy_vs_net.set_mode('test')

    def _rollout_context(self):
        """test-mode VideoStateNet output for every (take, start) episode window, evaluated in one batched sweep
        (replaces pre_episode's per-episode initialize, agent_ego.py:21-22): the kernel looks v_out[t] up on the
        device across auto-resets"""
        if isinstance(self.policy_vs_net, (VideoStateNet, VideoForecastNet)):
            return self.policy_vs_net.context_table(self.env.cnn_feat, self.env.cfg.env_episode_len)
        return None, None

    def _rollout_extra(self):
        """VideoForecastNet (ego_forecast.py:53-56): v_out is one constant row per episode window and the state LSTM
        is stepped inside the rollout kernel (video_forecast_net.py:57-61,86-93)"""
        vs = self.policy_vs_net
        if isinstance(vs, VideoForecastNet):
            return dict(ctx_const=True, snet=vs.snet_packed())
        return {}

    def _inputs(self, states, v_metas, masks, horizon):
        """trans_policy / trans_value (agent_ego.py:28-32, 44-47): VideoStateNet in train mode, or the per-frame
        table gathered on the device"""
        out = []
        cache = None
        for vs, flat in ((self.policy_vs_net, self._pf), (self.value_vs_net, self._vf)):
            if isinstance(vs, (VideoStateNet, VideoForecastNet)):
                vs.set_mode('train')
                vs.initialize((masks, self.env.cnn_feat, v_metas))
                out.append(_NetInput(vs=vs, states=states, flat=flat))
                continue
            if cache is None:
                if self.env is None or self.env.kernel.ctx_dim == 0:
                    cache = _NetInput(x_const=states)
                else:
                    if horizon is None:
                        raise lib.EgpError('context gather needs the env-major [E, T] batch produced by sample()')
                    cache = _NetInput(x_const=self.env.kernel.build_input(states, v_metas, masks, horizon))
            out.append(cache)
        return out[0], out[1]
