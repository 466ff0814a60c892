"""Host mirrors of the reference's network classes on the hot path, same constructor arguments,
attribute names and state-dict keys so checkpoints and callers carry over:

  MLP             models/mlp.py:5-25          (net.affine_layers.{i}.{weight,bias}, .out_dim)
  DiagGaussian    core/distributions.py:6-25
  Policy          core/policy.py:4-23
  PolicyGaussian  core/policy_gaussian.py:8-38 (action_mean.*, action_log_std [1, A], .type == 'gaussian')
  Value           core/critic.py:5-18         (value_head.*)

These modules only HOLD parameters and give the convenience forward (torch) the reference API promises;
the rollout and the PPO update read the parameter storage directly from the CUDA kernels
(egopose_b200/agent.py) and never call forward().
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn
from torch.distributions import Normal


class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dims=(128, 128), activation='tanh'):
        super().__init__()
        self.activation_name = activation
        self.activation = {'tanh': torch.tanh, 'relu': torch.relu, 'sigmoid': torch.sigmoid}[activation]
        self.out_dim = hidden_dims[-1]
        self.affine_layers = nn.ModuleList()
        last = input_dim
        for nh in hidden_dims:
            self.affine_layers.append(nn.Linear(last, nh))
            last = nh

    def forward(self, x):
        for affine in self.affine_layers:
            x = self.activation(affine(x))
        return x


class DiagGaussian(Normal):
    def __init__(self, loc, scale):
        super().__init__(loc, scale)

    def kl(self):
        loc1, scale1, log_scale1 = self.loc, self.scale, self.scale.log()
        loc0, scale0, log_scale0 = loc1.detach(), scale1.detach(), log_scale1.detach()
        kl = log_scale1 - log_scale0 + (scale0.pow(2) + (loc0 - loc1).pow(2)) / (2.0 * scale1.pow(2)) - 0.5
        return kl.sum(1, keepdim=True)

    def log_prob(self, value):
        return super().log_prob(value).sum(1, keepdim=True)

    def mean_sample(self):
        return self.loc


class Policy(nn.Module):
    def forward(self, x):
        raise NotImplementedError

    def select_action(self, x, mean_action=False):
        dist = self.forward(x)
        return dist.mean_sample() if mean_action else dist.sample()

    def get_kl(self, x):
        return self.forward(x).kl()

    def get_log_prob(self, x, action):
        return self.forward(x).log_prob(action)


class PolicyGaussian(Policy):
    def __init__(self, net, action_dim, net_out_dim=None, log_std=0, fix_std=False):
        super().__init__()
        self.type = 'gaussian'
        self.net = net
        if net_out_dim is None:
            net_out_dim = net.out_dim
        self.action_mean = nn.Linear(net_out_dim, action_dim)
        self.action_mean.weight.data.mul_(0.1)
        self.action_mean.bias.data.mul_(0.0)
        self.action_log_std = nn.Parameter(torch.ones(1, action_dim) * log_std, requires_grad=not fix_std)

    def forward(self, x):
        mean = self.action_mean(self.net(x))
        return DiagGaussian(mean, torch.exp(self.action_log_std.expand_as(mean)))

    def get_fim(self, x):
        dist = self.forward(x)
        cov_inv = self.action_log_std.exp().pow(-2).squeeze(0).repeat(x.size(0))
        param_count, std_index, std_id = 0, 0, 0
        for i, (name, param) in enumerate(self.named_parameters()):
            if name == 'action_log_std':
                std_id, std_index = i, param_count
            param_count += param.view(-1).shape[0]
        return cov_inv.detach(), dist.loc, {'std_id': std_id, 'std_index': std_index}


class Value(nn.Module):
    def __init__(self, net, net_out_dim=None):
        super().__init__()
        self.net = net
        if net_out_dim is None:
            net_out_dim = net.out_dim
        self.value_head = nn.Linear(net_out_dim, 1)
        self.value_head.weight.data.mul_(0.1)
        self.value_head.bias.data.mul_(0.0)

    def forward(self, x):
        return self.value_head(self.net(x))


class FrameContext(nn.Module):
    """Stand-in for models/video_state_net.py on the fused path: the per-step video context v_out[t] that
    the reference concatenates in front of the state (video_state_net.py:62-64) is served from a per-frame
    table uploaded with the experts (one row per expert frame).  With the raw CNN features as the table this
    is VideoStateNet with an identity v_net; the BiLSTM producer is SURVEY 8f row 2 (next)."""

    def __init__(self, cnn_feat_dim, v_hdim=None, v_margin=10):
        super().__init__()
        self.cnn_feat_dim = cnn_feat_dim
        self.v_hdim = v_hdim or cnn_feat_dim
        self.v_margin = v_margin
        self.mode = 'test'

    def set_mode(self, mode):
        self.mode = mode

    def initialize(self, x):
        pass

    def forward(self, x):
        raise RuntimeError('FrameContext is consumed inside the fused kernels (egp_rollout / egp_build_input)')


class _LstmRecurrence(torch.autograd.Function):
    """Sequential half of an LSTM sweep on the fused kernels (csrc/lstm.cu, egp_lstm_seq_fwd / bwd_f64): packed gate
    pre-activations xi [Np, 4H] (= x W_ih^T + biases, a plain GEMM of the caller) -> hidden states [Np, H].  Rows are time-major
    packed, step s owns rows [off[s], off[s+1]); ``prev`` maps a row to the row of the same batch element one step earlier
    (-1 at its first step).  backward returns d xi and dW_hh = d xi^T H_prev (GEMM)."""

    @staticmethod
    def forward(ctx, xi, weight_hh, off, prev, L, B):
        from . import lib
        H = weight_hh.shape[1]
        wf, wb = lib.lstm_pack_whh(weight_hh.detach())
        h, gates, c = lib.lstm_seq_fwd(xi.detach().contiguous(), off, L, B, H, wf)
        ctx.save_for_backward(h, gates, c, off, prev, wb)
        ctx.dims = (L, B, H)
        return h

    @staticmethod
    def backward(ctx, dh):
        from . import lib
        h, gates, c, off, prev, wb = ctx.saved_tensors
        L, B, H = ctx.dims
        dxi = lib.lstm_seq_bwd(dh.contiguous(), gates, c, off, L, B, H, wb)
        # dW_hh = sum over rows of dxi_row^T h_prev_row: rows at their first step have h_prev = 0
        hp = h.index_select(0, prev.clamp(min=0)) * (prev >= 0).to(h.dtype).unsqueeze(1)
        return dxi, dxi.t() @ hp, None, None, None, None


def _fused_lstm_ok(cell, x):
    from . import lib
    return x.is_cuda and x.dtype == torch.float64 and cell.hidden_size in lib.LSTM_FUSED_H and os.environ.get('EGP_LSTM', 'fused') == 'fused'


class RNN(nn.Module):
    """models/rnn.py:5-61 mirror (LSTM cells only): parameters live in nn.LSTMCell modules named rnn_f / rnn_b so
    the state-dict keys match the reference; batch mode runs the whole sequence.  The input projection of all
    time steps is one GEMM, only the recurrent half is stepped."""

    def __init__(self, input_dim, out_dim, cell_type='lstm', bi_dir=False):
        super().__init__()
        if cell_type != 'lstm':
            raise NotImplementedError('only the lstm cell is on the hot path (egomimic_config.py:54,63)')
        self.input_dim, self.out_dim, self.cell_type, self.bi_dir = input_dim, out_dim, cell_type, bi_dir
        self.mode = 'batch'
        hidden = out_dim // 2 if bi_dir else out_dim
        self.rnn_f = nn.LSTMCell(input_dim, hidden)
        if bi_dir:
            self.rnn_b = nn.LSTMCell(input_dim, hidden)
        self.hx = self.cx = None

    def set_mode(self, mode):
        self.mode = mode

    def initialize(self, batch_size=1):
        if self.mode == 'step':
            p = self.rnn_f.weight_ih
            self.hx = torch.zeros((batch_size, self.rnn_f.hidden_size), dtype=p.dtype, device=p.device)
            self.cx = torch.zeros_like(self.hx)

    @staticmethod
    def _sweep(cell, x, reverse):
        """x [L, B, F] -> hidden states [L, B, H] of one direction, zero initial state"""
        L, B, _ = x.shape
        H = cell.hidden_size
        xi = torch.addmm(cell.bias_ih + cell.bias_hh, x.reshape(L * B, -1), cell.weight_ih.t()).view(L, B, 4 * H)
        if _fused_lstm_ok(cell, x):
            # fused recurrence (one launch per sweep): iteration order = time order, reversed for the backward direction
            xs = xi.flip(0) if reverse else xi
            off = torch.arange(L + 1, device=x.device, dtype=torch.int64) * B
            rows = torch.arange(L * B, device=x.device, dtype=torch.int64)
            prev = torch.where(rows >= B, rows - B, torch.full_like(rows, -1))
            hs = _LstmRecurrence.apply(xs.reshape(L * B, 4 * H), cell.weight_hh, off, prev, L, B).view(L, B, H)
            return hs.flip(0) if reverse else hs
        h = x.new_zeros((B, H))
        c = x.new_zeros((B, H))
        whh_t = cell.weight_hh.t()
        outs = [None] * L
        for t in (range(L - 1, -1, -1) if reverse else range(L)):
            gates = xi[t] + h @ whh_t
            i, f, g, o = gates.chunk(4, 1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs[t] = h
        return torch.stack(outs, 0)

    def forward(self, x):
        if self.mode == 'step':
            self.hx, self.cx = self.rnn_f(x, (self.hx.to(x.device), self.cx.to(x.device)))
            return self.hx
        out = self._sweep(self.rnn_f, x, False)
        if self.bi_dir:
            out = torch.cat((out, self._sweep(self.rnn_b, x, True)), 2)
        return out


class VideoStateNet(nn.Module):
    """models/video_state_net.py:8-79 mirror (lstm v_net): same constructor, modes and forward semantics.

    test mode  : initialize(cnn_feat[start-m : start+T+m]) runs the (Bi)LSTM once per episode, forward(state)
                 prepends v_out[t]                                                            (:36-39,61-64)
    train mode : initialize((masks, cnn_feat, v_metas)) packs every episode of the batch into a padded
                 [Tmax+2m, n_ep, F] context, forward re-runs the (Bi)LSTM and gathers rows       (:40-59,65-69)

    Fused-path entry points (used by egopose_b200.agent.AgentEgo): ``context_table`` evaluates test mode for EVERY
    (take, start) window in one batched sweep so the rollout kernel can look the context up on the device across
    auto-resets; ``train_context`` is train-mode forward without the concatenation (autograd graph attached)."""

    def __init__(self, cnn_feat_dim, v_hdim=128, v_margin=10, v_net_type='lstm', v_net_param=None, causal=False):
        super().__init__()
        if v_net_type != 'lstm':
            raise NotImplementedError('v_net_type %r: only lstm is selected by the shipped configs' % v_net_type)
        self.mode = 'test'
        self.cnn_feat_dim, self.v_net_type, self.v_hdim, self.v_margin = cnn_feat_dim, v_net_type, v_hdim, v_margin
        self.v_net = RNN(cnn_feat_dim, v_hdim, v_net_type, bi_dir=not causal)
        self.v_out, self.t = None, 0
        self.indices = self.cnn_feat_ctx = None

    def set_mode(self, mode):
        self.mode = mode

    def _p(self):
        return self.v_net.rnn_f.weight_ih

    def forward_v_net(self, x):
        return self.v_net(x)

    def initialize(self, x):
        if self.mode == 'test':
            m = self.v_margin
            self.v_out = self.forward_v_net(x.unsqueeze(1)).squeeze(1)[m:-m]
            self.t = 0
            return
        masks, cnn_feat, v_metas = x
        p = self._p()
        m = self.v_margin
        masks_np = masks.detach().cpu().numpy() if torch.is_tensor(masks) else np.asarray(masks)
        v_metas = v_metas.detach().cpu().numpy() if torch.is_tensor(v_metas) else np.asarray(v_metas)
        ends = np.nonzero(masks_np == 0)[0]
        starts = np.concatenate([[0], ends[:-1] + 1])
        lens = ends - starts + 1
        tmax = int(lens.max())
        n = masks_np.shape[0]
        ep_of = np.repeat(np.arange(len(ends)), lens)
        self.indices = torch.as_tensor(ep_of * tmax + (np.arange(n) - starts[ep_of]), dtype=torch.long, device=p.device)
        # frame gather instead of the reference's per-episode python loop
        feats = cnn_feat if torch.is_tensor(cnn_feat) else None
        if feats is None:
            offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in cnn_feat])])
            feats = torch.as_tensor(np.concatenate(cnn_feat), dtype=p.dtype, device=p.device)
        else:
            raise ValueError('cnn_feat must be the per-take list (env.cnn_feat)')
        meta = v_metas[ends].astype(np.int64)
        base = offs[meta[:, 0]] + meta[:, 1] - m
        frame = base[None, :] + np.arange(tmax + 2 * m)[:, None]
        if frame.min() < 0 or frame.max() >= feats.shape[0]:
            raise IndexError('episode context window leaves the take (video_state_net.py:55)')
        self.cnn_feat_ctx = feats[torch.as_tensor(frame, device=p.device)]          # [tmax + 2m, n_ep, F]

    def train_context(self):
        m = self.v_margin
        v_ctx = self.forward_v_net(self.cnn_feat_ctx)[m:-m]
        v_ctx = v_ctx.transpose(0, 1).reshape(-1, self.v_hdim)
        return v_ctx.index_select(0, self.indices)

    def forward(self, x):
        if self.mode == 'test':
            x = torch.cat((self.v_out[[self.t], :], x), dim=1)
            self.t += 1
            return x
        return torch.cat((self.train_context(), x), dim=1)

    @torch.no_grad()
    def context_table(self, cnn_feat, episode_len, max_batch=4096):
        """-> (table [n_windows * episode_len, v_hdim], win_off int32 [n_takes + 1]); window w of take k starts at
        frame start = v_margin + (w - win_off[k]) (the range reset_model samples from, humanoid_v1.py:214) and row
        w * episode_len + t equals test-mode v_out[t] of an episode started there."""
        p, m, T = self._p(), self.v_margin, int(episode_len)
        rows, win_off = [], [0]
        for feat in cnn_feat:
            f = torch.as_tensor(feat, dtype=p.dtype, device=p.device)
            nwin = f.shape[0] - T - 2 * m
            if nwin <= 0:
                raise ValueError('take shorter than episode_len + 2 * fr_margin')
            for lo in range(0, nwin, max_batch):
                hi = min(nwin, lo + max_batch)
                idx = torch.arange(T + 2 * m, device=p.device)[:, None] + torch.arange(lo, hi, device=p.device)[None, :]
                out = self.forward_v_net(f[idx])[m:-m]                      # [T, hi - lo, H]
                rows.append(out.transpose(0, 1).reshape(-1, self.v_hdim))
            win_off.append(win_off[-1] + nwin)
        return torch.cat(rows).contiguous(), torch.tensor(win_off, dtype=torch.int32, device=p.device)


class VideoForecastNet(nn.Module):
    """models/video_forecast_net.py:8-111 mirror (lstm v_net, 'lstm' | 'id' s_net, dynamic_v False - what the shipped
    egoforecast configs select): same constructor, modes, forward semantics and state-dict keys (v_net.rnn_f.*,
    s_net.rnn_f.*).

    test mode  : initialize(cnn_feat[start-m : ...]) runs the causal LSTM over the m = v_margin frames BEFORE the episode
                 start, v_out = last hidden state (constant over the episode); forward(state) steps the state LSTM
                 once and returns cat(v_out, h_t)                                                   (:57-61, 86-93)
    train mode : initialize((masks, cnn_feat, v_metas)) splits the flat batch into episodes; forward(states) re-runs
                 the causal LSTM per episode and unrolls the state LSTM over every episode from a zero state (:62-107)

    Fused-path entry points (egopose_b200.agent.AgentEgo): ``context_table`` = test-mode v_out for EVERY (take, start)
    window (one row per window, rollout ctx_mode 2); ``snet_packed`` = the state-LSTM weights in the row order the
    rollout kernel steps them in; ``train_context(states)`` = train-mode forward with the autograd graph attached.
    The reference pads every episode to the longest one (:97-101); here episodes are sorted by length and each unroll
    step only touches the episodes still alive, so the work is proportional to the batch, not to n_ep x Tmax."""

    def __init__(self, cnn_feat_dim, state_dim, v_hdim=128, v_margin=10, v_net_type='lstm', v_net_param=None,
                 s_hdim=None, s_net_type='id', dynamic_v=False):
        super().__init__()
        if v_net_type != 'lstm':
            raise NotImplementedError('v_net_type %r: only lstm is selected by the shipped configs' % v_net_type)
        if s_net_type not in ('lstm', 'id'):
            raise NotImplementedError('s_net_type %r' % s_net_type)
        if dynamic_v:
            raise NotImplementedError('dynamic_v is not selected by any shipped config (egoforecast_config.py:52,65)')
        s_hdim = state_dim if s_hdim is None else s_hdim
        self.mode = 'test'
        self.cnn_feat_dim, self.state_dim = cnn_feat_dim, state_dim
        self.v_net_type, self.v_hdim, self.v_margin = v_net_type, v_hdim, v_margin
        self.s_net_type, self.s_hdim, self.dynamic_v = s_net_type, s_hdim, dynamic_v
        self.out_dim = v_hdim + s_hdim
        self.v_net = RNN(cnn_feat_dim, v_hdim, v_net_type, bi_dir=False)
        if s_net_type == 'lstm':
            self.s_net = RNN(state_dim, s_hdim, s_net_type, bi_dir=False)
        self.v_out, self.t = None, 0
        self._ep = None
        self.set_mode('test')

    def set_mode(self, mode):
        self.mode = mode
        if self.s_net_type == 'lstm':
            self.s_net.set_mode('batch' if mode == 'train' else 'step')

    def _p(self):
        return self.v_net.rnn_f.weight_ih

    def forward_v_net(self, x):
        return self.v_net(x)

    def initialize(self, x):
        if self.mode == 'test':
            self.v_out = self.forward_v_net(x.unsqueeze(1)[:self.v_margin])[-1]
            if self.s_net_type == 'lstm':
                self.s_net.initialize()
            self.t = 0
            return
        masks, cnn_feat, v_metas = x
        p, m = self._p(), self.v_margin
        masks_np = masks.detach().cpu().numpy() if torch.is_tensor(masks) else np.asarray(masks)
        v_metas = v_metas.detach().cpu().numpy() if torch.is_tensor(v_metas) else np.asarray(v_metas)
        ends = np.nonzero(masks_np == 0)[0]
        starts = np.concatenate([[0], ends[:-1] + 1])
        lens = ends - starts + 1
        order = np.argsort(-lens, kind='stable')                    # longest episode first
        lens_s, starts_s = lens[order], starts[order]
        tmax = int(lens_s[0])
        # alive[t] = number of (sorted) episodes with length > t; rows[t] = flat batch rows of their step t
        alive = np.searchsorted(-lens_s, -np.arange(tmax), side='left')
        rows = [torch.as_tensor(starts_s[:alive[t]] + t, dtype=torch.long, device=p.device) for t in range(tmax)]
        n = masks_np.shape[0]
        rank = np.empty(len(ends), dtype=np.int64)
        rank[order] = np.arange(len(ends))
        ep_of = np.repeat(np.arange(len(ends)), lens)
        offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in cnn_feat])])
        feats = torch.as_tensor(np.concatenate(cnn_feat), dtype=p.dtype, device=p.device)
        meta = v_metas[ends].astype(np.int64)
        base = offs[meta[:, 0]] + meta[:, 1] - m
        if base.min() < 0:
            raise IndexError('episode context window leaves the take (video_forecast_net.py:82)')
        frame = base[None, :] + np.arange(m)[:, None]
        self._ep = dict(n=n, tmax=tmax, alive=alive, rows=rows,
                        ep_of=torch.as_tensor(ep_of, dtype=torch.long, device=p.device),
                        cnn_ctx=feats[torch.as_tensor(frame, device=p.device)])      # [m, n_ep, F]

    def train_context(self, states):
        """-> [N, v_hdim + s_hdim] = cat(v_out of the row's episode, state-LSTM output at the row)"""
        e = self._ep
        v_ep = self.forward_v_net(e['cnn_ctx'])[-1]                                  # [n_ep, v_hdim]
        v_out = v_ep.index_select(0, e['ep_of'])
        if self.s_net_type != 'lstm':
            return torch.cat((v_out, states), dim=1)
        cell = self.s_net.rnn_f
        H = cell.hidden_size
        xi = torch.addmm(cell.bias_ih + cell.bias_hh, states, cell.weight_ih.t())    # input projection of every row
        if _fused_lstm_ok(cell, states):
            # fused recurrence over the ragged episodes: rows gathered time-major (episodes sorted longest first)
            if 'packed' not in e:
                alive = np.asarray(e['alive'][:e['tmax']], dtype=np.int64)
                off = np.concatenate([[0], np.cumsum(alive)])
                prev = np.full(int(off[-1]), -1, dtype=np.int64)
                for t in range(1, e['tmax']):
                    prev[off[t]:off[t + 1]] = off[t - 1] + np.arange(alive[t])
                e['packed'] = dict(order=torch.cat(e['rows']), off=torch.as_tensor(off, device=states.device),
                                   prev=torch.as_tensor(prev, device=states.device), B=int(alive[0]))
            pk = e['packed']
            hs = _LstmRecurrence.apply(xi.index_select(0, pk['order']), cell.weight_hh, pk['off'], pk['prev'], e['tmax'], pk['B'])
            s_out = states.new_zeros((e['n'], H)).index_copy(0, pk['order'], hs)
            return torch.cat((v_out, s_out), dim=1)
        whh_t = cell.weight_hh.t()
        h = states.new_zeros((int(e['alive'][0]), H))
        c = torch.zeros_like(h)
        outs = []
        for t in range(e['tmax']):
            na = int(e['alive'][t])
            h, c = h[:na], c[:na]
            gates = xi.index_select(0, e['rows'][t]) + h @ whh_t
            i, f, g, o = gates.chunk(4, 1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        s_out = states.new_zeros((e['n'], H)).index_copy(0, torch.cat(e['rows']), torch.cat(outs))
        return torch.cat((v_out, s_out), dim=1)

    def forward(self, x):
        if self.mode == 'test':
            if self.s_net_type == 'lstm':
                x = self.s_net(x)
            x = torch.cat((self.v_out.to(x.device), x), dim=1)
            self.t += 1
            return x
        return self.train_context(x)

    @torch.no_grad()
    def context_table(self, cnn_feat, episode_len, max_batch=8192):
        """-> (table [n_windows, v_hdim], win_off int32 [n_takes + 1]); window w of take k starts at frame
        start = v_margin + (w - win_off[k]) (the range reset_model samples from, humanoid_v1.py:214)."""
        p, m, T = self._p(), self.v_margin, int(episode_len)
        rows, win_off = [], [0]
        for feat in cnn_feat:
            f = torch.as_tensor(feat, dtype=p.dtype, device=p.device)
            nwin = f.shape[0] - T - 2 * m
            if nwin <= 0:
                raise ValueError('take shorter than episode_len + 2 * fr_margin')
            for lo in range(0, nwin, max_batch):
                hi = min(nwin, lo + max_batch)
                idx = torch.arange(m, device=p.device)[:, None] + torch.arange(lo, hi, device=p.device)[None, :]
                rows.append(self.forward_v_net(f[idx])[-1])
            win_off.append(win_off[-1] + nwin)
        return torch.cat(rows).contiguous(), torch.tensor(win_off, dtype=torch.int32, device=p.device)

    @torch.no_grad()
    def snet_packed(self):
        """state-LSTM weights for egp_rollout_f64 (include/egopose_b200.h, EgpRolloutIn.d_snet_W / d_snet_b):
        row 4u + g = cat(weight_ih[g H + u], weight_hh[g H + u]), bias = bias_ih + bias_hh, g = (i, f, g, o)"""
        if self.s_net_type != 'lstm':
            return None
        cell = self.s_net.rnn_f
        H = cell.hidden_size
        W = torch.cat((cell.weight_ih, cell.weight_hh), dim=1).view(4, H, -1).transpose(0, 1).reshape(4 * H, -1)
        b = (cell.bias_ih + cell.bias_hh).view(4, H).t().reshape(-1)
        return W.contiguous(), b.contiguous(), H


def trunk_ok(net):
    """the fused kernels implement the two-hidden-layer relu trunk every shipped yml selects
    (config/egomimic/subject_03.yml:14-16,21-23)"""
    return isinstance(net, MLP) and len(net.affine_layers) == 2 and net.activation_name == 'relu'


LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))
