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


def trunk_ok(net):
    """the fused kernels implement the two-hidden-layer relu trunk every shipped yml selects
    (config/egomimic/subject_03.yml:14-16,21-23)"""
    return isinstance(net, MLP) and len(net.affine_layers) == 2 and net.activation_name == 'relu'


LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))
