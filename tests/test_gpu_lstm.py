"""Fused LSTM sequence kernels (csrc/lstm.cu, egp_lstm_seq_fwd / bwd_f64) vs the eager per-step LSTMCell loop of the torch
mirror - which tests/test_vsnet_host.py / test_fcnet_host.py pin to the reference's own models/rnn.py, video_state_net.py and
video_forecast_net.py - at the reference's real sizes: BiLSTM 128 -> 2 x 64 over T + 2 m frames (egomimic VideoStateNet),
state LSTM 115 -> 128 over ragged episodes (egoforecast VideoForecastNet.s_net).  Forward values and every parameter gradient."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def _both(fn):
    """run fn() with the fused kernels and with the eager loop"""
    out = {}
    for mode in ('fused', 'eager'):
        os.environ['EGP_LSTM'] = mode
        try:
            out[mode] = fn()
        finally:
            os.environ.pop('EGP_LSTM', None)
    return out['fused'], out['eager']


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-300))


@pytest.mark.parametrize('L,B,bi', [(25, 70, True), (220, 40, True), (30, 33, False), (1, 5, True)])
def test_rnn_sweep_fused_matches_eager(L, B, bi):
    from egopose_b200.nets import RNN
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(L + B)
    net = RNN(128, 128, 'lstm', bi_dir=bi).cuda()
    x = torch.randn(L, B, 128, device='cuda')
    wgt = torch.randn(L, B, 128, device='cuda')

    def run():
        for p in net.parameters():
            p.grad = None
        out = net(x)
        (out * wgt).sum().backward()
        return out.detach().clone(), {n: p.grad.clone() for n, p in net.named_parameters()}
    (of, gf), (oe, ge) = _both(run)
    assert _rel(of, oe) < 1e-12
    for n in ge:
        assert _rel(gf[n], ge[n]) < 1e-10, n


def test_video_state_net_train_context_fused_matches_eager():
    """egomimic train mode (video_state_net.py:40-59,65-69): padded [Tmax + 2m, n_ep, 128] context, BiLSTM, row gather"""
    from egopose_b200.nets import VideoStateNet
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(5)
    rng = np.random.RandomState(2)
    net = VideoStateNet(128, 128, 10, 'lstm', None, False).cuda()
    cnn = [rng.randn(120, 128), rng.randn(90, 128)]
    lens = rng.randint(1, 30, size=37)
    N = int(lens.sum())
    masks = np.ones(N)
    masks[np.cumsum(lens) - 1] = 0
    v_metas = np.zeros((N, 2), dtype=np.int64)
    pos = 0
    for ln in lens:
        take = rng.randint(0, 2)
        v_metas[pos:pos + ln] = (take, rng.randint(10, cnn[take].shape[0] - 30 - 10))
        pos += ln
    wgt = torch.randn(N, 128, device='cuda')
    net.set_mode('train')
    net.initialize((torch.as_tensor(masks, device='cuda'), cnn, v_metas))

    def run():
        for p in net.parameters():
            p.grad = None
        ctx = net.train_context()
        (ctx * wgt).sum().backward()
        return ctx.detach().clone(), {n: p.grad.clone() for n, p in net.named_parameters()}
    (of, gf), (oe, ge) = _both(run)
    assert of.shape == (N, 128) and _rel(of, oe) < 1e-12
    for n in ge:
        assert _rel(gf[n], ge[n]) < 1e-10, n


def test_video_forecast_net_state_lstm_fused_matches_eager():
    """egoforecast train mode (video_forecast_net.py:95-111): state LSTM 115 -> 128 unrolled over ragged episodes"""
    from egopose_b200.nets import VideoForecastNet
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(7)
    rng = np.random.RandomState(3)
    net = VideoForecastNet(128, 115, 128, 30, 'lstm', None, 128, 'lstm').cuda()
    cnn = [rng.randn(200, 128), rng.randn(170, 128)]
    lens = np.concatenate([rng.randint(1, 20, size=60), [90, 90, 1]])
    N = int(lens.sum())
    masks = np.ones(N)
    masks[np.cumsum(lens) - 1] = 0
    v_metas = np.zeros((N, 2), dtype=np.int64)
    pos = 0
    for ln in lens:
        take = rng.randint(0, 2)
        v_metas[pos:pos + ln] = (take, rng.randint(30, cnn[take].shape[0] - 90 - 30))
        pos += ln
    states = torch.randn(N, 115, device='cuda')
    wgt = torch.randn(N, 256, device='cuda')
    net.set_mode('train')
    net.initialize((torch.as_tensor(masks, device='cuda'), cnn, v_metas))

    def run():
        for p in net.parameters():
            p.grad = None
        ctx = net.train_context(states)
        (ctx * wgt).sum().backward()
        return ctx.detach().clone(), {n: p.grad.clone() for n, p in net.named_parameters()}
    (of, gf), (oe, ge) = _both(run)
    assert of.shape == (N, 256) and _rel(of, oe) < 1e-12
    for n in ge:
        assert _rel(gf[n], ge[n]) < 1e-10, n
