"""VideoForecastNet mirror (CPU, torch) vs the golden produced by the reference's own
models/video_forecast_net.py + models/rnn.py (tests/golden/make_golden.py gen_fcnet)."""
import numpy as np
import torch


def _load(net, g, prefix):
    net.load_state_dict({k[len(prefix) + 1:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix + '.')})


def _net(g, prefix='pvs0'):
    from egopose_b200.nets import VideoForecastNet
    torch.set_default_dtype(torch.float64)
    F, VH, M, S, A, T, SH = [int(x) for x in g['dims']]
    net = VideoForecastNet(F, S, VH, M, 'lstm', None, SH, 'lstm', False)
    assert sorted(net.state_dict().keys()) == sorted(k[len(prefix) + 1:] for k in g.files if k.startswith(prefix + '.'))
    _load(net, g, prefix)
    return net


def test_test_mode_steps_and_window_table(golden):
    g = golden('fcnet_small')
    F, VH, M, S, A, T, SH = [int(x) for x in g['dims']]
    net = _net(g)
    assert net.out_dim == VH + SH
    cnn = [torch.from_numpy(c) for c in g['cnn_feat']]
    net.set_mode('test')
    with torch.no_grad():
        net.initialize(cnn[1][9 - M: 9 + T + M])
        assert np.allclose(net.v_out.numpy(), g['test_v_out'], rtol=1e-12, atol=1e-13)
        x = np.concatenate([net(torch.from_numpy(g['test_states'][[i]])).numpy() for i in range(5)])
    assert np.allclose(x, g['test_x'], rtol=1e-12, atol=1e-13) and net.t == 5
    # one row per (take, start) window == test-mode v_out of an episode started there
    table, win_off = net.context_table([c.numpy() for c in cnn], T)
    nwin = 34 - T - 2 * M
    assert win_off.tolist() == [0, nwin, 2 * nwin, 3 * nwin] and table.shape == (3 * nwin, VH)
    assert np.allclose(table[int(win_off[1]) + 9 - M].numpy(), g['test_v_out'][0], rtol=1e-12, atol=1e-13)


def test_snet_packed_row_order(golden):
    """row 4u + g of the packed matrix reproduces gate g of unit u (the order the rollout kernel steps in)"""
    g = golden('fcnet_small')
    net = _net(g)
    W, b, H = net.snet_packed()
    cell = net.s_net.rnn_f
    x = torch.randn(3, cell.input_size)
    h = torch.randn(3, H)
    gates = torch.cat((x, h), 1) @ W.t() + b
    ref = x @ cell.weight_ih.t() + cell.bias_ih + h @ cell.weight_hh.t() + cell.bias_hh
    for gi in range(4):
        assert torch.allclose(gates[:, gi::4], ref[:, gi * H:(gi + 1) * H], rtol=1e-12, atol=1e-13)


def test_train_mode_matches_reference_padded_unroll(golden):
    """length-sorted unroll == the reference's padded [Tmax, n_ep] unroll (video_forecast_net.py:94-107)"""
    g = golden('fcnet_small')
    net = _net(g)
    net.set_mode('train')
    net.initialize((torch.from_numpy(g['batch.masks']), [c for c in g['cnn_feat']], g['batch.v_metas']))
    out = net(torch.from_numpy(g['batch.states']))
    assert np.allclose(out.detach().numpy(), g['train_x'], rtol=1e-11, atol=1e-12)
    # gradients reach both LSTMs
    out.sum().backward()
    assert net.v_net.rnn_f.weight_ih.grad.abs().sum() > 0 and net.s_net.rnn_f.weight_hh.grad.abs().sum() > 0
