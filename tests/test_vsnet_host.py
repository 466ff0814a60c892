"""VideoStateNet / RNN mirrors (CPU, torch) vs the golden produced by the reference's own
models/video_state_net.py + models/rnn.py (tests/golden/make_golden.py gen_vsnet)."""
import numpy as np
import torch


def _load(net, g, prefix):
    net.load_state_dict({k[len(prefix) + 1:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix + '.')})


def test_state_dict_keys_and_test_mode(golden):
    from egopose_b200.nets import VideoStateNet
    torch.set_default_dtype(torch.float64)
    g = golden('vsnet_small')
    F, VH, M, S, A, T = [int(x) for x in g['dims']]
    net = VideoStateNet(F, VH, M)
    assert sorted(net.state_dict().keys()) == sorted(k[5:] for k in g.files if k.startswith('pvs0.'))
    _load(net, g, 'pvs0')
    cnn = [torch.from_numpy(c) for c in g['cnn_feat']]
    net.set_mode('test')
    with torch.no_grad():
        net.initialize(cnn[1][5 - M: 5 + T + M])
    assert np.allclose(net.v_out.numpy(), g['test_v_out'], rtol=1e-12, atol=1e-13)
    x = net(torch.zeros(1, S))
    assert x.shape == (1, VH + S) and net.t == 1
    # all-windows table: row (win, t) == test-mode v_out[t] of the episode started at frame M + win
    table, win_off = net.context_table([c.numpy() for c in cnn], T)
    nwin = 30 - T - 2 * M
    assert win_off.tolist() == [0, nwin, 2 * nwin, 3 * nwin] and table.shape == (3 * nwin * T, VH)
    w = int(win_off[1]) + (5 - M)
    assert np.allclose(table[w * T:(w + 1) * T].numpy(), g['test_v_out'], rtol=1e-12, atol=1e-13)


def test_train_mode_context_matches_reference_packing(golden):
    """train-mode rows equal a per-episode BiLSTM over [start - m, start + Tmax + m) (video_state_net.py:53-57)"""
    from egopose_b200.nets import VideoStateNet
    torch.set_default_dtype(torch.float64)
    g = golden('vsnet_small')
    F, VH, M, S, A, T = [int(x) for x in g['dims']]
    net = VideoStateNet(F, VH, M)
    _load(net, g, 'pvs0')
    cnn = [c for c in g['cnn_feat']]
    masks, v_metas = g['batch.masks'], g['batch.v_metas']
    net.set_mode('train')
    net.initialize((torch.from_numpy(masks), cnn, v_metas))
    ctx = net.train_context().detach().numpy()
    ends = np.nonzero(masks == 0)[0]
    tmax = int(np.diff(np.concatenate([[-1], ends])).max())
    ref_net = VideoStateNet(F, VH, M)
    _load(ref_net, g, 'pvs0')
    ref_net.set_mode('test')
    row = 0
    for e, end in enumerate(ends):
        take, start = v_metas[end]
        with torch.no_grad():
            full = ref_net.forward_v_net(torch.from_numpy(cnn[take][start - M: start + tmax + M]).unsqueeze(1)).squeeze(1)[M:-M]
        n = end - row + 1
        assert np.allclose(ctx[row:row + n], full[:n].numpy(), rtol=1e-12, atol=1e-13)
        row = end + 1
    out = net(torch.from_numpy(g['batch.states']))
    assert out.shape == (masks.shape[0], VH + S)
