"""CUDA update kernels (through the C ABI) vs the CPU oracle and the reference-generated golden fixture.
Tolerances: float64 arithmetic on both sides; the north-star bound is 1e-5 on the PPO loss."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from oracle import ppo as oppo  # noqa: E402
import helpers  # noqa: E402


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda')


@pytest.mark.parametrize('n', [1, 5, 1023, 1024, 1025, 2047, 2048, 2049, 4096 * 3 + 17, 300 * 700])
def test_gae_vs_oracle(n):
    from egopose_b200 import lib
    rng = np.random.RandomState(n)
    r, v = rng.rand(n), rng.randn(n)
    m = (rng.rand(n) > 0.02).astype(np.float64)
    if n > 300:
        m[299::300] = 0.0
    m[-1] = 0.0
    adv, ret, stats = lib.gae(cu(r), cu(m), cu(v), 0.95, 0.95)
    if n > 1:
        adv_n, ret_o = oppo.gae(r, m, v, 0.95, 0.95)
        got = lib.standardize_(adv.clone(), stats).cpu().numpy()
        assert np.allclose(got, adv_n, rtol=1e-10, atol=1e-11)
        assert np.allclose(ret.cpu().numpy(), ret_o, rtol=1e-12, atol=1e-12)
    else:
        assert abs(adv.item() - (r[0] - v[0])) < 1e-14


def test_gae_no_episode_boundaries_long_chain():
    """masks all one: the look-back must chain through every tile (no early exit)."""
    from egopose_b200 import lib
    n = 2048 * 40 + 3
    rng = np.random.RandomState(0)
    r, v, m = rng.rand(n), rng.randn(n), np.ones(n)
    adv, ret, stats = lib.gae(cu(r), cu(m), cu(v), 0.99, 0.97)
    adv_n, ret_o = oppo.gae(r, m, v, 0.99, 0.97)
    assert np.allclose(ret.cpu().numpy(), ret_o, rtol=1e-11, atol=1e-11)
    assert np.allclose(lib.standardize_(adv, stats).cpu().numpy(), adv_n, rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize('n,mask', [(1, 'rand'), (2049, 'rand'), (300 * 700, 'episodes'), (2048 * 40 + 3, 'ones'),
                                    (2048 * 700 + 11, 'rand'), (6_000_001, 'episodes')])
def test_gae_one_pass_is_bit_identical_to_two_pass(n, mask):
    """the ticketed one-pass scan (large batches) and the two-pass scan do the same per-tile arithmetic in the same merge
    order: advantages, returns and moments must agree bit for bit; also against the oracle"""
    from egopose_b200 import lib
    rng = np.random.RandomState(7)
    r, v = rng.rand(n), rng.randn(n)
    m = np.ones(n) if mask == 'ones' else (rng.rand(n) > 0.02).astype(np.float64)
    if mask == 'episodes':
        m[149::150] = 0.0
    m[-1] = 0.0
    old = lib.gae_set_onepass_min(1 << 62)
    try:
        a2, r2, s2 = lib.gae(cu(r), cu(m), cu(v), 0.95, 0.95)
        lib.gae_set_onepass_min(0)
        for _ in range(3):                              # repeated calls reuse nothing between launches
            a1, r1, s1 = lib.gae(cu(r), cu(m), cu(v), 0.95, 0.95)
            assert torch.equal(a1, a2) and torch.equal(r1, r2) and torch.equal(s1, s2)
    finally:
        lib.gae_set_onepass_min(old)
    if 1 < n < 1_000_000:
        adv_n, ret_o = oppo.gae(r, m, v, 0.95, 0.95)
        assert np.allclose(r1.cpu().numpy(), ret_o, rtol=1e-11, atol=1e-11)
        assert np.allclose(lib.standardize_(a1, s1).cpu().numpy(), adv_n, rtol=1e-9, atol=1e-10)


def test_gae_unaligned_views():
    """8-byte-offset tensor views take the scalar load path"""
    from egopose_b200 import lib
    n = 5000
    rng = np.random.RandomState(1)
    r, v = rng.rand(n + 1), rng.randn(n + 1)
    m = (rng.rand(n + 1) > 0.05).astype(np.float64)
    m[-1] = 0
    adv, ret, stats = lib.gae(cu(r)[1:], cu(m)[1:], cu(v)[1:], 0.95, 0.9)
    adv_n, ret_o = oppo.gae(r[1:], m[1:], v[1:], 0.95, 0.9)
    assert np.allclose(ret.cpu().numpy(), ret_o, rtol=1e-12, atol=1e-12)
    assert np.allclose(lib.standardize_(adv, stats).cpu().numpy(), adv_n, rtol=1e-10, atol=1e-11)


def test_gae_golden(golden):
    from egopose_b200 import lib
    g = golden('ppo_small')
    adv, ret, stats = lib.gae(cu(g['rewards']), cu(g['masks']), cu(g['values0'].ravel()), 0.95, 0.95)
    assert np.allclose(ret.cpu().numpy(), g['returns'].ravel(), rtol=1e-12, atol=1e-12)
    assert np.allclose(lib.standardize_(adv, stats).cpu().numpy(), g['advantages'].ravel(), rtol=1e-10, atol=1e-11)


def test_logp_and_loss_golden(golden):
    """fixed log-probs, surrogate loss and its gradient wrt mu vs torch autograd of the reference formula."""
    from egopose_b200 import lib
    g = golden('ppo_small')
    pol = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('p0.')}
    st, ac = torch.from_numpy(g['states']), torch.from_numpy(g['actions'])
    mu = oppo.policy_mean(st, pol)
    ls = pol['action_log_std']
    logp0 = lib.gauss_logp(cu(mu.numpy()), cu(g['actions']), cu(ls.numpy().ravel()))
    assert np.allclose(logp0.cpu().numpy(), g['fixed_log_probs'].ravel(), rtol=1e-12, atol=1e-12)
    # perturbed mu so that ratios leave the clip range on both sides
    rng = np.random.RandomState(4)
    mu2 = (mu + 0.03 * torch.from_numpy(rng.randn(*mu.shape))).requires_grad_(True)
    ls2 = ls.clone().requires_grad_(True)
    adv_raw, ret, stats = lib.gae(cu(g['rewards']), cu(g['masks']), cu(g['values0'].ravel()), 0.95, 0.95)
    adv = torch.from_numpy(g['advantages'])
    ind = torch.from_numpy(g['exps']).nonzero().squeeze(1)
    lp = oppo.log_prob(mu2[ind], ls2, ac[ind])
    ratio = torch.exp(lp - torch.from_numpy(g['fixed_log_probs'])[ind])
    a = adv[ind]
    assert ((ratio < 0.8).sum() > 5) and ((ratio > 1.2).sum() > 5)
    surr = -torch.min(ratio * a, torch.clamp(ratio, 0.8, 1.2) * a).mean()
    surr.backward()
    dmu = torch.zeros_like(cu(g['actions']))
    dls = torch.zeros(mu.shape[1], dtype=torch.float64, device='cuda')
    loss = torch.zeros(1, dtype=torch.float64, device='cuda')
    lib.ppo_loss_grad(cu(mu2.detach().numpy()), cu(g['actions']), cu(ls.numpy().ravel()), adv_raw, stats, logp0,
                      cu(g['exps']), 0.2, 1.0 / len(ind), dmu, dls, loss)
    assert abs(loss.item() - surr.item()) < 1e-12 * max(1, abs(surr.item())) + 1e-13
    assert np.allclose(dmu.cpu().numpy(), mu2.grad.numpy(), rtol=1e-9, atol=1e-13)
    assert np.allclose(dls.cpu().numpy(), ls2.grad.numpy().ravel(), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('n,adim', [(1, 52), (31, 4), (33, 52), (257, 17), (4099, 52)])
def test_loss_kernels_ragged_sizes(n, adim):
    """row counts that leave half a warp without a row, action dims that are not a multiple of the 16 lanes per row
    (regression: full-mask shuffles under a divergent row loop)"""
    from egopose_b200 import lib
    rng = np.random.RandomState(n + adim)
    mu, ac = rng.randn(n, adim) * 0.1, rng.randn(n, adim) * 0.2
    ls = rng.uniform(-1.5, -0.5, size=adim)
    adv = rng.randn(n) * 2 + 1
    exps = (rng.rand(n) > 0.2).astype(np.float64)
    exps[0] = 1.0
    mu0 = mu + 0.05 * rng.randn(n, adim)
    lp0 = oppo.log_prob(torch.from_numpy(mu0), torch.from_numpy(ls), torch.from_numpy(ac))
    logp0 = lib.gauss_logp(cu(mu0), cu(ac), cu(ls))
    assert np.allclose(logp0.cpu().numpy(), lp0.numpy().ravel(), rtol=1e-12, atol=1e-12)
    mean, m2 = adv.mean(), ((adv - adv.mean()) ** 2).sum()
    stats = cu(np.array([n, mean, m2 if n > 1 else 1.0, 0.0]))
    nn_ = max(n - 1, 1)
    a_std = (torch.from_numpy(adv) - mean) / np.sqrt((m2 if n > 1 else 1.0) / nn_) if n > 1 else torch.from_numpy(adv - mean)
    mu_t = torch.from_numpy(mu).requires_grad_(True)
    ls_t = torch.from_numpy(ls).requires_grad_(True)
    ind = torch.from_numpy(exps).nonzero().squeeze(1)
    ratio = torch.exp(oppo.log_prob(mu_t[ind], ls_t, torch.from_numpy(ac)[ind]) - lp0[ind])
    a = a_std[ind].view(-1, 1)
    surr = -torch.min(ratio * a, torch.clamp(ratio, 0.8, 1.2) * a).mean()
    surr.backward()
    dmu = torch.zeros(n, adim, dtype=torch.float64, device='cuda')
    dls = torch.zeros(adim, dtype=torch.float64, device='cuda')
    loss = torch.zeros(1, dtype=torch.float64, device='cuda')
    if n == 1:
        stats = cu(np.array([2.0, mean, 1.0, 0.0]))        # (adv - mean) / sqrt(1 / 1): same standardisation as a_std
    lib.ppo_loss_grad(cu(mu), cu(ac), cu(ls), cu(adv), stats, logp0, cu(exps), 0.2, 1.0 / len(ind), dmu, dls, loss)
    assert abs(loss.item() - surr.item()) < 1e-11 * max(1, abs(surr.item()))
    assert np.allclose(dmu.cpu().numpy(), mu_t.grad.numpy(), rtol=1e-9, atol=1e-13)
    assert np.allclose(dls.cpu().numpy(), ls_t.grad.numpy(), rtol=1e-9, atol=1e-12)


def test_value_loss_and_helpers():
    from egopose_b200 import lib
    rng = np.random.RandomState(2)
    n = 5000
    v, r = rng.randn(n), rng.randn(n)
    dv = torch.empty(n, dtype=torch.float64, device='cuda')
    loss = torch.zeros(1, dtype=torch.float64, device='cuda')
    lib.value_loss_grad(cu(v), cu(r), 1.0 / n, dv, loss)
    assert abs(loss.item() - np.mean((v - r) ** 2)) < 1e-12
    assert np.allclose(dv.cpu().numpy(), 2 * (v - r) / n, rtol=1e-14)
    y, b = rng.randn(777, 300), rng.randn(300)
    yy = lib.bias_relu_(cu(y), cu(b))
    ref = np.maximum(y + b, 0)
    assert np.array_equal(yy.cpu().numpy(), ref)
    dy = rng.randn(777, 300)
    assert np.array_equal(lib.relu_bwd_(cu(dy), yy).cpu().numpy(), dy * (ref > 0))
    fused_out = torch.empty(300, dtype=torch.float64, device='cuda')
    fused = lib.relu_bwd_colsum_(cu(dy), yy, fused_out)
    assert np.array_equal(fused.cpu().numpy(), dy * (ref > 0))
    assert np.allclose(fused_out.cpu().numpy(), (dy * (ref > 0)).sum(0), rtol=1e-12, atol=1e-12)
    cs = lib.colsum(cu(dy), torch.empty(300, dtype=torch.float64, device='cuda'))
    assert np.allclose(cs.cpu().numpy(), dy.sum(0), rtol=1e-12, atol=1e-12)
    sh = rng.randn(300)
    mo = lib.col_moments(cu(y), cu(sh)).cpu().numpy()
    assert np.allclose(mo[:300], (y - sh).sum(0), rtol=1e-12, atol=1e-11)
    assert np.allclose(mo[300:], ((y - sh) ** 2).sum(0), rtol=1e-12)


def test_clip_adam_vs_torch():
    from egopose_b200 import lib
    rng = np.random.RandomState(3)
    n = 143904
    p0 = rng.randn(n) * 0.1
    for max_norm in (0.0, 40.0, 0.5):
        p_ref = torch.nn.Parameter(torch.from_numpy(p0.copy()))
        opt = torch.optim.Adam([p_ref], lr=5e-5)
        p, m, v = cu(p0), cu(np.zeros(n)), cu(np.zeros(n))
        nrm = torch.zeros(1, dtype=torch.float64, device='cuda')
        for step in range(1, 4):
            g = rng.randn(n) * 0.01 * step
            p_ref.grad = torch.from_numpy(g.copy())
            if max_norm > 0:
                torch.nn.utils.clip_grad_norm_([p_ref], max_norm)
            opt.step()
            gg = cu(g)
            lib.sumsq(gg, nrm)
            assert abs(nrm.item() - (g ** 2).sum()) < 1e-12 * (g ** 2).sum()
            lib.adam_step(p, gg, m, v, 5e-5, 0.9, 0.999, 1e-8, step, max_norm, nrm)
        assert np.allclose(p.cpu().numpy(), p_ref.detach().numpy(), rtol=1e-12, atol=1e-14)
