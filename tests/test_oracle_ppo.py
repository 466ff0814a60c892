"""PPO-half oracle (oracle/ppo.py) vs golden outputs of the reference's own estimate_advantages /
PolicyGaussian / Value / AgentPPO.update_policy / ZFilter (tests/golden/make_golden.py)."""
import numpy as np

from oracle import ppo


def test_gae_matches_reference(golden):
    g = golden('ppo_small')
    gamma, tau = g['hyper'][:2]
    adv, ret = ppo.gae(g['rewards'], g['masks'], g['values0'], gamma, tau)
    assert np.allclose(adv, g['advantages'].ravel(), rtol=1e-12, atol=1e-13)
    assert np.allclose(ret, g['returns'].ravel(), rtol=1e-13, atol=1e-13)


def test_ppo_update_matches_reference(golden):
    g = golden('ppo_small')
    gamma, tau, clip, lr_p, lr_v, max_norm = g['hyper']
    pol = {k[3:]: g[k] for k in g.files if k.startswith('p0.')}
    val = {k[3:]: g[k] for k in g.files if k.startswith('v0.')}
    new_p, new_v, info = ppo.ppo_update(pol, val, g['states'], g['actions'], g['returns'], g['advantages'], g['exps'],
                                        clip, lr_p, lr_v, max_norm, epochs=3)
    assert np.allclose(info['fixed_log_probs'], g['fixed_log_probs'], rtol=1e-12, atol=1e-12)
    assert np.allclose(info['surr_loss'], g['surr_loss'], rtol=1e-10, atol=1e-12)
    assert np.allclose(info['value_loss'], g['value_loss'], rtol=1e-11)
    assert np.allclose(info['grad_norm'], g['grad_norm'], rtol=1e-10)
    for k in new_p:
        assert np.allclose(new_p[k], g['p3.' + k], rtol=1e-9, atol=1e-11), k
    for k in new_v:
        assert np.allclose(new_v[k], g['v3.' + k], rtol=1e-9, atol=1e-11), k


def test_zfilter_matches_reference(golden):
    g = golden('zfilter')
    ys, n, M, S = ppo.zfilter_sequence(g['xs'], clip=5.0)
    assert np.allclose(ys, g['ys'], rtol=1e-12, atol=1e-12)
    assert n == int(g['n']) and np.allclose(M, g['mean'], rtol=1e-13) and np.allclose(S, g['S'], rtol=1e-12)
    # batched (Chan) merge reproduces the sequential moments
    n2, M2, S2 = ppo.zfilter_merge(0, None, None, g['xs'][:13])
    n2, M2, S2 = ppo.zfilter_merge(n2, M2, S2, g['xs'][13:])
    assert n2 == n and np.allclose(M2, M, rtol=1e-12) and np.allclose(S2, S, rtol=1e-11)


def test_ppo_minibatch_matches_reference(golden):
    g, m = golden('ppo_small'), golden('ppo_minibatch')
    gamma, tau, clip, lr_p, lr_v, max_norm = g['hyper']
    pol = {k[3:]: g[k] for k in g.files if k.startswith('p0.')}
    val = {k[3:]: g[k] for k in g.files if k.startswith('v0.')}
    new_p, new_v, info = ppo.ppo_update(pol, val, g['states'], g['actions'], g['returns'], g['advantages'], g['exps'],
                                        clip, lr_p, lr_v, max_norm, epochs=int(m['epochs']),
                                        mini_batch=int(m['opt_batch_size']), perm_rng=np.random.RandomState(int(m['seed'])))
    assert np.allclose(info['surr_loss'], m['surr_loss'], rtol=1e-10, atol=1e-12)
    for k in new_p:
        assert np.allclose(new_p[k], m['p.' + k], rtol=1e-9, atol=1e-11), k
    for k in new_v:
        assert np.allclose(new_v[k], m['v.' + k], rtol=1e-9, atol=1e-11), k
