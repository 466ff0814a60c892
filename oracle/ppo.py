"""TEST INFRASTRUCTURE ONLY - CPU float64 restatement of the reference's PPO update half.

Pinned against the reference's own code through tests/golden/ppo_small.npz (tests/test_oracle_ppo.py).
Never imported by egopose_b200/.  Uses torch CPU autograd + torch.optim.Adam exactly as the reference
does (ego_pose/ego_mimic.py:70-77), so it also serves as the CPU baseline of the update phase.
"""
import math

import numpy as np
import torch


def gae(rewards, masks, values, gamma, tau):
    """core/common.py:5-25 estimate_advantages: flat reverse scan, carries zeroed by masks;
    returns (standardised advantages [N], returns [N]); std is the unbiased torch.std (:22)."""
    r = np.asarray(rewards, dtype=np.float64).ravel().tolist()
    m = np.asarray(masks, dtype=np.float64).ravel().tolist()
    v = np.asarray(values, dtype=np.float64).ravel().tolist()
    n = len(r)
    adv = [0.0] * n
    prev_v = 0.0
    prev_a = 0.0
    for i in range(n - 1, -1, -1):
        delta = r[i] + gamma * prev_v * m[i] - v[i]
        adv[i] = delta + gamma * tau * prev_a * m[i]
        prev_v = v[i]
        prev_a = adv[i]
    adv = np.array(adv)
    ret = np.asarray(values, dtype=np.float64).ravel() + adv
    adv_n = (adv - adv.mean()) / adv.std(ddof=1)
    return adv_n, ret


def mlp_forward(x, layers):
    """models/mlp.py:22-25 with relu; ``layers`` = [(W, b), ...] hidden layers (torch [out, in])."""
    for W, b in layers:
        x = torch.relu(torch.addmm(b, x, W.t()))
    return x


def policy_mean(x, p):
    """core/policy_gaussian.py:19-24 action mean; p = dict of state-dict tensors."""
    h = mlp_forward(x, _hidden(p))
    return torch.addmm(p['action_mean.bias'], h, p['action_mean.weight'].t())


def value_forward(x, p):
    """core/critic.py:15-18"""
    h = mlp_forward(x, _hidden(p))
    return torch.addmm(p['value_head.bias'], h, p['value_head.weight'].t())


def _hidden(p):
    out, i = [], 0
    while 'net.affine_layers.%d.weight' % i in p:
        out.append((p['net.affine_layers.%d.weight' % i], p['net.affine_layers.%d.bias' % i]))
        i += 1
    return out


def log_prob(mean, log_std, actions):
    """core/distributions.py:21-22 (Normal.log_prob summed over action dims)"""
    std = torch.exp(log_std)
    var = std * std
    lp = -((actions - mean) ** 2) / (2 * var) - log_std - math.log(math.sqrt(2 * math.pi))
    return lp.sum(1, keepdim=True)


def ppo_update(pol, val, states, actions, returns, advantages, exps, clip_epsilon, lr_p, lr_v, max_norm, epochs,
               fix_std=True, threads=None, states_v=None, mini_batch=None, perm_rng=None):
    """agents/agent_ppo.py:16-51 (full-batch branch) + agent_pg.py:19-26 update_value + :53-56 clip.
    ``pol`` / ``val`` are dicts name -> float64 numpy arrays in state-dict layout; returns
    (new_pol, new_val, info) with per-epoch surrogate loss, pre-step value loss and policy grad norm."""
    if threads:
        torch.set_num_threads(threads)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))  # noqa: E731
    P = {k: tt(v).clone().requires_grad_(not (fix_std and k == 'action_log_std')) for k, v in pol.items()}
    V = {k: tt(v).clone().requires_grad_(True) for k, v in val.items()}
    pparams = [t for t in P.values() if t.requires_grad]
    opt_p = torch.optim.Adam(pparams, lr=lr_p)
    opt_v = torch.optim.Adam(list(V.values()), lr=lr_v)
    st, ac = tt(states), tt(actions)
    stv = st if states_v is None else tt(states_v)
    ret, adv = tt(returns).reshape(-1, 1), tt(advantages).reshape(-1, 1)
    ind = torch.from_numpy(np.asarray(exps).ravel()).nonzero().squeeze(1)
    with torch.no_grad():
        fixed_lp = log_prob(policy_mean(st, P), P['action_log_std'], ac)
    info = dict(surr_loss=[], value_loss=[], grad_norm=[], fixed_log_probs=fixed_lp.numpy().copy())
    if mini_batch:
        # agents/agent_ppo.py:24-43: cumulative np.random.shuffle permutations, contiguous slices, per-slice means
        rng = perm_rng if perm_rng is not None else np.random
        ex = torch.from_numpy(np.asarray(exps, dtype=np.float64).ravel())
        n = st.shape[0]
        for _ in range(epochs):
            perm = np.arange(n)
            rng.shuffle(perm)
            perm = torch.from_numpy(perm)
            st, stv, ac, ret, adv, fixed_lp, ex = st[perm], stv[perm], ac[perm], ret[perm], adv[perm], fixed_lp[perm], ex[perm]
            for i in range(int(math.ceil(n / mini_batch))):
                sl = slice(i * mini_batch, min((i + 1) * mini_batch, n))
                bi = ex[sl].nonzero().squeeze(1)
                vloss = (value_forward(stv[sl], V) - ret[sl]).pow(2).mean()
                opt_v.zero_grad()
                vloss.backward()
                opt_v.step()
                lp = log_prob(policy_mean(st[sl][bi], P), P['action_log_std'], ac[sl][bi])
                ratio = torch.exp(lp - fixed_lp[sl][bi])
                a = adv[sl][bi]
                surr = -torch.min(ratio * a, torch.clamp(ratio, 1.0 - clip_epsilon, 1.0 + clip_epsilon) * a).mean()
                opt_p.zero_grad()
                surr.backward()
                if max_norm is not None:
                    torch.nn.utils.clip_grad_norm_(pparams, max_norm)
                opt_p.step()
                info['surr_loss'].append(surr.item())
                info['value_loss'].append(vloss.item())
        return ({k: v.detach().numpy() for k, v in P.items()}, {k: v.detach().numpy() for k, v in V.items()}, info)
    for _ in range(epochs):
        vloss = (value_forward(stv, V) - ret).pow(2).mean()
        opt_v.zero_grad()
        vloss.backward()
        opt_v.step()
        lp = log_prob(policy_mean(st[ind], P), P['action_log_std'], ac[ind])
        ratio = torch.exp(lp - fixed_lp[ind])
        a = adv[ind]
        surr = -torch.min(ratio * a, torch.clamp(ratio, 1.0 - clip_epsilon, 1.0 + clip_epsilon) * a).mean()
        opt_p.zero_grad()
        surr.backward()
        gn = torch.sqrt(sum((p.grad ** 2).sum() for p in pparams))
        if max_norm is not None:
            torch.nn.utils.clip_grad_norm_(pparams, max_norm)
        opt_p.step()
        info['surr_loss'].append(surr.item())
        info['value_loss'].append(vloss.item())
        info['grad_norm'].append(gn.item())
    return ({k: v.detach().numpy() for k, v in P.items()}, {k: v.detach().numpy() for k, v in V.items()}, info)


def zfilter_sequence(xs, clip=5.0):
    """utils/zfilter.py:7-67: Welford push then (x-mean)/(std+1e-8), clip."""
    n, M, S = 0, None, None
    ys = []
    for x in np.asarray(xs, dtype=np.float64):
        n += 1
        if n == 1:
            M, S = x.copy(), np.zeros_like(x)
        else:
            old = M.copy()
            M = old + (x - old) / n
            S = S + (x - old) * (x - M)
        var = S / (n - 1) if n > 1 else np.square(M)
        y = (x - M) / (np.sqrt(var) + 1e-8)
        ys.append(np.clip(y, -clip, clip))
    return np.stack(ys), n, M, S


def zfilter_merge(n_a, mean_a, S_a, xs):
    """Chan et al. parallel merge of running moments with a batch xs [k, dim] - the documented batched
    replacement for the reference's sequential worker-0 updates (SURVEY 7 'ZFilter semantics')."""
    xs = np.asarray(xs, dtype=np.float64)
    n_b = xs.shape[0]
    mean_b = xs.mean(0)
    S_b = ((xs - mean_b) ** 2).sum(0)
    if n_a == 0:
        return n_b, mean_b, S_b
    n = n_a + n_b
    delta = mean_b - mean_a
    return n, mean_a + delta * n_b / n, S_a + S_b + delta * delta * n_a * n_b / n
