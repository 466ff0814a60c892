"""Parity at the BENCHMARK shapes (BASELINE configs 2 / 3), not only at toy sizes:

  * the fused rollout at E = 4096 x T = 300 with the 2 x 300 policy (one full wave of 128 CTAs) and at E = 8192 with the
    2 x 512 policy (two waves, chunked layer-2/3 plan): pre-drawn noise / resets for every environment, 64 randomly chosen
    environments re-run by the CPU oracle (per-environment streams do not depend on E, so the oracle only pays for 64);
  * a full-scale PPO update (N = 1,228,800 rows, 10 epochs): float64 on the int8 tensor cores (S = 6 and 7) vs cuBLAS DGEMM
    vs torch autograd on the GPU, loss tolerance 1e-5 (north star) with the measured figures printed;
  * two consecutive updates that reuse the rollout buffers (stale input-slice cache regression);
  * N-rank == 1-rank equivalence with two ranks sharing cuda:0 (gloo on device tensors), so the 1-GPU test lease runs it.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from oracle import cphys  # noqa: E402
import helpers  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cu(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device='cuda')


def _rollout_vs_oracle(E, T, hidden, ctx_dim, n_check, seed):
    EPL = 200 if T >= 200 else 40
    n_takes, L = 4, max(T, EPL) + 84
    orc = cphys.Oracle(episode_len=EPL)
    takes = cphys.synthetic_takes(orc.md, n_takes, L, seed=seed)
    ctx = np.random.RandomState(seed + 1).randn(n_takes * L, ctx_dim)
    orc.make_expert(takes, ctx)
    orc.cfg.fix_head_lb = float('nan')
    orc.cfg.end_reward = 0.21
    model = helpers.make_model()
    model.upload_experts(orc._keep['x_rows'], orc._keep['x_off'], orc._keep['x_lb'], ctx)
    S, nu = orc.S, orc.nu
    w = helpers.policy_weights(S + ctx_dim, hidden[0], hidden[1], nu, seed=3)
    rng = np.random.RandomState(seed + 2)
    max_resets = T
    reset_take = rng.randint(0, n_takes, size=(E, max_resets)).astype(np.int32)
    reset_start = rng.randint(10, L - EPL - 10, size=(E, max_resets)).astype(np.int32)
    eps = rng.standard_normal((E * T, nu))
    zf_mean, zf_std = rng.randn(S) * 0.1, rng.uniform(0.5, 2.0, size=S)
    wd = {k: cu(v.ravel() if k == 'log_std' else v) for k, v in w.items()}
    out = model.rollout(wd, E, T, episode_len=EPL, end_reward=0.21, zf_mean=cu(zf_mean), zf_std=cu(zf_std), eps=cu(eps),
                        reset_take=cu(reset_take, torch.int32), reset_start=cu(reset_start, torch.int32))
    torch.cuda.synchronize()
    pick = np.sort(np.random.RandomState(seed + 3).choice(E, size=n_check, replace=False))
    pick[0], pick[-1] = 0, E - 1                         # first / last environment (first / last CTA) always checked
    rows = (pick[:, None] * T + np.arange(T)[None, :]).ravel()
    pol = orc.make_policy(w['W1'], w['b1'], w['W2'], w['b2'], w['W3'], w['b3'], w['log_std'])
    ref = orc.rollout(pol, n_check, T, reset_take[pick], reset_start[pick], eps[rows], None, zf_mean, zf_std, 5.0,
                      n_threads=min(16, os.cpu_count() or 1))
    g = {k: out[k][torch.as_tensor(rows, device='cuda')].cpu().numpy() for k in ('states', 'actions', 'next_states', 'rewards',
                                                                                   'masks', 'exps', 'c_info', 'v_metas')}
    assert np.array_equal(g['masks'], ref['masks'])
    assert np.array_equal(g['exps'], ref['exps'])
    assert np.array_equal(g['v_metas'], ref['v_metas'])
    errs = dict(states=helpers.relerr(g['states'], ref['states']), actions=helpers.relerr(g['actions'], ref['actions']),
                next_states=helpers.relerr(g['next_states'], ref['next_states']),
                rewards=float(np.abs(g['rewards'] - ref['rewards']).max()))
    print('rollout E=%d T=%d hidden=%s: %d envs vs oracle, max rel err %s' % (E, T, hidden, n_check, errs))
    # chaotic dynamics amplify rounding inside an episode; per-row tolerance stays far below the 1e-4 north star
    assert errs['states'] < 1e-6 and errs['actions'] < 1e-6 and errs['next_states'] < 1e-6 and errs['rewards'] < 1e-7
    assert np.allclose(g['c_info'], ref['c_info'], rtol=1e-6, atol=1e-8)
    # size-independent properties over the WHOLE batch: last step of every environment masked, logger totals
    masks = out['masks'].view(E, T)
    assert float(masks[:, -1].abs().sum()) == 0.0
    lg = out['logger'].cpu().numpy()
    assert lg[0] == E * T and lg[1] == float((out['masks'] == 0).sum())
    assert abs(lg[3] - float(out['rewards'].sum())) < 1e-7 * max(1.0, lg[3])
    assert torch.isfinite(out['states']).all() and torch.isfinite(out['rewards']).all()
    model.close()


def test_rollout_headline_shape_vs_oracle():
    """BASELINE config 2: 4096 envs x 300 steps, 243 -> 300 -> 300 -> 52 policy"""
    _rollout_vs_oracle(4096, 300, (300, 300), 128, 64, seed=11)


def test_rollout_config3_shape_vs_oracle():
    """BASELINE config 3: 8192 envs (256 CTAs on 148 SMs: two waves), 2 x 512 policy (chunked layer-2/3 plan)"""
    _rollout_vs_oracle(8192, 60, (512, 512), 128, 48, seed=13)


def _synthetic_batch(N, D, A, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    f64 = dict(dtype=torch.float64, device='cuda')
    states = torch.randn(N, D, generator=g, **f64)
    actions = torch.randn(N, A, generator=g, **f64) * 0.3
    rewards = torch.rand(N, generator=g, **f64)
    masks = (torch.rand(N, generator=g, **f64) > 0.1).to(torch.float64)
    masks[-1] = 0
    exps = (torch.rand(N, generator=g, **f64) > 0.05).to(torch.float64)
    return states, actions, rewards, masks, exps


def _agent(gemm, D, H, A, seed, oz_slices=None, epochs=10):
    from egopose_b200.agent import AgentPPO
    from egopose_b200.nets import MLP, PolicyGaussian, Value
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(seed)
    pol = PolicyGaussian(MLP(D, H, 'relu'), A, log_std=-2.3, fix_std=True).cuda()
    val = Value(MLP(D, H, 'relu')).cuda()
    opt_p = torch.optim.Adam(pol.parameters(), lr=5e-5)
    opt_v = torch.optim.Adam(val.parameters(), lr=3e-4)
    return AgentPPO(env=None, dtype=torch.float64, device=torch.device('cuda'), policy_net=pol, value_net=val,
                    optimizer_policy=opt_p, optimizer_value=opt_v, opt_num_epochs=epochs, gamma=0.95, tau=0.95,
                    clip_epsilon=0.2, policy_grad_clip=[(list(pol.parameters()), 40)], gemm=gemm, oz_slices=oz_slices)


def _torch_reference_update(agent0, states, actions, rewards, masks, exps, epochs):
    """agents/agent_pg.py:40-57 + agent_ppo.py:16-65 restated with torch autograd ON THE GPU (test oracle at full scale)"""
    import copy
    from egopose_b200 import lib
    pol, val = copy.deepcopy(agent0.policy_net), copy.deepcopy(agent0.value_net)
    opt_p = torch.optim.Adam(pol.parameters(), lr=5e-5)
    opt_v = torch.optim.Adam(val.parameters(), lr=3e-4)
    with torch.no_grad():
        values = val(states).view(-1)
    adv, ret, stats = lib.gae(rewards, masks, values.contiguous(), 0.95, 0.95)      # checked against the reference elsewhere
    adv = lib.standardize_(adv.clone(), stats)
    with torch.no_grad():
        logp0 = pol.get_log_prob(states, actions)
    ind = exps.nonzero().squeeze(1)
    surr, vl = [], []
    for _ in range(epochs):
        loss_v = (val(states).view(-1) - ret).pow(2).mean()
        opt_v.zero_grad(); loss_v.backward(); opt_v.step()
        lp = pol.get_log_prob(states[ind], actions[ind])
        ratio = torch.exp(lp - logp0[ind])
        a = adv[ind].view(-1, 1)
        loss_p = -torch.min(ratio * a, torch.clamp(ratio, 0.8, 1.2) * a).mean()
        opt_p.zero_grad(); loss_p.backward()
        torch.nn.utils.clip_grad_norm_(pol.parameters(), 40)
        opt_p.step()
        surr.append(float(loss_p)); vl.append(float(loss_v))
    return np.array(surr), np.array(vl), pol, val


def test_update_full_scale_ozaki_vs_cublas_vs_torch():
    from egopose_b200.trajbatch import TrajBatch
    N, D, H, A, EPOCHS = 1228800, 243, (300, 300), 52, 10
    states, actions, rewards, masks, exps = _synthetic_batch(N, D, A, seed=5)
    res = {}
    for name, gemm, S in (('ozaki6', 'ozaki', 6), ('ozaki7', 'ozaki', 7), ('cublas', 'cublas', None)):
        ag = _agent(gemm, D, H, A, seed=9, oz_slices=S, epochs=EPOCHS)
        if name == 'ozaki6':
            ref_agent = _agent('cublas', D, H, A, seed=9, epochs=EPOCHS)
            t_surr, t_vl, t_pol, t_val = _torch_reference_update(ref_agent, states, actions, rewards, masks, exps, EPOCHS)
            del ref_agent
        batch = TrajBatch(dev=dict(states=states, actions=actions, rewards=rewards, masks=masks, exps=exps), horizon=None)
        ag.update_params(batch)
        ls = ag.losses()
        res[name] = (ls['surr_loss'], ls['value_loss'],
                     torch.cat([p.data.view(-1) for p in ag.policy_net.parameters()]).clone(),
                     torch.cat([p.data.view(-1) for p in ag.value_net.parameters()]).clone())
        del ag, batch
        torch.cuda.empty_cache()
    t_p = torch.cat([p.data.view(-1) for p in t_pol.parameters()])
    t_v = torch.cat([p.data.view(-1) for p in t_val.parameters()])
    rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))  # noqa: E731
    for name, (surr, vl, pp, vp) in res.items():
        e = dict(surr=rel(surr, t_surr), value=rel(vl, t_vl), policy_params=float((pp - t_p).abs().max() / t_p.abs().max()),
                 value_params=float((vp - t_v).abs().max() / t_v.abs().max()))
        print('full-scale update N=%d x %d epochs, %s vs torch autograd: %s' % (N, EPOCHS, name, e))
        # north star: 1e-5 on the PPO loss
        assert e['surr'] < 1e-5 and e['value'] < 1e-5, (name, e)
        assert e['policy_params'] < 1e-5 and e['value_params'] < 1e-5, (name, e)
    # the two back ends against each other (same kernels around the GEMMs)
    assert rel(res['ozaki6'][0], res['cublas'][0]) < 1e-6 and rel(res['ozaki7'][0], res['cublas'][0]) < 1e-6


def test_two_updates_reuse_rollout_buffers_ozaki_matches_cublas():
    """the input-slice cache of the tensor-core back end must not survive an update: the second update sees NEW states at the
    SAME addresses (sample() reuses its output buffers)"""
    from egopose_b200.trajbatch import TrajBatch
    N, D, H, A = 6000, 40, (48, 32), 7
    bufs = _synthetic_batch(N, D, A, seed=1)
    agents = {g: _agent(g, D, H, A, seed=4, epochs=2) for g in ('ozaki', 'cublas')}
    for it in range(3):
        new = _synthetic_batch(N, D, A, seed=10 + it)
        for dst, src in zip(bufs, new):
            dst.copy_(src)                               # same tensors, new contents (what sample() does with self._out)
        losses = {}
        for g, ag in agents.items():
            ag.update_params(TrajBatch(dev=dict(states=bufs[0], actions=bufs[1], rewards=bufs[2], masks=bufs[3], exps=bufs[4]),
                                       horizon=None))
            losses[g] = ag.losses()
        assert np.allclose(losses['ozaki']['surr_loss'], losses['cublas']['surr_loss'], rtol=1e-8, atol=1e-11), it
        assert np.allclose(losses['ozaki']['value_loss'], losses['cublas']['value_loss'], rtol=1e-8), it
    po = torch.cat([p.data.view(-1) for p in agents['ozaki'].policy_net.parameters()])
    pc = torch.cat([p.data.view(-1) for p in agents['cublas'].policy_net.parameters()])
    assert float((po - pc).abs().max()) < 1e-9


@pytest.mark.parametrize('stock', [False, True])
def test_flat_buffers_survive_to_cpu_round_trip(stock):
    """ego_mimic.py:134 wraps checkpointing in to_cpu(...).  The shim's to_cpu swaps .data in place (parameters keep their
    identity); a stock module.to(cpu) / .to(cuda) round trip REPLACES the Parameter objects on current PyTorch - the next
    update must find the new objects by name, re-alias them to the flat buffers and re-point the caller's optimizer."""
    from egopose_b200.torch_utils import to_cpu
    from egopose_b200.trajbatch import TrajBatch
    N, D, H, A = 3000, 24, (32, 16), 5
    bufs = _synthetic_batch(N, D, A, seed=2)
    a1, a2 = _agent('cublas', D, H, A, seed=6, epochs=1), _agent('cublas', D, H, A, seed=6, epochs=1)
    mk = lambda: TrajBatch(dev=dict(states=bufs[0], actions=bufs[1], rewards=bufs[2], masks=bufs[3], exps=bufs[4]), horizon=None)  # noqa: E731
    for it in range(3):
        a1.update_params(mk())
        a2.update_params(mk())
        if stock:
            for net in (a2.policy_net, a2.value_net):
                net.to('cpu')
            sd = {k: v.clone() for k, v in a2.policy_net.state_dict().items()}
            for net in (a2.policy_net, a2.value_net):
                net.to('cuda')
        else:
            with to_cpu(a2.policy_net, a2.value_net):
                sd = {k: v.clone() for k, v in a2.policy_net.state_dict().items()}      # what a checkpoint would pickle
        assert all(v.device.type == 'cpu' for v in sd.values())
    a1.update_params(mk())
    a2.update_params(mk())
    # a lost alias freezes the nets at the first checkpoint's weights (differences ~1e-4); column sums with atomics make
    # two healthy runs differ in the last bits only
    for p1, p2 in zip(list(a1.policy_net.parameters()) + list(a1.value_net.parameters()),
                      list(a2.policy_net.parameters()) + list(a2.value_net.parameters())):
        assert p2.is_cuda and torch.allclose(p1.data, p2.data, rtol=1e-9, atol=1e-12)
    assert int(a2.optimizer_policy.state[a2.policy_net.action_mean.weight]['step']) == 4


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_on_one_gpu_equal_single_rank():
    """G ranks on the global batch == 1 rank on the same batch.  Two processes share cuda:0, so this runs on the
    1-GPU test lease: the small collectives go through gloo (device tensors staged by torch.distributed), the per-epoch
    gradient sum through the peer-memory kernel (csrc/p2p.cu: CUDA IPC works between processes on one device; the two
    kernels time-slice); the same path over NVLink at 2 - 8 GPUs is exercised by bench.py --gpus N / tools/check_multi_gpu.py.
    EGP_GRAD_EXCHANGE=nccl (torch.distributed all_reduce instead of the kernel) must give the same parameters."""
    script = os.path.join(ROOT, 'tools', 'check_multi_gpu.py')
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT=str(_free_port()), EGP_CHECK_BACKEND='gloo',
               EGP_CHECK_SAME_DEVICE='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', env['MASTER_PORT'], script, '--envs', '64', '--horizon', '12']
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'EQUIVALENT' in res.stdout, res.stdout[-3000:]
    assert 'peer memory (csrc/p2p.cu), barrier status 0' in res.stdout, res.stdout[-3000:]
    env['EGP_GRAD_EXCHANGE'] = 'nccl'
    env['MASTER_PORT'] = str(_free_port())
    cmd[cmd.index('--master-port') + 1] = env['MASTER_PORT']
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'EQUIVALENT' in res.stdout and 'torch.distributed all_reduce' in res.stdout, res.stdout[-3000:]


def test_streamed_host_batch_updates_like_the_device_batch():
    """sample(to_host=True) downloads on a side stream with one event per field (TrajBatch.to_host(stream=...)): the update
    from the host-only batch waits for the downloads ON THE DEVICE and re-uploads field by field; a numpy reader waits for
    the array it touches.  Same parameters as the update from the device batch; arrays identical to the device tensors."""
    from egopose_b200.trajbatch import TrajBatch
    N, D, H, A = 300000, 40, (48, 32), 7
    bufs = _synthetic_batch(N, D, A, seed=3)
    mk = lambda: TrajBatch(dev=dict(states=bufs[0], actions=bufs[1], rewards=bufs[2], masks=bufs[3], exps=bufs[4]), horizon=None)  # noqa: E731
    a_dev, a_host = _agent('cublas', D, H, A, seed=9, epochs=2), _agent('cublas', D, H, A, seed=9, epochs=2)
    side, pool = torch.cuda.Stream(), {}
    for it in range(2):
        a_dev.update_params(mk())
        b = mk().to_host(pool, stream=side)
        assert b.host_event('states') is not None            # nothing was synchronised
        hb = b.host_only()
        assert not hb.dev and hb.pinned('states') is not None
        a_host.update_params(hb)
        hb.wait_host()
        assert b.host_event('states') is None and np.array_equal(b.states, bufs[0].cpu().numpy())
        assert b.masks.dtype == np.int64 and np.array_equal(b.masks, bufs[3].cpu().numpy().astype(np.int64))
        for dst, src in zip(bufs, _synthetic_batch(N, D, A, seed=20 + it)):
            side.synchronize()
            dst.copy_(src)
    for p1, p2 in zip(list(a_dev.policy_net.parameters()) + list(a_dev.value_net.parameters()),
                      list(a_host.policy_net.parameters()) + list(a_host.value_net.parameters())):
        assert torch.allclose(p1.data, p2.data, rtol=1e-9, atol=1e-12)
    # a reader that touches one array waits for that array only
    b = mk().to_host(pool, stream=side)
    r = b.rewards
    assert np.array_equal(r, bufs[2].cpu().numpy()) and b.host_event('rewards') is None
    b.wait_host()
