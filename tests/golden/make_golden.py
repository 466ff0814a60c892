#!/usr/bin/env python
"""Generate the committed golden fixtures by executing the UNMODIFIED reference sources.

Runs only in the build container (needs /root/reference):   python tests/golden/make_golden.py

  ppo_small.npz     core/common.py estimate_advantages, core/policy_gaussian.py, core/critic.py,
                    agents/agent_ppo.py update_policy (3 epochs, torch Adam, grad-norm clip) on a seeded
                    synthetic trajbatch with small layer sizes (fixture size), float64
  math_helpers.npz  utils/math.py + utils/transformation.py helpers on random inputs
  zfilter.npz       utils/zfilter.py sequential filtering
  env_traj.npz      ego_pose/envs/humanoid_v1.py + ego_pose/core/reward_function.py stepped on the
                    restated physics through oracle/mujoco_shim.py (pins env logic, NOT MuJoCo itself),
                    and the gen_expert.py feature pipeline re-driven with the reference's own helpers
  expert_file.npz   ego_pose/data_process/gen_expert.py:28-83 get_expert, all 13 keys + scalars, with an lb:ub cut
  eval_forecast.npz ego_pose/ego_forecast_eval.py:95-180 windows with and without --gt-init (reference env + sync_traj)
  pose_metrics.npz  ego_pose/utils/metrics.py + the aggregation of ego_pose/eval_pose.py:31-66
  eval_traj.npz     ego_pose/ego_mimic_eval.py:93-177 evaluation roll-out ('naivefs' fail-safe) re-driven with the
                    reference's own env / align_human_state / PolicyGaussian / ZFilter on the restated physics
"""
import os
import pickle
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import cphys, mujoco_shim, refimport  # noqa: E402

refimport.install()
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)


def gen_ppo():
    from agents.agent_ppo import AgentPPO
    from core.common import estimate_advantages
    from core.critic import Value
    from core.policy_gaussian import PolicyGaussian
    from models.mlp import MLP

    torch.manual_seed(1)
    rng = np.random.RandomState(1)
    N, D, A, H = 640, 24, 6, (32, 16)
    policy = PolicyGaussian(MLP(D, H, 'relu'), A, log_std=-2.3, fix_std=True)
    value = Value(MLP(D, H, 'relu'))
    states = rng.randn(N, D)
    masks = np.ones(N)
    ends = np.sort(rng.choice(N - 1, size=24, replace=False))
    masks[ends] = 0
    masks[-1] = 0
    rewards = rng.rand(N)
    exps = (rng.rand(N) > 0.1).astype(np.float64)
    with torch.no_grad():
        mean = policy(torch.from_numpy(states)).loc.numpy()
    actions = mean + np.exp(-2.3) * rng.randn(N, A) * 1.5
    out = dict(states=states, actions=actions, masks=masks, rewards=rewards, exps=exps)
    for k, v in policy.state_dict().items():
        out['p0.' + k] = v.numpy().copy()
    for k, v in value.state_dict().items():
        out['v0.' + k] = v.numpy().copy()
    st, ac = torch.from_numpy(states), torch.from_numpy(actions)
    with torch.no_grad():
        values = value(st)
        out['values0'] = values.numpy().copy()
    adv, ret = estimate_advantages(torch.from_numpy(rewards), torch.from_numpy(masks), values, 0.95, 0.95)
    out['advantages'], out['returns'] = adv.numpy().copy(), ret.numpy().copy()
    with torch.no_grad():
        out['fixed_log_probs'] = policy.get_log_prob(st, ac).numpy().copy()

    opt_p = torch.optim.Adam(policy.parameters(), lr=5e-3)
    opt_v = torch.optim.Adam(value.parameters(), lr=3e-3)
    pparams = list(policy.parameters())
    agent = AgentPPO(env=None, dtype=torch.float64, device=torch.device('cpu'), policy_net=policy, value_net=value,
                     optimizer_policy=opt_p, optimizer_value=opt_v, opt_num_epochs=3, gamma=0.95, tau=0.95,
                     clip_epsilon=0.2, policy_grad_clip=[(pparams, 0.05)], use_mini_batch=False)
    ex = torch.from_numpy(exps)
    surr, vloss, gnorm = [], [], []
    # wrap to record per-epoch losses (calls the reference's own ppo_loss / update_value unchanged)
    orig_loss = agent.ppo_loss

    def rec_loss(*a):
        loss = orig_loss(*a)
        surr.append(loss.item())
        return loss
    agent.ppo_loss = rec_loss
    orig_clip = agent.clip_policy_grad

    def rec_clip():
        gnorm.append(float(torch.sqrt(sum((p.grad ** 2).sum() for p in pparams if p.grad is not None))))
        orig_clip()
    agent.clip_policy_grad = rec_clip
    orig_uv = agent.update_value

    def rec_uv(s_, r_):
        with torch.no_grad():
            vloss.append(float((value(s_) - r_).pow(2).mean()))
        orig_uv(s_, r_)
    agent.update_value = rec_uv
    agent.update_policy(st, ac, ret, adv, ex)
    out['surr_loss'], out['value_loss'], out['grad_norm'] = np.array(surr), np.array(vloss), np.array(gnorm)
    for k, v in policy.state_dict().items():
        out['p3.' + k] = v.numpy().copy()
    for k, v in value.state_dict().items():
        out['v3.' + k] = v.numpy().copy()
    out['hyper'] = np.array([0.95, 0.95, 0.2, 5e-3, 3e-3, 0.05])   # gamma tau clip lr_p lr_v max_norm
    np.savez_compressed(os.path.join(OUT, 'ppo_small.npz'), **out)
    print('ppo_small: surr', surr, 'vloss', vloss, 'gnorm', gnorm)


def gen_ppo_minibatch():
    """agents/agent_ppo.py:24-43 (use_mini_batch=True) on the ppo_small batch: 2 epochs x ceil(640/200) steps"""
    from agents.agent_ppo import AgentPPO
    from core.critic import Value
    from core.policy_gaussian import PolicyGaussian
    from models.mlp import MLP
    g = np.load(os.path.join(OUT, 'ppo_small.npz'))
    torch.manual_seed(1)
    D, A, H = 24, 6, (32, 16)
    policy = PolicyGaussian(MLP(D, H, 'relu'), A, log_std=-2.3, fix_std=True)
    value = Value(MLP(D, H, 'relu'))
    policy.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('p0.')})
    value.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('v0.')})
    opt_p = torch.optim.Adam(policy.parameters(), lr=5e-3)
    opt_v = torch.optim.Adam(value.parameters(), lr=3e-3)
    pparams = list(policy.parameters())
    agent = AgentPPO(env=None, dtype=torch.float64, device=torch.device('cpu'), policy_net=policy, value_net=value,
                     optimizer_policy=opt_p, optimizer_value=opt_v, opt_num_epochs=2, gamma=0.95, tau=0.95,
                     clip_epsilon=0.2, policy_grad_clip=[(pparams, 0.05)], use_mini_batch=True, opt_batch_size=200)
    surr = []
    orig_loss = agent.ppo_loss

    def rec_loss(*a):
        loss = orig_loss(*a)
        surr.append(loss.item())
        return loss
    agent.ppo_loss = rec_loss
    np.random.seed(123)
    agent.update_policy(torch.from_numpy(g['states']), torch.from_numpy(g['actions']), torch.from_numpy(g['returns']),
                        torch.from_numpy(g['advantages']), torch.from_numpy(g['exps']))
    out = {'surr_loss': np.array(surr), 'seed': np.array(123), 'opt_batch_size': np.array(200), 'epochs': np.array(2)}
    for k, v in policy.state_dict().items():
        out['p.' + k] = v.numpy().copy()
    for k, v in value.state_dict().items():
        out['v.' + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(OUT, 'ppo_minibatch.npz'), **out)
    print('ppo_minibatch: surr', np.round(surr, 5))


def gen_vsnet():
    """models/video_state_net.py + models/rnn.py (BiLSTM context) and ego_pose/core/agent_ego.py update_params with
    real vs-nets on a small synthetic batch (3 takes, margin 2, feature dim 8, v_hdim 6)."""
    import types
    from core.critic import Value
    from core.policy_gaussian import PolicyGaussian
    from ego_pose.core.agent_ego import AgentEgo
    from models.mlp import MLP
    from models.video_state_net import VideoStateNet
    torch.manual_seed(5)
    rng = np.random.RandomState(5)
    F, VH, M, S, A, T = 8, 6, 2, 10, 4, 7
    cnn_feat = [rng.randn(30, F) for _ in range(3)]
    pvs, vvs = VideoStateNet(F, VH, M), VideoStateNet(F, VH, M)
    pol = PolicyGaussian(MLP(S + VH, (16, 12), 'relu'), A, log_std=-1.0, fix_std=True)
    val = Value(MLP(S + VH, (16, 12), 'relu'))
    out = {'cnn_feat': np.stack(cnn_feat), 'dims': np.array([F, VH, M, S, A, T])}
    for name, net in (('pvs0', pvs), ('vvs0', vvs), ('p0', pol), ('v0', val)):
        for k, v in net.state_dict().items():
            out['%s.%s' % (name, k)] = v.numpy().copy()
    # test mode: one episode window -> v_out [T, VH]  (video_state_net.py:36-39)
    pvs.set_mode('test')
    with torch.no_grad():
        pvs.initialize(torch.from_numpy(cnn_feat[1][5 - M: 5 + T + M]))
        out['test_v_out'] = pvs.v_out.numpy().copy()
    # batch: episodes of ragged length (<= T), masks 0 at episode ends, v_metas (take, start)
    lens = [7, 3, 5, 7, 2, 6]
    metas = [(0, 4), (1, 9), (2, 3), (1, 5), (0, 12), (2, 8)]
    N = sum(lens)
    masks = np.ones(N)
    v_metas = np.zeros((N, 2), dtype=np.int64)
    i = 0
    for ln, mt in zip(lens, metas):
        v_metas[i:i + ln] = mt
        masks[i + ln - 1] = 0
        i += ln
    batch = types.SimpleNamespace(states=rng.randn(N, S), actions=rng.randn(N, A) * 0.5, rewards=rng.rand(N), masks=masks,
                                  exps=(rng.rand(N) > 0.15).astype(np.float64), v_metas=v_metas)
    for k in ('states', 'actions', 'rewards', 'masks', 'exps', 'v_metas'):
        out['batch.' + k] = getattr(batch, k)
    pparams = list(pol.parameters()) + list(pvs.parameters())
    vparams = list(val.parameters()) + list(vvs.parameters())
    opt_p = torch.optim.Adam(pparams, lr=3e-3)
    opt_v = torch.optim.Adam(vparams, lr=2e-3)
    env = types.SimpleNamespace(cnn_feat=cnn_feat)
    agent = AgentEgo(env=env, dtype=torch.float64, device=torch.device('cpu'), running_state=None, custom_reward=None,
                     policy_net=pol, policy_vs_net=pvs, value_net=val, value_vs_net=vvs, optimizer_policy=opt_p,
                     optimizer_value=opt_v, opt_num_epochs=3, gamma=0.95, tau=0.95, clip_epsilon=0.2,
                     policy_grad_clip=[(pparams, 0.5)])
    surr = []
    orig_loss = agent.ppo_loss

    def rec_loss(*a):
        loss = orig_loss(*a)
        surr.append(loss.item())
        return loss
    agent.ppo_loss = rec_loss
    agent.update_params(batch)
    out['surr_loss'] = np.array(surr)
    for name, net in (('pvs3', pvs), ('vvs3', vvs), ('p3', pol), ('v3', val)):
        for k, v in net.state_dict().items():
            out['%s.%s' % (name, k)] = v.numpy().copy()
    out['hyper'] = np.array([0.95, 0.95, 0.2, 3e-3, 2e-3, 0.5])
    np.savez_compressed(os.path.join(OUT, 'vsnet_small.npz'), **out)
    print('vsnet_small: surr', np.round(surr, 6), 'test_v_out', out['test_v_out'].shape)


def gen_fcnet():
    """models/video_forecast_net.py + models/rnn.py (causal LSTM v_out + per-step state LSTM) in test mode, and
    ego_pose/core/agent_ego.py update_params with real forecast vs-nets (train mode, padded unroll) on a small
    synthetic batch (3 takes, margin 4, feature dim 8, v_hdim 6, s_hdim 10)."""
    import types
    from core.critic import Value
    from core.policy_gaussian import PolicyGaussian
    from ego_pose.core.agent_ego import AgentEgo
    from models.mlp import MLP
    from models.video_forecast_net import VideoForecastNet
    torch.manual_seed(11)
    rng = np.random.RandomState(11)
    F, VH, M, S, A, T, SH = 8, 6, 4, 10, 4, 7, 10
    cnn_feat = [rng.randn(34, F) for _ in range(3)]
    pvs = VideoForecastNet(F, S, VH, M, 'lstm', None, SH, 'lstm', False)
    vvs = VideoForecastNet(F, S, VH, M, 'lstm', None, SH, 'lstm', False)
    pol = PolicyGaussian(MLP(pvs.out_dim, (16, 12), 'relu'), A, log_std=-1.0, fix_std=True)
    val = Value(MLP(vvs.out_dim, (16, 12), 'relu'))
    out = {'cnn_feat': np.stack(cnn_feat), 'dims': np.array([F, VH, M, S, A, T, SH])}
    for name, net in (('pvs0', pvs), ('vvs0', vvs), ('p0', pol), ('v0', val)):
        for k, v in net.state_dict().items():
            out['%s.%s' % (name, k)] = v.numpy().copy()
    # test mode (agent_ego.py:21-22 + agents/agent.py:44): initialize on the episode window, then step 5 states
    pvs.set_mode('test')
    test_states = rng.randn(5, S)
    with torch.no_grad():
        pvs.initialize(torch.from_numpy(cnn_feat[1][9 - M: 9 + T + M]))
        out['test_v_out'] = pvs.v_out.numpy().copy()
        out['test_x'] = np.concatenate([pvs(torch.from_numpy(test_states[[i]])).numpy() for i in range(5)])
        out['test_mu'] = pol(torch.from_numpy(out['test_x'])).loc.numpy().copy()
    out['test_states'] = test_states
    lens = [7, 3, 5, 7, 2, 6, 1]
    metas = [(0, 4), (1, 9), (2, 5), (1, 6), (0, 12), (2, 8), (1, 20)]
    N = sum(lens)
    masks = np.ones(N)
    v_metas = np.zeros((N, 2), dtype=np.int64)
    i = 0
    for ln, mt in zip(lens, metas):
        v_metas[i:i + ln] = mt
        masks[i + ln - 1] = 0
        i += ln
    batch = types.SimpleNamespace(states=rng.randn(N, S), actions=rng.randn(N, A) * 0.5, rewards=rng.rand(N), masks=masks,
                                  exps=(rng.rand(N) > 0.15).astype(np.float64), v_metas=v_metas)
    for k in ('states', 'actions', 'rewards', 'masks', 'exps', 'v_metas'):
        out['batch.' + k] = getattr(batch, k)
    # train-mode forward of the policy context on the batch (video_forecast_net.py:94-109)
    pvs.set_mode('train')
    pvs.initialize((torch.from_numpy(masks), cnn_feat, v_metas))
    with torch.no_grad():
        out['train_x'] = pvs(torch.from_numpy(batch.states)).numpy().copy()
    pparams = list(pol.parameters()) + list(pvs.parameters())
    vparams = list(val.parameters()) + list(vvs.parameters())
    opt_p = torch.optim.Adam(pparams, lr=3e-3)
    opt_v = torch.optim.Adam(vparams, lr=2e-3)
    env = types.SimpleNamespace(cnn_feat=cnn_feat)
    agent = AgentEgo(env=env, dtype=torch.float64, device=torch.device('cpu'), running_state=None, custom_reward=None,
                     policy_net=pol, policy_vs_net=pvs, value_net=val, value_vs_net=vvs, optimizer_policy=opt_p,
                     optimizer_value=opt_v, opt_num_epochs=3, gamma=0.95, tau=0.95, clip_epsilon=0.2,
                     policy_grad_clip=[(pparams, 0.5)])
    surr = []
    orig_loss = agent.ppo_loss

    def rec_loss(*a):
        loss = orig_loss(*a)
        surr.append(loss.item())
        return loss
    agent.ppo_loss = rec_loss
    agent.update_params(batch)
    out['surr_loss'] = np.array(surr)
    for name, net in (('pvs3', pvs), ('vvs3', vvs), ('p3', pol), ('v3', val)):
        for k, v in net.state_dict().items():
            out['%s.%s' % (name, k)] = v.numpy().copy()
    out['hyper'] = np.array([0.95, 0.95, 0.2, 3e-3, 2e-3, 0.5])
    np.savez_compressed(os.path.join(OUT, 'fcnet_small.npz'), **out)
    print('fcnet_small: surr', np.round(surr, 6), 'test_x', out['test_x'].shape)


def gen_math():
    from utils.math import (de_heading, get_angvel_fd, get_heading_q, get_qvel_fd, multi_quat_diff, multi_quat_norm,
                            transform_vec)
    from utils.transformation import (quaternion_from_euler, quaternion_inverse, quaternion_matrix,
                                      quaternion_multiply, rotation_from_quaternion)
    rng = np.random.RandomState(7)
    n = 16
    q = rng.randn(n, 4)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q2 = rng.randn(n, 4)
    q2 /= np.linalg.norm(q2, axis=1, keepdims=True)
    v = rng.randn(n, 3)
    eul = rng.uniform(-2, 2, size=(n, 3))
    qa = np.concatenate([rng.randn(n, 3), q, rng.randn(n, 52)], axis=1)
    dq = rng.randn(n, 3) * 0.05
    qb = qa.copy()
    qb[:, :3] += rng.randn(n, 3) * 0.02
    qb[:, 7:] += rng.randn(n, 52) * 0.03
    for i in range(n):   # small relative rotation
        ang = np.linalg.norm(dq[i])
        dqq = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * dq[i] / ang])
        qb[i, 3:7] = quaternion_multiply(dqq, qa[i, 3:7])
    qb[0, 3:7] = qa[0, 3:7]         # identical rotation -> 1 - w < 1e-8 branch
    bq0 = rng.randn(n, 21, 4)
    bq0 /= np.linalg.norm(bq0, axis=2, keepdims=True)
    bq1 = bq0 + 0.05 * rng.randn(n, 21, 4)
    bq1 /= np.linalg.norm(bq1, axis=2, keepdims=True)
    out = dict(q=q, q2=q2, v=v, eul=eul, qa=qa, qb=qb, bq0=bq0.reshape(n, 84), bq1=bq1.reshape(n, 84))
    out['mul'] = np.stack([quaternion_multiply(q[i], q2[i]) for i in range(n)])
    out['inv'] = np.stack([quaternion_inverse(q[i] * (1 + 0.1 * i)) for i in range(n)])
    out['inv_in'] = np.stack([q[i] * (1 + 0.1 * i) for i in range(n)])
    out['from_euler'] = np.stack([quaternion_from_euler(*eul[i]) for i in range(n)])
    out['matrix'] = np.stack([quaternion_matrix(q[i] * (1 + 0.1 * i))[:3, :3] for i in range(n)])
    out['rot_from_quat'] = np.stack([rotation_from_quaternion(q[i]) for i in range(n)])
    out['heading_q'] = np.stack([get_heading_q(q[i]) for i in range(n)])
    out['de_heading'] = np.stack([de_heading(q[i]) for i in range(n)])
    out['tv_root'] = np.stack([transform_vec(v[i], q[i], 'root') for i in range(n)])
    out['tv_heading'] = np.stack([transform_vec(v[i], q[i], 'heading') for i in range(n)])
    out['qvel_fd_none'] = np.stack([get_qvel_fd(qa[i], qb[i], 1 / 30.0) for i in range(n)])
    out['qvel_fd_heading'] = np.stack([get_qvel_fd(qa[i], qb[i], 1 / 30.0, 'heading') for i in range(n)])
    out['angvel_fd'] = np.stack([get_angvel_fd(out['bq0'][i], out['bq1'][i], 1 / 30.0) for i in range(n)])
    diff = np.stack([multi_quat_diff(out['bq1'][i], out['bq0'][i]) for i in range(n)])
    out['quat_diff'] = diff
    out['quat_norm'] = np.stack([multi_quat_norm(diff[i]) for i in range(n)])
    np.savez_compressed(os.path.join(OUT, 'math_helpers.npz'), **out)
    print('math_helpers ok')


def gen_zfilter():
    from utils.zfilter import ZFilter
    rng = np.random.RandomState(3)
    xs = rng.randn(40, 9) * rng.uniform(0.1, 20, size=9) + rng.randn(9)
    zf = ZFilter((9,), clip=5)
    ys = np.stack([zf(x) for x in xs])
    frozen = np.stack([zf(x, update=False) for x in xs[:5]])
    np.savez_compressed(os.path.join(OUT, 'zfilter.npz'), xs=xs, ys=ys, frozen=frozen, n=zf.rs.n, mean=zf.rs.mean,
                        std=zf.rs.std, S=zf.rs._S)
    print('zfilter ok')


def gen_env():
    orc = cphys.Oracle()
    mujoco_shim.install(orc)
    work = tempfile.mkdtemp(prefix='egopose_golden_')
    os.symlink(os.path.join(refimport.REF, 'config'), os.path.join(work, 'config'))
    os.symlink(os.path.join(refimport.REF, 'assets'), os.path.join(work, 'assets'))
    os.makedirs(os.path.join(work, 'datasets', 'meta'))
    os.makedirs(os.path.join(work, 'datasets', 'features'))
    take_names = ['take_a', 'take_b']
    import yaml
    yaml.safe_dump({'train': take_names, 'test': take_names},
                   open(os.path.join(work, 'datasets', 'meta', 'meta_subject_03.yml'), 'w'))
    os.chdir(work)

    from ego_pose.core.reward_function import quat_space_reward_v3
    from ego_pose.envs.humanoid_v1 import HumanoidEnv
    from ego_pose.utils.egomimic_config import Config
    from utils.math import de_heading, get_angvel_fd, get_qvel_fd, transform_vec

    L, EPLEN = 48, 10
    takes = cphys.synthetic_takes(orc.md, 2, L, seed=5)
    cfg = Config('subject_03', create_dirs=False)
    cfg.env_episode_len = EPLEN
    env = HumanoidEnv(cfg)
    env.seed(1)

    # ---- gen_expert.py:28-83 re-driven with the reference's own env methods / helpers -----------
    def get_expert(expert_qpos):
        keys = ['qvel', 'rlinv', 'rlinv_local', 'rangv', 'rq_rmh', 'head_pos', 'ee_pos', 'bquat', 'bangvel']
        ex = {k: [] for k in keys}
        for i in range(expert_qpos.shape[0]):
            qpos = expert_qpos[i]
            env.data.qpos[:] = qpos
            env.sim.forward()
            ex['rq_rmh'].append(de_heading(qpos[3:7]))
            ex['ee_pos'].append(env.get_ee_pos(cfg.obs_coord))
            ex['bquat'].append(env.get_body_quat())
            ex['head_pos'].append(env.get_body_com('Head').copy())
            if i > 0:
                qvel = get_qvel_fd(expert_qpos[i - 1], qpos, env.dt)
                ex['qvel'].append(qvel)
                ex['rlinv'].append(qvel[:3].copy())
                ex['rlinv_local'].append(transform_vec(qvel[:3].copy(), qpos[3:7], cfg.obs_coord))
                ex['rangv'].append(qvel[3:6].copy())
        for k in ('qvel', 'rlinv', 'rlinv_local', 'rangv'):
            ex[k].insert(0, ex[k][0].copy())
        for i in range(1, expert_qpos.shape[0]):
            ex['bangvel'].append(get_angvel_fd(ex['bquat'][i - 1], ex['bquat'][i], env.dt))
        ex['bangvel'].insert(0, ex['bangvel'][0].copy())
        out = {k: np.vstack(v) for k, v in ex.items()}
        out['qpos'] = expert_qpos
        out['len'] = expert_qpos.shape[0]
        out['height_lb'] = expert_qpos[:, 2].min()
        out['head_height_lb'] = out['head_pos'][:, 2].min()
        return out

    expert_dict = {n: get_expert(q) for n, q in zip(take_names, takes)}
    cnn = {n: np.random.RandomState(11).randn(L, 128) for n in take_names}
    pickle.dump(expert_dict, open(cfg.expert_feat_file, 'wb'))
    pickle.dump((cnn, {}), open(cfg.cnn_feat_file, 'wb'))
    env.load_experts(take_names, cfg.expert_feat_file, cfg.cnn_feat_file)

    out = dict(takes_qpos=np.stack(takes), episode_len=EPLEN)
    for k in ('qvel', 'rlinv_local', 'rangv', 'rq_rmh', 'ee_pos', 'bquat', 'bangvel'):
        out['expert.' + k] = np.stack([expert_dict[n][k] for n in take_names])
    out['expert.head_height_lb'] = np.array([expert_dict[n]['head_height_lb'] for n in take_names])

    # ---- episodes --------------------------------------------------------------------------------
    rng = np.random.RandomState(21)
    episodes = [dict(take=0, start=12, head_lb=None, end_reward=0.0),       # natural fail rule
                dict(take=1, start=15, head_lb=-10.0, end_reward=1.7)]      # runs to the time limit
    for ei, ep in enumerate(episodes):
        env.set_fix_sampling(expert_ind=ep['take'], start_ind=ep['start'])
        env.set_fix_head_lb(ep['head_lb'])
        env.end_reward = ep['end_reward']
        obs0 = env.reset()
        rec = dict(obs=[obs0], qpos=[env.data.qpos.copy()], qvel=[env.data.qvel.copy()], reward=[], c_info=[], fail=[],
                   end=[], action=[], head_z=[], torque0=[], qM_diag=[], bias=[])
        for t in range(EPLEN + 3):
            action = 0.3 * rng.randn(env.model.nu)
            rec['torque0'].append(env.compute_torque(cfg.a_ref + action * cfg.a_scale))   # PD torque at sub-step 0
            obs, _r, done, info = env.step(action)
            rew, c_info = quat_space_reward_v3(env, None, action, info)
            rec['action'].append(action)
            rec['obs'].append(obs)
            rec['qpos'].append(env.data.qpos.copy())
            rec['qvel'].append(env.data.qvel.copy())
            rec['reward'].append(rew)
            rec['c_info'].append(c_info)
            rec['fail'].append(info['fail'])
            rec['end'].append(info['end'])
            rec['head_z'].append(env.get_body_com('Head')[2])
            rec['bias'].append(env.data.qfrc_bias.copy())
            if done:
                break
        for k, v in rec.items():
            out['ep%d.%s' % (ei, k)] = np.array(v)
        out['ep%d.meta' % ei] = np.array([ep['take'], ep['start'], np.nan if ep['head_lb'] is None else ep['head_lb'],
                                          ep['end_reward']])
        print('episode', ei, 'steps', len(rec['reward']), 'fail', rec['fail'][-1], 'end', rec['end'][-1],
              'rewards', np.round(rec['reward'], 4))
    np.savez_compressed(os.path.join(OUT, 'env_traj.npz'), **out)


def gen_eval():
    """ego_pose/ego_mimic_eval.py:93-177 (eval_expert) re-driven in the script's call order with the reference's own
    HumanoidEnv (set_fix_sampling / set_fix_head_lb / reset / step), align_human_state, PolicyGaussian + MLP and
    ZFilter(update=False) on the restated physics; 'naivefs' fail-safe; state_pred = perturbed expert observations
    (the state-regression net of the script is outside the hot path)."""
    orc = cphys.Oracle()
    mujoco_shim.install(orc)
    work = tempfile.mkdtemp(prefix='egopose_golden_eval_')
    os.symlink(os.path.join(refimport.REF, 'config'), os.path.join(work, 'config'))
    os.symlink(os.path.join(refimport.REF, 'assets'), os.path.join(work, 'assets'))
    os.makedirs(os.path.join(work, 'datasets', 'meta'))
    os.makedirs(os.path.join(work, 'datasets', 'features'))
    take_names = ['take_a', 'take_b']
    import yaml
    yaml.safe_dump({'train': take_names, 'test': take_names},
                   open(os.path.join(work, 'datasets', 'meta', 'meta_subject_03.yml'), 'w'))
    os.chdir(work)

    from core.policy_gaussian import PolicyGaussian
    from ego_pose.core.reward_function import quat_space_reward_v3
    from ego_pose.envs.humanoid_v1 import HumanoidEnv
    from ego_pose.utils.egomimic_config import Config
    from models.mlp import MLP
    from utils.tools import align_human_state
    from utils.zfilter import ZFilter

    FM, CTX = 5, 6
    lens = [40, 33]
    takes = [cphys.synthetic_takes(orc.md, 1, L, seed=31 + i)[0] for i, L in enumerate(lens)]
    cfg = Config('subject_03', create_dirs=False)
    cfg.fr_margin = FM
    env = HumanoidEnv(cfg)
    env.seed(1)
    # expert dict through the oracle's gen_expert restatement (itself pinned by env_traj.npz)
    X = cphys.X
    expert_dict = {}
    for n, q in zip(take_names, takes):
        rows, lb = orc.expert_features(q)
        ex = {'qpos': q, 'len': q.shape[0], 'height_lb': q[:, 2].min(), 'head_height_lb': lb}
        for key, col, w in (('qvel', 'QVEL', 58), ('rlinv_local', 'RLINV_LOCAL', 3), ('rangv', 'RANGV', 3), ('rq_rmh', 'RQ_RMH', 4),
                            ('ee_pos', 'EE_POS', 15), ('bquat', 'BQUAT', 84), ('bangvel', 'BANGVEL', 63)):
            ex[key] = rows[:, X[col]:X[col] + w].copy()
        expert_dict[n] = ex
    rng = np.random.RandomState(41)
    cnn = {n: rng.randn(L, CTX) for n, L in zip(take_names, lens)}
    pickle.dump(expert_dict, open(cfg.expert_feat_file, 'wb'))
    pickle.dump((cnn, {}), open(cfg.cnn_feat_file, 'wb'))
    env.load_experts(take_names, cfg.expert_feat_file, cfg.cnn_feat_file)

    state_dim, action_dim = env.observation_space.shape[0], env.action_space.shape[0]
    torch.manual_seed(7)
    policy_net = PolicyGaussian(MLP(state_dim + CTX, (32, 16), 'relu'), action_dim, log_std=-2.3, fix_std=True)
    running_state = ZFilter((state_dim,), clip=5)
    for _ in range(50):                                  # some statistics, then frozen
        running_state(0.5 * rng.randn(state_dim))
    out = dict(fr_margin=FM, lens=np.array(lens), zf_mean=running_state.rs.mean.copy(), zf_std=running_state.rs.std.copy())
    for k, v in policy_net.state_dict().items():
        out['policy.' + k] = v.numpy().copy()

    def reset_env_state(state, ref_qpos):                # ego_mimic_eval.py:93-99
        qpos = ref_qpos.copy()
        qpos[2:] = state[:qpos.size - 2]
        qvel = state[qpos.size - 2:].copy()
        align_human_state(qpos, qvel, ref_qpos)
        env.set_state(qpos, qvel)
        return env.get_obs()

    for ti, name in enumerate(take_names):
        L = lens[ti]
        test_len = L - 2 * FM
        # state_pred: the expert's own observation of every frame plus a perturbation
        sp = np.zeros((L, state_dim))
        for fr in range(L):
            env.set_state(expert_dict[name]['qpos'][fr].copy(), expert_dict[name]['qvel'][fr].copy())
            sp[fr] = env.get_obs()
        sp[:, 5:] += 0.02 * rng.randn(L, state_dim - 5)       # joints / velocities; root height + quaternion kept
        # fail line a little below the starting head height: a contact-less humanoid crosses it every few steps
        env.set_state(expert_dict[name]['qpos'][FM].copy(), expert_dict[name]['qvel'][FM].copy())
        head_lb = env.get_body_com('Head')[2] - 0.03
        env.set_fix_head_lb(head_lb)
        env.set_fix_sampling(ti, FM, test_len)
        state = env.reset()
        cnn_feat = env.get_episode_cnn_feat()
        assert cnn_feat.shape[0] == L
        state = reset_env_state(sp[FM], env.data.qpos)
        state = running_state(state, update=False)
        traj_pred, vel_pred, states, actions, rewards, num_reset = [], [], [], [], [], 0
        for t in range(test_len):
            traj_pred.append(env.data.qpos.copy())
            vel_pred.append(env.data.qvel.copy())
            x = torch.from_numpy(np.concatenate([cnn_feat[FM + t], state])).unsqueeze(0)
            with torch.no_grad():
                action = policy_net.select_action(x, mean_action=True)[0].numpy()
            next_state, _r, done, info = env.step(action)
            next_state = running_state(next_state, update=False)
            reward, _ = quat_space_reward_v3(env, state, action, info)
            states.append(state)
            actions.append(action)
            rewards.append(reward)
            if info['end']:
                break
            if info['fail']:                             # --fail-safe naivefs
                num_reset += 1
                state = reset_env_state(sp[FM + t + 1], env.data.qpos)
                state = running_state(state, update=False)
            else:
                state = next_state
        pre = 'take%d.' % ti
        out[pre + 'qpos'] = takes[ti]
        out[pre + 'cnn'] = cnn[name]
        out[pre + 'state_pred'] = sp
        out[pre + 'head_lb'] = head_lb
        out[pre + 'traj_pred'] = np.vstack(traj_pred)
        out[pre + 'vel_pred'] = np.vstack(vel_pred)
        out[pre + 'states'] = np.vstack(states)
        out[pre + 'actions'] = np.vstack(actions)
        out[pre + 'rewards'] = np.array(rewards)
        out[pre + 'num_reset'] = num_reset
        print('eval take', ti, 'steps', len(rewards), 'num_reset', num_reset)
    # ---- the script's default fail-safe 'valuefs' (:156-159,167): value net in the loop, ONE RunningStat over both takes
    from core.critic import Value
    from utils.zfilter import RunningStat
    torch.manual_seed(17)
    value_net = Value(MLP(state_dim + CTX, (32, 16), 'relu'))
    with torch.no_grad():                                # a head that reacts to the state (the default init is x0.1, bias 0)
        value_net.value_head.weight.mul_(60.0)
        value_net.value_head.bias.fill_(1.0)
    for k, v in value_net.state_dict().items():
        out['value.' + k] = v.numpy().copy()
    vcnn = {n: rng.randn(L, CTX) for n, L in zip(take_names, lens)}      # value_vs_net has its own context
    value_stat = RunningStat(1)
    env.set_fix_head_lb(None)
    margin = np.inf
    for ti, name in enumerate(take_names):
        L = lens[ti]
        test_len = L - 2 * FM
        sp = out['take%d.state_pred' % ti]
        env.set_fix_sampling(ti, FM, test_len)
        state = env.reset()
        state = running_state(reset_env_state(sp[FM], env.data.qpos), update=False)
        traj_pred, vel_pred, values, num_reset = [], [], [], 0
        for t in range(test_len):
            traj_pred.append(env.data.qpos.copy())
            vel_pred.append(env.data.qvel.copy())
            with torch.no_grad():
                value = value_net(torch.from_numpy(np.concatenate([vcnn[name][FM + t], state])).unsqueeze(0)).item()
                value_stat.push(np.array([value]))
                x = torch.from_numpy(np.concatenate([cnn[name][FM + t], state])).unsqueeze(0)
                action = policy_net.select_action(x, mean_action=True)[0].numpy()
            values.append(value)
            next_state, _r, done, info = env.step(action)
            next_state = running_state(next_state, update=False)
            if info['end']:
                break
            margin = min(margin, abs(value - 0.6 * value_stat.mean[0]))
            if value < 0.6 * value_stat.mean[0]:
                num_reset += 1
                state = running_state(reset_env_state(sp[FM + t + 1], env.data.qpos), update=False)
            else:
                state = next_state
        pre = 'vfs.take%d.' % ti
        out[pre + 'vcnn'] = vcnn[name]
        out[pre + 'traj_pred'] = np.vstack(traj_pred)
        out[pre + 'vel_pred'] = np.vstack(vel_pred)
        out[pre + 'values'] = np.array(values)
        out[pre + 'num_reset'] = num_reset
        print('valuefs take', ti, 'steps', len(values), 'num_reset', num_reset, 'min decision margin', margin)
    out['vfs.value_stat'] = np.array([value_stat.n, value_stat.mean[0]])
    np.savez_compressed(os.path.join(OUT, 'eval_traj.npz'), **out)



def gen_expert_file():
    """ego_pose/data_process/gen_expert.py:28-83 get_expert re-driven with the reference's own env methods for ALL keys
    of the expert dict (incl. the ones the rollout never reads: rlinv, com, head_pos, obs, ee_wpos) and the lb:ub cut"""
    orc = cphys.Oracle()
    mujoco_shim.install(orc)
    work = tempfile.mkdtemp(prefix='egopose_golden_gx_')
    os.symlink(os.path.join(refimport.REF, 'config'), os.path.join(work, 'config'))
    os.symlink(os.path.join(refimport.REF, 'assets'), os.path.join(work, 'assets'))
    os.makedirs(os.path.join(work, 'datasets', 'meta'))
    import yaml
    yaml.safe_dump({'train': ['t'], 'test': ['t']}, open(os.path.join(work, 'datasets', 'meta', 'meta_subject_03.yml'), 'w'))
    os.chdir(work)
    from ego_pose.envs.humanoid_v1 import HumanoidEnv
    from ego_pose.utils.egomimic_config import Config
    from utils.math import de_heading, get_angvel_fd, get_qvel_fd, transform_vec
    cfg = Config('subject_03', create_dirs=False)
    env = HumanoidEnv(cfg)
    rng = np.random.RandomState(61)
    expert_qpos = cphys.synthetic_takes(orc.md, 1, 20, seed=61)[0]
    expert_qpos[:, 32:35] = 0.3 * rng.randn(20, 3)          # noisy hand data that get_expert removes (:38-39)
    expert_qpos[:, 42:45] = 0.3 * rng.randn(20, 3)
    raw = expert_qpos.copy()
    lb, ub = 2, 17
    feat_keys = ['qvel', 'rlinv', 'rlinv_local', 'rangv', 'rq_rmh', 'com', 'head_pos', 'obs', 'ee_pos', 'ee_wpos', 'bquat', 'bangvel']
    expert = {k: [] for k in feat_keys}
    for i in range(expert_qpos.shape[0]):
        qpos = expert_qpos[i]
        qpos[slice(*env.body_qposaddr['LeftHand'])] = 0.0
        qpos[slice(*env.body_qposaddr['RightHand'])] = 0.0
        env.data.qpos[:] = qpos
        env.sim.forward()
        expert['rq_rmh'].append(de_heading(qpos[3:7]))
        expert['obs'].append(env.get_obs())
        expert['ee_pos'].append(env.get_ee_pos(cfg.obs_coord))
        expert['ee_wpos'].append(env.get_ee_pos(None))
        expert['bquat'].append(env.get_body_quat())
        expert['com'].append(env.get_com())
        expert['head_pos'].append(env.get_body_com('Head').copy())
        if i > 0:
            qvel = get_qvel_fd(expert_qpos[i - 1], qpos, env.dt)
            expert['qvel'].append(qvel)
            expert['rlinv'].append(qvel[:3].copy())
            expert['rlinv_local'].append(transform_vec(qvel[:3].copy(), qpos[3:7], cfg.obs_coord))
            expert['rangv'].append(qvel[3:6].copy())
    for k in ('qvel', 'rlinv', 'rlinv_local', 'rangv'):
        expert[k].insert(0, expert[k][0].copy())
    for i in range(1, expert_qpos.shape[0]):
        expert['bangvel'].append(get_angvel_fd(expert['bquat'][i - 1], expert['bquat'][i], env.dt))
    expert['bangvel'].insert(0, expert['bangvel'][0].copy())
    out = {'raw_qpos': raw, 'lb': lb, 'ub': ub, 'qpos': expert_qpos[lb:ub]}
    for k in feat_keys:
        out[k] = np.vstack(expert[k][lb:ub])
    out['len'] = out['qpos'].shape[0]
    out['height_lb'] = out['qpos'][:, 2].min()
    out['head_height_lb'] = out['head_pos'][:, 2].min()
    np.savez_compressed(os.path.join(OUT, 'expert_file.npz'), **out)
    print('expert_file ok', out['len'], out['head_height_lb'])



def gen_eval_forecast():
    """ego_pose/ego_forecast_eval.py:95-180 re-driven in the script's call order (reference HumanoidEnv, sync_traj,
    PolicyGaussian, ZFilter) with and without --gt-init; the ego-mimic result it starts from is a perturbed, rotated and
    shifted copy of the expert (the kind of output ego_mimic_eval.py writes)."""
    orc = cphys.Oracle()
    mujoco_shim.install(orc)
    work = tempfile.mkdtemp(prefix='egopose_golden_fc_')
    os.symlink(os.path.join(refimport.REF, 'config'), os.path.join(work, 'config'))
    os.symlink(os.path.join(refimport.REF, 'assets'), os.path.join(work, 'assets'))
    os.makedirs(os.path.join(work, 'datasets', 'meta'))
    os.makedirs(os.path.join(work, 'datasets', 'features'))
    import yaml
    yaml.safe_dump({'train': ['t'], 'test': ['t']}, open(os.path.join(work, 'datasets', 'meta', 'meta_subject_03.yml'), 'w'))
    os.chdir(work)
    from core.policy_gaussian import PolicyGaussian
    from ego_pose.envs.humanoid_v1 import HumanoidEnv
    from ego_pose.utils.egomimic_config import Config
    from ego_pose.utils.tools import sync_traj
    from models.mlp import MLP
    from utils.zfilter import ZFilter

    FM, EMO, T, CTX, L = 6, 4, 8, 5, 34
    q = cphys.synthetic_takes(orc.md, 1, L, seed=77)[0]
    cfg = Config('subject_03', create_dirs=False)
    cfg.fr_margin = FM
    env = HumanoidEnv(cfg)
    env.seed(1)
    X = cphys.X
    rows, lb = orc.expert_features(q)
    ex = {'qpos': q, 'len': L, 'height_lb': q[:, 2].min(), 'head_height_lb': lb}
    for key, col, w in (('qvel', 'QVEL', 58), ('rlinv_local', 'RLINV_LOCAL', 3), ('rangv', 'RANGV', 3), ('rq_rmh', 'RQ_RMH', 4),
                        ('ee_pos', 'EE_POS', 15), ('bquat', 'BQUAT', 84), ('bangvel', 'BANGVEL', 63)):
        ex[key] = rows[:, X[col]:X[col] + w].copy()
    rng = np.random.RandomState(78)
    cnn = {'t': rng.randn(L, CTX)}
    pickle.dump({'t': ex}, open(cfg.expert_feat_file, 'wb'))
    pickle.dump((cnn, {}), open(cfg.cnn_feat_file, 'wb'))
    env.load_experts(['t'], cfg.expert_feat_file, cfg.cnn_feat_file)
    env.set_fix_head_lb(-1e30)                           # the script does not stop on a fall (:171-176)
    # ego-mimic result: frames EMO .. L - EMO of the take, yawed by 0.7 rad, shifted, perturbed
    from utils.transformation import quaternion_about_axis, quaternion_multiply
    from utils.math import quat_mul_vec
    yaw = quaternion_about_axis(0.7, [0, 0, 1])
    em_traj, em_vel = q[EMO:L - EMO].copy(), ex['qvel'][EMO:L - EMO].copy()
    for i in range(em_traj.shape[0]):
        em_traj[i, :3] = quat_mul_vec(yaw, em_traj[i, :3]) + np.array([0.4, -0.3, 0.0])
        em_traj[i, 3:7] = quaternion_multiply(yaw, em_traj[i, 3:7])
        em_vel[i, :3] = quat_mul_vec(yaw, em_vel[i, :3])
    em_traj[:, 7:] += 0.02 * rng.randn(*em_traj[:, 7:].shape)
    em_vel[:, 6:] += 0.05 * rng.randn(*em_vel[:, 6:].shape)
    state_dim, action_dim = env.observation_space.shape[0], env.action_space.shape[0]
    torch.manual_seed(9)
    policy_net = PolicyGaussian(MLP(state_dim + CTX, (32, 16), 'relu'), action_dim, log_std=-2.3, fix_std=True)
    running_state = ZFilter((state_dim,), clip=5)
    for _ in range(50):
        running_state(0.5 * rng.randn(state_dim))
    out = dict(fr_margin=FM, em_offset=EMO, test_len=T, qpos=q, cnn=cnn['t'], em_traj=em_traj, em_vel=em_vel,
               zf_mean=running_state.rs.mean.copy(), zf_std=running_state.rs.std.copy())
    for k, v in policy_net.state_dict().items():
        out['policy.' + k] = v.numpy().copy()

    def eval_expert(start_ind, gt_init):                 # ego_forecast_eval.py:95-180
        traj_pred = []
        env.set_fix_sampling(0, start_ind, T)
        state = env.reset()
        miss_len = 0
        if not gt_init:
            state_pred = em_traj[max(0, start_ind - FM - EMO): start_ind + T - EMO]
            vel_pred = em_vel[max(0, start_ind - FM - EMO): start_ind + T - EMO]
            miss_len = FM + T - state_pred.shape[0]
            if start_ind - FM - EMO >= 0:
                ref_qpos = env.get_expert_attr('qpos', env.get_expert_index(-FM)).copy()
                state_pred, vel_pred = sync_traj(state_pred, vel_pred, ref_qpos)
            ind = FM - miss_len
            env.set_state(state_pred[ind].copy(), vel_pred[ind].copy())
            state = env.get_obs()
        state = running_state(state, update=False)
        for t in range(-FM, 0):
            epos = env.get_expert_attr('qpos', env.get_expert_index(t)).copy()
            traj_pred.append(epos.copy() if (gt_init or t + FM < miss_len) else state_pred[t + FM - miss_len].copy())
        for t in range(T):
            traj_pred.append(env.data.qpos.copy())
            x = torch.from_numpy(np.concatenate([cnn['t'][start_ind + t], state])).unsqueeze(0)
            with torch.no_grad():
                action = policy_net.select_action(x, mean_action=True)[0].numpy()
            next_state, _r, _done, _info = env.step(action)
            state = running_state(next_state, update=False)
        return np.vstack(traj_pred)

    starts = list(range(FM, L - T + 1, FM))
    out['starts'] = np.array(starts)
    out['traj_pred_gt'] = np.stack([eval_expert(s0, True) for s0 in starts])
    out['traj_pred_em'] = np.stack([eval_expert(s0, False) for s0 in starts])
    print('eval_forecast starts', starts, out['traj_pred_em'].shape)
    np.savez_compressed(os.path.join(OUT, 'eval_forecast.npz'), **out)



def gen_metrics():
    """ego_pose/utils/metrics.py (get_joint_angles / get_joint_vels / get_joint_accels / get_mean_dist / get_mean_abs) and the
    aggregation of ego_pose/eval_pose.py:31-66 on two synthetic (prediction, ground truth) trajectory pairs"""
    from ego_pose.utils.metrics import get_joint_accels, get_joint_angles, get_joint_vels, get_mean_abs, get_mean_dist
    orc = cphys.Oracle()
    rng = np.random.RandomState(91)
    dt = 1 / 30.0
    out, tot = {}, np.zeros(3)
    for i, L in enumerate((17, 12)):
        gt = cphys.synthetic_takes(orc.md, 1, L, seed=90 + i)[0]
        pred = gt.copy()
        pred[:, :3] += 0.05 * rng.randn(L, 3)
        q = pred[:, 3:7] + 0.1 * rng.randn(L, 4)
        pred[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
        pred[:, 7:] += 0.1 * rng.randn(L, gt.shape[1] - 7)
        pred[3] = pred[2]                                    # identical consecutive frames: the zero-rotation branch
        angs_gt, vels_gt = get_joint_angles(gt), get_joint_vels(gt, dt)
        angs, vels = get_joint_angles(pred), get_joint_vels(pred, dt)
        accels = get_joint_accels(vels, dt)
        m = np.array([get_mean_dist(angs, angs_gt), get_mean_dist(vels, vels_gt), get_mean_abs(accels)])
        tot += m
        out.update({'gt%d' % i: gt, 'pred%d' % i: pred, 'angs%d' % i: angs, 'vels%d' % i: vels, 'accels%d' % i: accels,
                    'metrics%d' % i: m})
    out['global'] = tot / 2
    np.savez_compressed(os.path.join(OUT, 'pose_metrics.npz'), **out)
    print('pose_metrics ok', out['global'])


if __name__ == '__main__':
    which = sys.argv[1:] or ['ppo', 'math', 'zfilter', 'env', 'vsnet', 'fcnet', 'eval', 'expert_file', 'eval_forecast', 'metrics']
    if 'ppo' in which:
        gen_ppo()
    if 'ppo_mb' in which or 'ppo' in which:
        gen_ppo_minibatch()
    if 'vsnet' in which:
        gen_vsnet()
    if 'fcnet' in which:
        gen_fcnet()
    if 'math' in which:
        gen_math()
    if 'zfilter' in which:
        gen_zfilter()
    if 'eval' in which:
        gen_eval()
    if 'expert_file' in which:
        gen_expert_file()
    if 'eval_forecast' in which:
        gen_eval_forecast()
    if 'metrics' in which:
        gen_metrics()
    if 'env' in which:
        gen_env()
