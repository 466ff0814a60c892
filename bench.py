#!/usr/bin/env python
"""Headline benchmark: env-steps/sec of one full PPO iteration (fused humanoid rollout + PPO update).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...        (one rank per GPU, NCCL; environments sharded, weak scaling)

A "step" = one PPO iteration on the BASELINE.json configs[1] workload: subject_03 egomimic humanoid
(nq 59 / nv 58 / nu 52), E = 4096 environments x T = 300 steps per GPU, policy / value MLP 243 -> 300 ->
300 -> 52 | 1 (relu), 10 full-batch PPO epochs, float64 like the reference; synthetic experts and CNN
features (the dataset is not redistributable), random-init weights.
`value`  = E*T*N / (device time of rollout + update), inputs resident in HBM.
`e2e`    = same metric through the reference-facing Python API with HOST trajbatches: agent.sample()
           returns numpy arrays (D2H inside the timed region), agent.update_params(host batch) uploads them.
`--impl reference` times the CPU oracle port (C physics/rollout on all host cores + torch-CPU float64 PPO
update, the reference's own libraries) on a bounded sample of the same workload; /root/reference is not
needed at run time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

WORKLOAD = 'subject_03 egomimic humanoid PPO iteration: 4096 envs x 300 steps per GPU, 2x300 MLP policy/value, 10 epochs'

# BASELINE.json configs (per-GPU shares; `--gpus N` weak-scales them).  2 is the default the driver measures; the others are
# run explicitly (`--config K`) and their lines kept under profiles/.
CONFIGS = {
    2: dict(task='egomimic', cfg='subject_03', envs=4096, horizon=300, hidden=[300, 300], n_takes=8, minibatch=0,
            workload=WORKLOAD),
    3: dict(task='egomimic', cfg='cross_01', envs=8192, horizon=300, hidden=[512, 512], n_takes=32, minibatch=0,
            workload='cross_01 egomimic humanoid PPO iteration: 8192 envs x 300 steps per GPU, 2x512 MLP policy/value, 10 epochs'),
    4: dict(task='egoforecast', cfg='subject_03', envs=4096, horizon=90, hidden=[300, 200], n_takes=8, minibatch=0,
            workload='subject_03 egoforecast humanoid PPO iteration: 4096 envs x 90 steps per GPU (16384 envs on 4 GPUs), '
                     'VideoForecastNet context (causal LSTM over 30 past CNN-feature frames, precomputed features) + state LSTM '
                     '115->128 stepped inside the rollout kernel, 2-layer [300, 200] MLP policy/value on the 256-wide input, 10 epochs'),
    5: dict(task='egomimic', cfg='cross_01', envs=8192, horizon=300, hidden=[512, 512], n_takes=32, minibatch=65536,
            workload='cross_01 egomimic humanoid PPO iteration: 8192 envs x 300 steps per GPU (65536 envs on 8 GPUs), 2x512 MLP, '
                     'mini-batch PPO (opt_batch_size 65536, agents/agent_ppo.py:24-43), 10 epochs'),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=sorted(CONFIGS), help='BASELINE.json config (2 = headline)')
    ap.add_argument('--envs', type=int, default=None)
    ap.add_argument('--horizon', type=int, default=None)
    ap.add_argument('--hidden', type=int, nargs=2, default=None)
    ap.add_argument('--epochs', type=int, default=10)
    ap.add_argument('--no-variants', action='store_true', help='skip the cuBLAS / 7-slice update timings')
    ap.add_argument('--physics', choices=['smooth', 'full'], default='smooth',
                    help='smooth = the north-star scope (no constraints); full = + the 52 joint ranges and floor contact '
                         '(MuJoCo soft constraints; runs on the one-warp rollout kernel)')
    ap.add_argument('--vsnet', action='store_true', help='egomimic with the learned VideoStateNet (BiLSTM 128 -> 2 x 64 over T + 2 m '
                    'CNN-feature frames) as policy / value context instead of the per-frame table; episodes run the full horizon')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-envs', type=int, default=0, help='envs in the CPU sample (0 = 2 per core)')
    ap.add_argument('--cpu-horizon', type=int, default=20)
    ap.add_argument('--cpu-update-rows', type=int, default=12288)
    args = ap.parse_args()
    c = CONFIGS[args.config]
    args.envs = args.envs or c['envs']
    args.horizon = args.horizon or c['horizon']
    args.hidden = args.hidden or list(c['hidden'])
    args.task, args.cfg_id, args.n_takes, args.minibatch, args.workload = c['task'], c['cfg'], c['n_takes'], c['minibatch'], c['workload']
    if args.vsnet:
        args.workload += ' [VideoStateNet context nets in sample and update (video_state_net.py), head-height fail rule off]'
    if (args.envs, args.horizon, args.hidden) != (c['envs'], c['horizon'], list(c['hidden'])):
        args.workload += ' [overridden: %d envs x %d steps, hidden %s]' % (args.envs, args.horizon, args.hidden)
    return args


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def tensor_peak():
    """dense bf16 TFLOP/s measured on this pool (burst, for a kernel timed alone); int8 tcgen05 runs at twice that rate"""
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)).get('bf16_tflops', 1590.0), 'measured (MEASURED_PEAKS.json)'
    return 1590.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def synthetic_problem(args):
    """seeded synthetic experts: smooth in-range trajectories, N(0,1) CNN features"""
    from egopose_b200.mjcf import load_builtin
    from egopose_b200.synthetic import synthetic_cnn_feat, synthetic_takes
    md = load_builtin()
    margin = 30 if args.task == 'egoforecast' else 10
    L = args.horizon + 2 * margin + 64
    return md, synthetic_takes(md, args.n_takes, L, seed=1), synthetic_cnn_feat(args.n_takes, L)


def build_agent(args, device, takes, cnn, gemm=None, oz_slices=None):
    import torch
    from egopose_b200.agent import AgentEgo
    from egopose_b200.config import Config
    from egopose_b200.env import HumanoidEnv
    from egopose_b200.nets import MLP, FrameContext, PolicyGaussian, Value, VideoForecastNet, VideoStateNet
    from egopose_b200.zfilter import ZFilter
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(1)
    cfg = Config(args.cfg_id, task=args.task)
    cfg.env_episode_len = args.horizon
    env = HumanoidEnv(cfg, device=device.index)
    if args.physics == 'full':
        env.kernel.set_joint_limits(True)
        env.kernel.set_contacts(True)
    env.seed(cfg.seed)
    env.set_expert_qpos(['take_%d' % i for i in range(len(takes))], takes, cnn)
    sd, ad = env.observation_space.shape[0], env.action_space.shape[0]
    if args.task == 'egoforecast':          # ego_forecast.py:53-58: forecast context nets (v_hdim 128, state LSTM 128)
        mk = lambda: VideoForecastNet(128, sd, 128, cfg.fr_margin, 'lstm', None, 128, 'lstm').to(device)  # noqa: E731
        pvs, vvs = mk(), mk()
        in_dim = pvs.out_dim
    elif getattr(args, 'vsnet', False):     # ego_mimic.py:52-53: BiLSTM video context, re-run on every forward of the update
        pvs = VideoStateNet(128, cfg.policy_v_hdim, cfg.fr_margin, 'lstm', None, False).to(device)
        vvs = VideoStateNet(128, cfg.value_v_hdim, cfg.fr_margin, 'lstm', None, False).to(device)
        in_dim = sd + cfg.policy_v_hdim
        env.set_fix_head_lb(-10.0)          # full-length episodes, as in a trained run (contact-less bodies fall after ~9 steps)
    else:
        pvs, vvs, in_dim = FrameContext(128), FrameContext(128), sd + 128
    policy = PolicyGaussian(MLP(in_dim, args.hidden, 'relu'), ad, log_std=cfg.log_std, fix_std=cfg.fix_std).to(device)
    value = Value(MLP(in_dim, args.hidden, 'relu')).to(device)
    pparams = list(policy.parameters()) + [p for p in pvs.parameters()]
    vparams = list(value.parameters()) + [p for p in vvs.parameters()]
    opt_p = torch.optim.Adam(pparams, lr=cfg.policy_lr)
    opt_v = torch.optim.Adam(vparams, lr=cfg.value_lr)
    agent = AgentEgo(env=env, dtype=torch.float64, device=device, running_state=ZFilter((sd,), clip=5),
                     custom_reward=None, num_threads=1, policy_net=policy, policy_vs_net=pvs,
                     value_net=value, value_vs_net=vvs, optimizer_policy=opt_p, optimizer_value=opt_v,
                     opt_num_epochs=args.epochs, gamma=cfg.gamma, tau=cfg.tau, clip_epsilon=cfg.clip_epsilon,
                     policy_grad_clip=[(pparams, 40)], num_envs=args.envs, horizon=args.horizon,
                     use_mini_batch=args.minibatch > 0, opt_batch_size=args.minibatch or 64, gemm=gemm, oz_slices=oz_slices)
    return agent, cfg


def cpu_reference(args, takes, cnn, hidden, threads):
    """CPU oracle port on a bounded sample -> (env-steps/s for a full iteration, detail dict)"""
    import torch
    from oracle import cphys, ppo as oppo
    import helpers
    cores = threads or os.cpu_count() or 1
    orc = cphys.Oracle(episode_len=args.horizon)
    ctx = np.concatenate(cnn)
    orc.make_expert(takes, ctx)
    S, nu = orc.S, orc.nu
    w = helpers.policy_weights(S + 128, hidden[0], hidden[1], nu, seed=1)
    E = args.cpu_envs or 2 * cores
    T = args.cpu_horizon
    rng = np.random.RandomState(5)
    L = takes[0].shape[0]
    rt = rng.randint(0, len(takes), size=(E, T))
    rs = rng.randint(10, L - args.horizon - 10, size=(E, T))
    eps = rng.randn(E * T, nu)
    pol = orc.make_policy(w['W1'], w['b1'], w['W2'], w['b2'], w['W3'], w['b3'], w['log_std'])
    t0 = time.perf_counter()
    out = orc.rollout(pol, E, T, rt, rs, eps, n_threads=cores)
    t_roll = time.perf_counter() - t0
    # update: reference path (torch CPU float64 autograd + torch.optim.Adam), epochs as configured
    n_up = min(args.cpu_update_rows, 1 << 20)
    reps = (n_up + E * T - 1) // (E * T)
    tile = lambda a: np.concatenate([a] * reps)[:n_up]  # noqa: E731
    states = np.concatenate([np.concatenate([ctx[:1].repeat(E * T, 0), out['states']], axis=1)] * reps)[:n_up]
    vw = helpers.policy_weights(S + 128, hidden[0], hidden[1], 1, seed=2)
    pdict = {'net.affine_layers.0.weight': w['W1'], 'net.affine_layers.0.bias': w['b1'], 'net.affine_layers.1.weight': w['W2'],
             'net.affine_layers.1.bias': w['b2'], 'action_mean.weight': w['W3'], 'action_mean.bias': w['b3'],
             'action_log_std': w['log_std']}
    vdict = {'net.affine_layers.0.weight': vw['W1'], 'net.affine_layers.0.bias': vw['b1'], 'net.affine_layers.1.weight': vw['W2'],
             'net.affine_layers.1.bias': vw['b2'], 'value_head.weight': vw['W3'], 'value_head.bias': vw['b3']}
    torch.set_num_threads(cores)
    t0 = time.perf_counter()
    with torch.no_grad():
        values = oppo.value_forward(torch.from_numpy(states), {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in vdict.items()}).numpy()
    masks = tile(out['masks'])
    masks[-1] = 0
    adv, ret = oppo.gae(tile(out['rewards']), masks, values, 0.95, 0.95)
    oppo.ppo_update(pdict, vdict, states, tile(out['actions']), ret, adv, tile(out['exps']), 0.2, 5e-5, 3e-4, 40.0,
                    epochs=args.epochs, threads=cores)
    t_up = time.perf_counter() - t0
    per_step = t_roll / (E * T) + t_up / n_up
    detail = dict(cores=cores, rollout_steps_per_s=E * T / t_roll, update_samples_per_s=n_up / t_up,
                  sample='rollout %d envs x %d steps (C oracle, %d threads) + GAE/%d-epoch PPO update on %d rows '
                         '(torch CPU f64, %d threads); iteration rate = 1/(t_roll/step + t_update/row)'
                         % (E, T, cores, args.epochs, n_up, cores))
    return 1.0 / per_step, detail


def cons_passes(lib):
    """average number of (backward, forward) passes of the active-set iteration per sub-step in the block-sweep kernel"""
    import ctypes
    out = (ctypes.c_int64 * 2)()
    if lib.load().egp_cons_passes(out, 0) != 0 or out[0] == 0:
        return None
    return out[1] / out[0]


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    _, takes, cnn = synthetic_problem(args)
    vals = []
    for _ in range(max(1, args.warmup > 0)):
        cpu_reference(args, takes, cnn, args.hidden, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, detail = cpu_reference(args, takes, cnn, args.hidden, 0)
        vals.append(v)
    ms = (time.perf_counter() - t0) / args.steps * 1e3
    v = float(np.mean(vals))
    line = {'impl': 'reference', 'metric': 'env-steps/sec (humanoid PPO rollout+update)', 'value': v, 'unit': 'env-steps/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': args.workload, 'baseline_config': args.config, 'envs_per_gpu': args.envs, 'horizon': args.horizon, 'hidden': args.hidden,
                       'epochs': args.epochs, 'parallelism': 'host threads (rank 0 only)'},
            'cpu_baseline': {'value': v, 'unit': 'env-steps/s', 'cores': detail['cores'], 'kind': 'port',
                             'sample': detail['sample'], 'rollout_steps_per_s': detail['rollout_steps_per_s'],
                             'update_samples_per_s': detail['update_samples_per_s']},
            'e2e': {'value': v, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from egopose_b200 import lib
    from egopose_b200.trajbatch import TrajBatchEgo
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    md, takes, cnn = synthetic_problem(args)
    agent, cfg = build_agent(args, device, takes, cnn)
    agent.env._seed = cfg.seed + 1000 * rank             # decorrelate ranks (agents/agent.py:30-32 analogue)
    E, T = args.envs, args.horizon
    N = E * T

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def iteration(host):
        batch, log = agent.sample(N, to_host=host)
        agent.env.end_reward = log.avg_c_reward * cfg.gamma / (1 - cfg.gamma)       # ego_mimic.py:112
        if host:        # reference-format call: a batch object that only carries host (numpy) arrays
            batch = batch.host_only()
        agent.update_params(batch)
        if host:
            batch.wait_host()           # the step ends when every downloaded array has landed (next_states comes down last)
        return log

    for _ in range(args.warmup):
        iteration(False)
    clk = ClockSampler(local)
    barrier()
    clk.start()
    l0 = lib.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    t_roll = t_upd = 0.0
    for _ in range(args.steps):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        batch, log = agent.sample(N, to_host=False)
        agent.env.end_reward = log.avg_c_reward * cfg.gamma / (1 - cfg.gamma)
        e1.record()
        agent.update_params(batch)
        e2.record()
        e2.synchronize()
        t_roll += e0.elapsed_time(e1)
        t_upd += e1.elapsed_time(e2)
    ev[1].record()
    barrier()
    launches = lib.launches - l0
    clocks = clk.stop()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    tm = torch.tensor([ms, t_roll / args.steps, t_upd / args.steps], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, ms_roll, ms_upd = tm.tolist()
    value = N * world / (ms / 1e3)

    # ---- kernel rooflines: rollout kernel alone, GAE kernel alone (CUDA events on the launching stream)
    hbm, peak_src = peaks()
    w = agent._policy_weights()
    zm, zs, clip = agent._zf()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    times = []
    for i in range(3):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rctx, rwin = agent._rollout_context()
        agent.env.kernel.rollout(w, E, T, T, cfg.fr_margin, zf_mean=zm, zf_std=zs, zf_clip=clip, seed=7, iteration=900 + i,
                                 want_next=False, want_raw=True, out=agent._out, ctx=rctx, win_off=rwin, **agent._rollout_extra())
        b.record()
        b.synchronize()
        times.append(a.elapsed_time(b))
    roll_ms = float(np.median(times))
    wbytes = 8                                            # float64
    roll_bytes = (128 + 166 + 115 + 52 + 115 + 5 + 5) * wbytes * N      # ctx + expert row + states/actions/raw_obs/c_info/scalars
    rewards, masks = agent._out['rewards'], agent._out['masks']
    # GAE: 12 distinct input sets (12 x 29 MB in + 12 x 20 MB out > L2) processed back to back so that the timed
    # region is device-bound (one call is ~3 launches of a few microseconds each)
    sets = [(rewards + 0.01 * i, masks.clone(), torch.randn(N, dtype=torch.float64, device=device)) for i in range(12)]
    work = torch.empty(lib.load().egp_gae_work_bytes(N), dtype=torch.uint8, device=device)
    outs = [(torch.empty(N, dtype=torch.float64, device=device), torch.empty(N, dtype=torch.float64, device=device),
             torch.empty(3, dtype=torch.float64, device=device)) for _ in sets]
    gt = []
    for rep in range(4):
        flush.fill_(rep)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for (r_i, m_i, v_i), o_i in zip(sets, outs):
            lib.gae(r_i, m_i, v_i, cfg.gamma, cfg.tau, work=work, out=o_i)
        b.record()
        b.synchronize()
        gt.append(a.elapsed_time(b) / len(sets))
    del sets, outs
    gae_ms = float(np.median(gt))
    # the same scan at BASELINE config 5's batch (65536 envs x 300 steps = 19.66 M samples, 786 MB algorithmic): at that
    # size the three launches are bandwidth- instead of latency-dominated
    gae_big = None
    try:
        NB = 65536 * 300
        rb = (torch.rand(NB, dtype=torch.float64, device=device), (torch.rand(NB, dtype=torch.float64, device=device) > 0.11).double(),
              torch.randn(NB, dtype=torch.float64, device=device))
        ob = (torch.empty(NB, dtype=torch.float64, device=device), torch.empty(NB, dtype=torch.float64, device=device),
              torch.empty(3, dtype=torch.float64, device=device))
        wb = torch.empty(lib.load().egp_gae_work_bytes(NB), dtype=torch.uint8, device=device)
        bt = []
        for rep in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            lib.gae(rb[0], rb[1], rb[2], cfg.gamma, cfg.tau, work=wb, out=ob)
            b.record()
            b.synchronize()
            bt.append(a.elapsed_time(b))
        bms = float(np.median(bt[1:]))
        old_min = lib.gae_set_onepass_min(1 << 62)        # the small-batch (two-pass) kernels at this size, for comparison
        bt2 = []
        for rep in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            lib.gae(rb[0], rb[1], rb[2], cfg.gamma, cfg.tau, work=wb, out=ob)
            b.record()
            b.synchronize()
            bt2.append(a.elapsed_time(b))
        lib.gae_set_onepass_min(old_min)
        gae_big = {'bound': 'hbm', 'samples': NB, 'ms': bms, 'achieved': 5 * wbytes * NB / (bms / 1e3) / 1e9, 'peak': hbm,
                   'unit': 'GB/s', 'frac': 5 * wbytes * NB / (bms / 1e3) / 1e9 / hbm, 'bytes_per_sample': 40,
                   'ms_two_pass_kernels': float(np.median(bt2[1:])),
                   'note': 'BASELINE config 5 batch (19.66 M samples, inputs + outputs 786 MB > L2), one call = one ticketed '
                           'one-pass launch (+ a memset of its flags); the two-pass kernels used below %d samples are timed beside it'
                           % old_min}
        del rb, ob, wb
    except Exception as e:        # noqa: BLE001
        gae_big = {'error': str(e)[:200]}
    # `traffic` (dram__bytes_read.sum + dram__bytes_write.sum of ONE launch) and pipe utilisation cannot be measured outside a
    # profiler: they are read from profiles/ncu_captures.json, written by tools/ncu_capture_summary.py from the ncu --set full
    # captures of THIS build at this configuration (the file names its .ncu-rep, commit and date), else null
    default_cfg = args.config == 2 and (E, T, tuple(args.hidden)) == (4096, 300, (300, 300)) and args.physics == 'smooth'
    cap = {}
    try:
        cap = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_captures.json'))) if default_cfg else {}
    except Exception:
        cap = {}
    cap_of = lambda k, f: (cap.get(k) or {}).get(f)     # noqa: E731
    flops_env_step = 15 * 2 * 21e3 + 2 * (243 * 300 + 300 * 300 + 300 * 52)      # ABA sweeps + policy MLP, FMA = 2
    # DFMA-pipe peak of the rollout kernel's arithmetic: 64 double-precision FMA lanes per SM and clock (measured:
    # tools/micro/fp64_probe.cu, 2.13 cycles per warp-wide DFMA and sub-partition) x the SM clock sampled during this run
    sm_mhz = clocks.get('sm_mhz') or 1965.0
    dfma_peak = torch.cuda.get_device_properties(device).multi_processor_count * 64 * 2 * sm_mhz * 1e6 / 1e12
    fp64_peak = None
    try:        # live float64 GEMM peak of this box (MEASURED_PEAKS.json carries no FP64 figure)
        sq = torch.randn(6144, 6144, dtype=torch.float64, device=device)
        torch.mm(sq, sq)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            torch.mm(sq, sq)
        b.record()
        b.synchronize()
        fp64_peak = 3 * 2 * 6144 ** 3 / (a.elapsed_time(b) / 1e3) / 1e12
        del sq
    except Exception:
        pass
    roofline = {'kernel': 'rollout_kernel_t4' if args.physics == 'smooth' else 'rollout_kernel (one warp per 32 environments)', 'bound': 'hbm', 'achieved': roll_bytes / (roll_ms / 1e3) / 1e9, 'peak': hbm,
                'unit': 'GB/s', 'frac': roll_bytes / (roll_ms / 1e3) / 1e9 / hbm,
                'traffic': cap_of('rollout_kernel_t4', 'dram_bytes'), 'traffic_source': cap_of('rollout_kernel_t4', 'source'),
                'peak_source': peak_src,
                'ms': roll_ms, 'share_of_step': roll_ms / ms, 'algorithmic_bytes_per_env_step': roll_bytes // N,
                'binding_resource': 'FP64 issue / dependent-issue latency of the tree sweeps (not HBM)',
                'note': 'the contract asks for the HBM figure of the dominant kernel; the fused rollout keeps all state on chip '
                        '(~200 FLOP/B), so its roofline is the DFMA pipe: see roofline_extra.rollout_fp64'}
    extra = {'rollout_fp64': {'achieved': flops_env_step * N / (roll_ms / 1e3) / 1e12, 'peak': dfma_peak, 'unit': 'TFLOP/s',
                              'frac': flops_env_step * N / (roll_ms / 1e3) / 1e12 / dfma_peak,
                              'peak_source': '64 DFMA lanes/SM/clk (tools/micro/fp64_probe.cu) x %d SMs x %.0f MHz sampled in this run'
                                             % (torch.cuda.get_device_properties(device).multi_processor_count, sm_mhz),
                              'flop_per_env_step': flops_env_step,
                              'ncu_fp64_pipe_active_pct': cap_of('rollout_kernel_t4', 'fp64_pipe_active_pct'),
                              'ncu_issue_active_pct': cap_of('rollout_kernel_t4', 'issue_active_pct'),
                              'ncu_source': cap_of('rollout_kernel_t4', 'source')},
             'gae_kernel': {'bound': 'hbm', 'achieved': 5 * wbytes * N / (gae_ms / 1e3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                            'frac': 5 * wbytes * N / (gae_ms / 1e3) / 1e9 / hbm, 'ms': gae_ms, 'bytes_per_sample': 40,
                            'traffic': cap_of('gae', 'dram_bytes'), 'traffic_source': cap_of('gae', 'source')},
             'gae_kernel_config5': (dict(gae_big, traffic=cap_of('gae_config5', 'dram_bytes'), traffic_source=cap_of('gae_config5', 'source'))
                                    if isinstance(gae_big, dict) and 'error' not in gae_big else gae_big),
             'update_dgemm': {'bound': 'tensor(fp64)', 'achieved': None, 'peak': fp64_peak, 'unit': 'TFLOP/s'},
             'rollout_env_substeps_per_s': N * 15 / (roll_ms / 1e3)}
    # the update's dominant kernel alone: [N, 243] x [300, 243]^T float64 product on the int8 tensor cores (first policy /
    # value layer), operands 1.9 GB + output 2.9 GB > L2.  Algorithmic int8 work = 2 M N K x S (S + 1) / 2 slice pairs.
    if agent.gemm == 'ozaki':
        try:
            S_oz = agent.oz_slices
            xg = torch.randn(N, 243, dtype=torch.float64, device=device)
            wg = torch.randn(args.hidden[0], 243, dtype=torch.float64, device=device) / 16
            a_sl, a_ex = lib.oz_slice_rows(xg, S_oz)
            b_sl, b_ex = lib.oz_slice_rows(wg, S_oz)
            og = torch.empty(N, args.hidden[0], dtype=torch.float64, device=device)
            for _ in range(3):
                lib.oz_gemm(a_sl, a_ex, b_sl, b_ex, out=og)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                lib.oz_gemm(a_sl, a_ex, b_sl, b_ex, out=og)
            b.record()
            b.synchronize()
            g_ms = a.elapsed_time(b) / 5
            pairs = S_oz * (S_oz + 1) // 2
            tops = 2.0 * N * args.hidden[0] * 243 * pairs / (g_ms / 1e3) / 1e12
            tp, tp_src = tensor_peak()
            extra['oz_gemm_kernel'] = {'kernel': 'oz_gemm_kernel<%d,80>' % S_oz, 'bound': 'tensor', 'achieved': tops, 'peak': 2 * tp,
                                       'unit': 'TOP/s (int8)', 'frac': tops / (2 * tp), 'ms': g_ms,
                                       'peak_source': '2 x dense bf16 ' + tp_src + ' (kind::i8 issues at twice the bf16 rate)',
                                       'f64_equivalent_tflops': 2.0 * N * args.hidden[0] * 243 / (g_ms / 1e3) / 1e12,
                                       'slices': S_oz, 'slice_pairs': pairs,
                                       'traffic': cap_of('oz_gemm_kernel', 'dram_bytes') if S_oz == 6 else None,
                                       'traffic_source': cap_of('oz_gemm_kernel', 'source') if S_oz == 6 else None,
                                       'note': 'unpadded algorithmic work; the tiles pad N 300->320 and K 243->256'}
            del xg, wg, a_sl, b_sl, og
        except Exception as exc:        # the headline numbers above do not depend on this probe
            extra['oz_gemm_kernel'] = {'error': str(exc)}
    D_in, H1, H2, A_out = 243, args.hidden[0], args.hidden[1], 52
    fwd = lambda o: 2 * N * (D_in * H1 + H1 * H2 + H2 * o)                       # noqa: E731
    bwd = lambda o: 2 * N * (2 * H2 * o + 2 * H1 * H2 + D_in * H1)               # noqa: E731  (no dL/dx of layer 1)
    upd_flops = args.epochs * (fwd(A_out) + fwd(1) + bwd(A_out) + bwd(1))        # the two initial forwards are not counted
    extra['update_dgemm']['achieved'] = upd_flops / (ms_upd / 1e3) / 1e12
    extra['update_dgemm']['frac'] = extra['update_dgemm']['achieved'] / fp64_peak if fp64_peak else None
    extra['update_dgemm']['backend'] = agent.gemm
    extra['update_dgemm']['note'] = ('float64 dense layers on the int8 tensor cores (Ozaki scheme, %d slices): useful float64 FLOPs of the '
                                     'whole update phase incl. slicing / loss / Adam, against the cuBLAS DGEMM peak' % agent.oz_slices
                                     if agent.gemm == 'ozaki' else
                                     'cuBLAS d884 DGEMMs + fused elementwise kernels; whole update phase incl. loss/Adam')

    # ---- the same update on the other dense-layer back ends (cuBLAS DGEMM = the reference's arithmetic, 7 slices = DGEMM-level
    # rounding), timed on the batch of the last rollout: one warm-up + one timed update each, fresh nets / optimizers
    variants = None
    if not args.no_variants and agent.gemm == 'ozaki' and world == 1 and args.task == 'egomimic':
        variants = {}
        for name, gm, sl in (('cublas', 'cublas', None), ('ozaki_7_slices', 'ozaki', 7)):
            try:
                ag2, _ = build_agent(args, device, takes, cnn, gemm=gm, oz_slices=sl)
                ag2.running_state = agent.running_state
                b2, _ = ag2.sample(N, to_host=False)
                ag2.update_params(b2)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ag2.update_params(b2)
                b.record()
                b.synchronize()
                u_ms = a.elapsed_time(b)
                variants[name] = {'ms_update': u_ms, 'value': N / ((ms_roll + u_ms) / 1e3), 'unit': 'env-steps/s',
                                  'note': 'ms_rollout of the timed region + this update'}
                ag2.env.close()
                del ag2, b2
                torch.cuda.empty_cache()
            except Exception as exc:        # noqa: BLE001
                variants[name] = {'error': str(exc)[:200]}

    # ---- e2e through the public API with host trajbatches
    e2e = None
    if not args.no_e2e:
        iteration(True)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        k = max(1, min(args.steps, 2))
        for _ in range(k):
            iteration(True)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) / k], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        S, nu = agent.env.obs_dim, agent.env.md.nu
        d2h = N * (2 * S + nu + 3) * 8 + N * 2 * 4 + 16 * 8
        h2d = N * (S + nu + 3) * 8 + N * 2 * 4
        e2e = {'value': N * world / (t.item() / 1e3), 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': d2h, 'ms_per_step': t.item()}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, detail = cpu_reference(args, takes, cnn, args.hidden, 0)
        cpu = {'value': v, 'unit': 'env-steps/s', 'cores': detail['cores'], 'kind': 'port', 'sample': detail['sample'],
               'rollout_steps_per_s': detail['rollout_steps_per_s'], 'update_samples_per_s': detail['update_samples_per_s']}
    if rank == 0:
        line = {'metric': 'env-steps/sec (humanoid PPO rollout+update)', 'value': value, 'unit': 'env-steps/s',
                'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': args.workload, 'baseline_config': args.config, 'envs_per_gpu': E, 'horizon': T, 'hidden': args.hidden,
                           'epochs': args.epochs, 'parallelism': 'env-sharded dp%d' % world,
                           'physics': ('smooth dynamics (north star)' if args.physics == 'smooth' else
                                       'smooth dynamics + joint limits + floor contact (one-warp rollout kernel)'),
                           'update_gemm': ('float64 in/out on int8 tcgen05 (Ozaki, %d slices, error <= %.0e of row x column maxima)'
                                           % (agent.oz_slices, (agent.oz_slices + 2) * 2.0 ** (-7 * agent.oz_slices))
                                           if agent.gemm == 'ozaki' else 'cuBLAS DGEMM'),
                           'l2': 'inputs larger than L2 (trajbatch %.1f GB per step)' % (N * (2 * 115 + 52) * 8 / 1e9)},
                'ms_rollout': ms_roll, 'ms_update': ms_upd, 'update_variants': variants, 'gpu_launches': launches, 'clocks': clocks,
                'roofline': roofline, 'roofline_extra': extra, 'e2e': e2e, 'cpu_baseline': cpu,
                'avg_c_reward': log.avg_c_reward, 'avg_episode_reward': log.avg_episode_reward,
                'avg_episode_len': log.num_steps / max(1, log.num_episodes),
                'nan_resets': log.num_nan_resets,
                'cons_cap_hits': (int(lib.load().egp_cons_cap_hits(0)) if args.physics == 'full' else None),
                'cons_passes_per_substep': cons_passes(lib) if args.physics == 'full' else None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
