"""debug helper: tiny wide-policy rollout (used under compute-sanitizer)"""
import sys, numpy as np, torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import helpers
import test_gpu_rollout as tr
hidden = tuple(int(x) for x in sys.argv[1:3]) if len(sys.argv) > 2 else (512, 512)
E, T, ctx_dim = 2, 2, 128
orc, model = tr._setup(3, 64, ctx_dim, seed=7)
S, nu = orc.S, orc.nu
w = helpers.policy_weights(S + ctx_dim, hidden[0], hidden[1], nu, seed=3)
rng = np.random.RandomState(9)
rt = rng.randint(0, 3, size=(E, T)); rs = rng.randint(10, 64 - 12 - 10, size=(E, T))
eps = rng.randn(E * T, nu)
pol = orc.make_policy(w['W1'], w['b1'], w['W2'], w['b2'], w['W3'], w['b3'], w['log_std'])
ref = orc.rollout(pol, E, T, rt, rs, eps, n_threads=2)
cu = tr.cu
wd = {k: cu(v.ravel() if k == 'log_std' else v) for k, v in w.items()}
out = model.rollout(wd, E, T, episode_len=12, eps=cu(eps), reset_take=cu(rt, torch.int32), reset_start=cu(rs, torch.int32))
torch.cuda.synchronize()
print(hidden, 'states err', helpers.relerr(out['states'].cpu().numpy(), ref['states']), 'actions err',
      helpers.relerr(out['actions'].cpu().numpy(), ref['actions']))
a = out['states'].cpu().numpy(); r = ref['states']
err = np.abs(a - r)
for row in range(E * T):
    bad = np.nonzero(err[row] > 1e-9)[0]
    print(' row', row, 'n_bad', len(bad), 'first bad dims', bad[:12], 'max', err[row].max())
print(' obs57-59 gpu', a[1,57:60], 'ref', r[1,57:60], 'final_qvel gpu', out['final_qvel'].cpu().numpy()[0,:3], 'ref', ref['final_qvel'][0,:3])
print(' final_qpos err', np.abs(out['final_qpos'].cpu().numpy() - ref['final_qpos']).max(1))
model.close()
