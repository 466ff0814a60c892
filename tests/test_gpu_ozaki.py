"""Float64 dense layers on the int8 tensor cores (csrc/ozaki.cu, csrc/oz_mlp.cu).

* the slices reconstruct their input to 2^(1-7S) of the row / column maximum
* the tcgen05 GEMM equals an exact integer matmul of the same slices BIT FOR BIT (integer accumulation in Tensor Memory,
  exact int64 Horner, one rounding) for every tile shape / split / epilogue option
* the chunked MLP step (forward + PPO / value loss + backward) matches torch float64 autograd of the reference's graph
  (models/mlp.py:22-25, core/policy_gaussian.py:19-24, core/critic.py:15-18, agents/agent_ppo.py:58-65,
  agents/agent_pg.py:22-23) far inside the north-star tolerance (1e-5 on the PPO loss)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from egopose_b200 import lib  # noqa: E402

DEV = 'cuda'


def _rb():
    return lib.load().egp_oz_radix_bits()


def digits(sl):
    """slice bytes -> digit values (radix 256: top slice signed, the others unsigned; radix 128: all signed)"""
    d = sl.double()
    if _rb() == 8 and sl.shape[0] > 1:
        d[1:] = sl[1:].view(torch.uint8).double()
    return d


def recon(sl, ex):
    S, rb = sl.shape[0], _rb()
    d = digits(sl)
    v = torch.zeros(sl.shape[1:], dtype=torch.float64, device=sl.device)
    for t in range(S):
        v += d[t] * 2.0 ** (rb * (S - 1 - t))
    return v * torch.ldexp(torch.ones_like(ex, dtype=torch.float64), ex + 1 - rb * S)[:, None]


def exact_ref(a, ea, b, eb, bias=None, relu=False, mask=None):
    """the kernel's arithmetic restated with exact float64 integer matmuls: one accumulator per d = t + u, two exact
    Horner groups joined by one rounding, scales, bias, relu, mask"""
    S, rb = a.shape[0], _rb()
    A, B = digits(a), digits(b)
    acc = []
    for d in range(S):
        s = torch.zeros((a.shape[1], b.shape[1]), dtype=torch.float64, device=a.device)
        for t in range(d + 1):
            s += A[t] @ B[d - t].t()
        acc.append(s)
    G = min(S, 3 if rb == 8 else 4)
    base = 2.0 ** rb
    hi = acc[0].clone()
    for d in range(1, G):
        hi = hi * base + acc[d]
    h = hi * 2.0 ** (-rb * (G - 1))
    if S > G:
        lo = acc[G].clone()
        for d in range(G + 1, S):
            lo = lo * base + acc[d]
        h = h + lo * 2.0 ** (-rb * (S - 1))
    h = h * (torch.ldexp(torch.ones_like(ea, dtype=torch.float64), ea + 2 - 2 * rb)[:, None]
             * torch.ldexp(torch.ones_like(eb, dtype=torch.float64), eb)[None, :])
    if bias is not None:
        h = h + bias[None, :]
    if relu:
        h = torch.relu(h)
    if mask is not None:
        h = torch.where(mask > 0, h, torch.zeros_like(h))
    return h


@pytest.mark.parametrize('M,N,K,S,bias,relu,mask', [
    (128, 64, 64, 4, False, False, False),
    (100, 52, 200, 5, True, False, False),
    (1000, 300, 243, 6, True, True, False),
    (4113, 304, 300, 3, False, False, False),
    (513, 1, 300, 6, True, False, False),          # odd leading dimension: direct-store epilogue
    (2000, 300, 640, 6, False, False, False),      # long contraction
    (3000, 300, 300, 6, False, False, True),       # relu-backward mask
    (777, 244, 52, 6, False, False, False),
])
def test_gemm_bit_exact_vs_integer_reference(M, N, K, S, bias, relu, mask):
    torch.manual_seed(M + N + K + S)
    x = torch.randn(M, K, device=DEV, dtype=torch.float64) * torch.exp(3 * torch.randn(M, 1, device=DEV, dtype=torch.float64))
    w = torch.randn(N, K, device=DEV, dtype=torch.float64) / K ** 0.5
    a, ea = lib.oz_slice_rows(x, S)
    b, eb = lib.oz_slice_rows(w, S)
    rb = _rb()
    if rb == 7:
        assert a.abs().max() <= 64 and b.abs().max() <= 64
    assert (a[:, :, K:] == 0).all()
    amax = x.abs().max(1, keepdim=True).values
    assert ((recon(a[:, :, :K], ea) - x).abs() / amax).max().item() <= 2.0 ** (2 - rb * S)
    bv = torch.randn(N, device=DEV, dtype=torch.float64) if bias else None
    mk = torch.randn(M, N, device=DEV, dtype=torch.float64) if mask else None
    c = lib.oz_gemm(a, ea, b, eb, bias=bv, relu=relu, mask=mk)
    assert torch.equal(c, exact_ref(a, ea, b, eb, bv, relu, mk))
    true = x @ w.t()
    if bias:
        true = true + bv
    if relu:
        true = torch.relu(true)
    if mask:
        true = torch.where(mk > 0, true, torch.zeros_like(true))
    bound = amax * w.abs().max(1).values[None, :] * K
    assert ((c - true).abs() / bound).max().item() <= 8 * (S + 2) * 2.0 ** (-rb * S)


def test_weight_gradient_gemm_with_ones_row():
    """contraction over the samples: transposed column-scaled slices, split-K, bias gradient from the row of ones"""
    torch.manual_seed(3)
    Ns, F1, F2, S = 20000 + 37, 52, 300, 6
    dy = torch.randn(Ns, F1, device=DEV, dtype=torch.float64) * torch.exp(torch.randn(Ns, 1, device=DEV, dtype=torch.float64))
    x = torch.relu(torch.randn(Ns, F2, device=DEV, dtype=torch.float64))
    a, ea = lib.oz_slice_colsT(x, S, lib.oz_colmax(x), ones_row=True)
    b, eb = lib.oz_slice_colsT(dy, S, lib.oz_colmax(dy))
    assert a.shape[1] == F2 + 1 and ea[F2].item() == 1
    assert torch.equal(recon(a[:, F2:, :Ns], ea[F2:]), torch.ones(1, Ns, device=DEV, dtype=torch.float64))
    assert (a[:, :, Ns:] == 0).all()
    c = lib.oz_gemm(a, ea, b, eb)
    assert torch.allclose(c, exact_ref(a, ea, b, eb), rtol=1e-14, atol=0)     # split-K partials are summed in float64
    true = torch.cat([x, torch.ones(Ns, 1, device=DEV, dtype=torch.float64)], 1).t() @ dy
    assert ((c - true).abs().max() / true.abs().max()).item() < 1e-9


def _torch_mlp(W, x):
    h1 = torch.relu(x @ W[0].t() + W[1])
    h2 = torch.relu(h1 @ W[2].t() + W[3])
    return h2 @ W[4].t() + W[5]


def _weights(dims, seed):
    g = torch.Generator(device='cpu').manual_seed(seed)
    i, h1, h2, o = dims
    shapes = [(h1, i), (h1,), (h2, h1), (h2,), (o, h2), (o,)]
    return [(torch.randn(s, generator=g, dtype=torch.float64) * (0.3 if len(s) == 1 else 1.0 / s[-1] ** 0.5)).to(DEV) for s in shapes]


@pytest.mark.parametrize('n,chunk', [(1000, 256), (300, 1024), (2500, 1024), (5, 128), (129, 128)])
def test_mlp_step_value_loss_matches_autograd(n, chunk):
    dims = (40, 64, 48, 1)
    W = _weights(dims, 1)
    torch.manual_seed(n)
    x = torch.randn(n, dims[0], device=DEV, dtype=torch.float64)
    ret = torch.randn(n, device=DEV, dtype=torch.float64)
    oz = lib.OzMlp(*dims, n_slices=6, chunk_rows=chunk, device=DEV)
    y = oz.step(W, x)
    Wt = [w.clone().requires_grad_(True) for w in W]
    yt = _torch_mlp(Wt, x)
    assert torch.allclose(y, yt.detach(), rtol=1e-10, atol=1e-11)
    loss_t = ((yt.view(-1) - ret) ** 2).mean()
    loss_t.backward()
    grads = [torch.full_like(w, float('nan')) for w in W]
    loss = torch.zeros(1, device=DEV, dtype=torch.float64)
    cache = oz.new_cache(n)
    for rep in range(2):            # second pass reads the cached input slices: identical bits
        loss.zero_()
        oz.step(W, x, grads=grads, loss=dict(kind='value', returns=ret, inv_n=1.0 / n, loss=loss), cache=cache)
        assert abs(loss.item() - loss_t.item()) < 1e-10 * abs(loss_t.item())
        for g, wt in zip(grads, Wt):
            assert torch.allclose(g, wt.grad, rtol=1e-8, atol=1e-10 * wt.grad.abs().max().item())
        if rep == 0:
            first = [g.clone() for g in grads]
    assert all(torch.equal(a, b) for a, b in zip(first, grads))


def test_mlp_step_ppo_loss_matches_autograd():
    dims = (243, 300, 300, 52)
    n, chunk = 3000, 1024
    W = _weights(dims, 2)
    torch.manual_seed(5)
    x = torch.randn(n, dims[0], device=DEV, dtype=torch.float64)
    log_std = torch.full((dims[3],), -2.3, device=DEV, dtype=torch.float64)
    mu0 = _torch_mlp(W, x)
    actions = mu0 + torch.exp(log_std) * torch.randn(n, dims[3], device=DEV, dtype=torch.float64)
    logp0 = lib.gauss_logp(mu0 + 0.02 * torch.randn_like(mu0), actions, log_std)
    adv = torch.randn(n, device=DEV, dtype=torch.float64) * 2 + 0.3
    exps = (torch.rand(n, device=DEV) > 0.2).double()
    stats = torch.tensor([float(n), adv.mean().item(), ((adv - adv.mean()) ** 2).sum().item()], device=DEV, dtype=torch.float64)
    inv_count = 1.0 / exps.sum().item()
    # torch reference (agents/agent_ppo.py:58-65)
    Wt = [w.clone().requires_grad_(True) for w in W]
    mu = _torch_mlp(Wt, x)
    var = torch.exp(log_std) ** 2
    logp = (-(actions - mu) ** 2 / (2 * var) - 0.5 * np.log(2 * np.pi) - log_std).sum(1)
    advn = (adv - adv.mean()) / adv.std()
    ratio = torch.exp(logp - logp0)
    surr = -torch.min(ratio * advn, torch.clamp(ratio, 0.8, 1.2) * advn)
    loss_t = (surr * exps).sum() * inv_count
    loss_t.backward()
    oz = lib.OzMlp(*dims, n_slices=6, chunk_rows=chunk, device=DEV)
    grads = [torch.zeros_like(w) for w in W]
    loss = torch.zeros(1, device=DEV, dtype=torch.float64)
    oz.step(W, x, grads=grads, loss=dict(kind='ppo', actions=actions, log_std=log_std, adv=adv, stats=stats, logp0=logp0, exps=exps,
                                          clip_eps=0.2, inv_count=inv_count, dlogstd=None, loss=loss))
    assert abs(loss.item() - loss_t.item()) < 1e-9 * abs(loss_t.item())          # north star: 1e-5
    for g, wt in zip(grads, Wt):
        assert torch.allclose(g, wt.grad, rtol=1e-7, atol=1e-9 * wt.grad.abs().max().item())


@pytest.mark.parametrize('dxc', [16, 40])
def test_mlp_step_input_gradient_matches_autograd(dxc):
    """dL/dx[:, :c]: the gradient handed to a learned video-context net (agent_ego.py:28-32)"""
    dims = (40, 64, 48, 1)
    n, chunk = 700, 256
    W = _weights(dims, 7)
    torch.manual_seed(11)
    x = torch.randn(n, dims[0], device=DEV, dtype=torch.float64)
    ret = torch.randn(n, device=DEV, dtype=torch.float64)
    xt = x.clone().requires_grad_(True)
    Wt = [w.clone().requires_grad_(True) for w in W]
    ((_torch_mlp(Wt, xt).view(-1) - ret) ** 2).mean().backward()
    oz = lib.OzMlp(*dims, n_slices=6, chunk_rows=chunk, device=DEV)
    grads = [torch.zeros_like(w) for w in W]
    loss = torch.zeros(1, device=DEV, dtype=torch.float64)
    dx = torch.full((n, dxc), float('nan'), device=DEV, dtype=torch.float64)
    oz.step(W, x, grads=grads, loss=dict(kind='value', returns=ret, inv_n=1.0 / n, loss=loss), dx=dx)
    ref = xt.grad[:, :dxc]
    assert torch.allclose(dx, ref, rtol=1e-8, atol=1e-10 * ref.abs().max().item())
    for g, wt in zip(grads, Wt):
        assert torch.allclose(g, wt.grad, rtol=1e-8, atol=1e-10 * wt.grad.abs().max().item())


@pytest.mark.parametrize('dims,n,chunk,dxc', [((243, 300, 300, 52), 3000, 1024, 0), ((243, 300, 300, 1), 2900, 1024, 0),
                                              ((40, 64, 48, 1), 700, 256, 16), ((243, 512, 512, 52), 1500, 640, 24),
                                              ((37, 50, 70, 3), 333, 128, 37)])
def test_mlp_step_fused_slicing_is_bit_identical(dims, n, chunk, dxc):
    """Producer-recorded abs-maxima + one-read two-orientation slicing (default) against the separate row / column-maximum /
    transposed passes: the same exponents and digits, so every gradient must agree BIT FOR BIT."""
    W = _weights(dims, 3)
    torch.manual_seed(17)
    x = torch.randn(n, dims[0], device=DEV, dtype=torch.float64)
    x[5] = 0.0                                           # an all-zero sample row
    od = dims[3]
    if od == 1:
        ls = dict(kind='value', returns=torch.randn(n, device=DEV, dtype=torch.float64), inv_n=1.0 / n)
    else:
        log_std = torch.full((od,), -2.3, device=DEV, dtype=torch.float64)
        mu0 = _torch_mlp(W, x)
        actions = mu0 + torch.exp(log_std) * torch.randn(n, od, device=DEV, dtype=torch.float64)
        adv = torch.randn(n, device=DEV, dtype=torch.float64)
        exps = (torch.rand(n, device=DEV) > 0.3).double()                # rows with zero gradient
        ls = dict(kind='ppo', actions=actions, log_std=log_std, adv=adv,
                  stats=torch.tensor([float(n), adv.mean().item(), ((adv - adv.mean()) ** 2).sum().item()], device=DEV, dtype=torch.float64),
                  logp0=lib.gauss_logp(mu0 + 0.02 * torch.randn_like(mu0), actions, log_std), exps=exps, clip_eps=0.2,
                  inv_count=1.0 / exps.sum().item(), dlogstd=None)
    out = {}
    L = lib.load()
    prev = L.egp_oz_mlp_set_fused_slicing(-1)
    try:
        for mode in (0, 1):
            assert L.egp_oz_mlp_set_fused_slicing(mode) == mode
            oz = lib.OzMlp(*dims, n_slices=6, chunk_rows=chunk, device=DEV)
            grads = [torch.full_like(w, float('nan')) for w in W]
            loss = torch.zeros(1, device=DEV, dtype=torch.float64)
            dx = torch.full((n, dxc), float('nan'), device=DEV, dtype=torch.float64) if dxc else None
            oz.step(W, x, grads=grads, loss=dict(ls, loss=loss), dx=dx)
            torch.cuda.synchronize()
            out[mode] = (grads, loss.clone(), dx)
    finally:
        L.egp_oz_mlp_set_fused_slicing(prev)
    assert abs(out[0][1].item() - out[1][1].item()) <= 1e-13 * abs(out[0][1].item())    # the loss is summed with atomics: order-dependent
    for a, b in zip(out[0][0], out[1][0]):
        assert torch.isfinite(a).all() and torch.equal(a, b)
    if dxc:
        assert torch.equal(out[0][2], out[1][2])


@pytest.mark.parametrize('M,N,K,relu,mask', [(300, 300, 243, True, False), (1000, 512, 300, True, False), (129, 52, 64, False, False),
                                             (700, 300, 52, False, True), (5, 7, 20, False, False), (2048, 512, 512, False, True)])
def test_gemm_records_maxima_and_slice_both_matches_separate_slicers(M, N, K, relu, mask):
    """egp_oz_gemm_max_f64 records the row / column abs-maxima of its FINAL output; egp_oz_slice_both_f64 then writes exactly the
    bytes of egp_oz_slice_rows_f64 + egp_oz_slice_cols_t_f64 from one read."""
    S = 6
    torch.manual_seed(M + N)
    a = torch.randn(M, K, device=DEV, dtype=torch.float64) * torch.exp(2 * torch.randn(M, 1, device=DEV, dtype=torch.float64))
    b = torch.randn(N, K, device=DEV, dtype=torch.float64)
    a[M // 2] = 0.0
    bias = torch.randn(N, device=DEV, dtype=torch.float64) if relu else None
    mk = torch.randn(M, N, device=DEV, dtype=torch.float64) if mask else None
    sa, ea = lib.oz_slice_rows(a, S)
    sb, eb = lib.oz_slice_rows(b, S)
    ref = lib.oz_gemm(sa, ea, sb, eb, bias=bias, relu=relu, mask=mk)
    rowmax = torch.zeros(M, dtype=torch.int32, device=DEV)
    colmax = torch.zeros(N, dtype=torch.float64, device=DEV)
    c = lib.oz_gemm(sa, ea, sb, eb, bias=bias, relu=relu, mask=mk, rowmax=rowmax, colmax=colmax)
    assert torch.equal(c, ref)
    hi = lambda t: (t.contiguous().view(torch.int64) >> 32).to(torch.int32)          # noqa: E731
    assert torch.equal(rowmax, hi(c.abs().max(1).values))
    assert torch.equal(hi(colmax), hi(c.abs().max(0).values))
    assert (colmax.view(torch.int64) & 0xffffffff).eq(0).all()
    for ones in (False, True):
        cm = torch.zeros(N, dtype=torch.float64, device=DEV)
        r_ref = lib.oz_slice_rows(c, S, colmax=cm)
        t_ref = lib.oz_slice_colsT(c, S, cm, ones_row=ones)
        (r, er), (t, et) = lib.oz_slice_both(c, S, rowmax, colmax, ones_row=ones)
        assert torch.equal(er, r_ref[1]) and torch.equal(et, t_ref[1])
        assert torch.equal(t, t_ref[0])
        kp16 = r_ref[0].shape[2]
        assert torch.equal(r[:, :, :kp16], r_ref[0]) and r[:, :, N:].eq(0).all()
