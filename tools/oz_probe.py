"""Probe / bring-up of the Ozaki int8 tcgen05 GEMM (csrc/ozaki.cu) on a B200:
slices reconstruct the input, the GEMM equals an exact integer matmul of the same slices bit for bit, error vs the
float64 product, timing against cuBLAS DGEMM.   python tools/oz_probe.py [quick]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from egopose_b200 import lib  # noqa: E402

torch.manual_seed(0)
dev = 'cuda'


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


def check(M, N, K, S, bias=False, relu=False, scale_rows=True, mask=False):
    x = torch.randn(M, K, device=dev, dtype=torch.float64)
    if scale_rows:
        x *= torch.exp(3 * torch.randn(M, 1, device=dev, dtype=torch.float64))
    w = torch.randn(N, K, device=dev, dtype=torch.float64) / K ** 0.5
    a, ea = lib.oz_slice_rows(x, S)
    b, eb = lib.oz_slice_rows(w, S)
    torch.cuda.synchronize()
    ra = recon(a[:, :, :K], ea)
    amax = x.abs().max(1, keepdim=True).values
    e_sl = ((ra - x).abs() / amax).max().item()
    assert (a[:, :, K:] == 0).all(), 'padding not zero'
    bv = torch.randn(N, device=dev, dtype=torch.float64) if bias else None
    mk = torch.randn(M, N, device=dev, dtype=torch.float64) if mask else None
    c = lib.oz_gemm(a, ea, b, eb, bias=bv, relu=relu, mask=mk)
    torch.cuda.synchronize()
    ref = exact_ref(a, ea, b, eb, bv, relu, mk)
    exact = torch.equal(c, ref)
    true = x @ w.t()
    if bias:
        true = true + bv
    if relu:
        true = torch.relu(true)
    if mask:
        true = torch.where(mk > 0, true, torch.zeros_like(true))
    bound = amax * w.abs().max(1).values[None, :] * K
    e_rel = ((c - true).abs() / bound).max().item()
    e_typ = ((c - true).abs().max() / true.abs().max()).item()
    print('M %7d N %4d K %4d S %d bias %d relu %d | slice resid %.2e (<= %.2e) | bit-exact vs integer ref: %s | err/bound %.2e  err/max|C| %.2e'
          % (M, N, K, S, bias, relu, e_sl, 2.0 ** (1 - _rb() * S), exact, e_rel, e_typ), flush=True)
    if not exact:
        d = (c - ref).abs()
        bad = (d > 0).nonzero()
        print('   mismatches %d of %d, first at %s: got %r want %r' % (bad.shape[0], c.numel(), bad[0].tolist(),
                                                                       c[tuple(bad[0])].item(), ref[tuple(bad[0])].item()))
        rows = torch.unique(bad[:, 0])[:20].tolist()
        cols = torch.unique(bad[:, 1])[:20].tolist()
        print('   bad rows (first 20)', rows, 'bad cols (first 20)', cols)
    return exact


def check_wgrad(Ns, F1, F2, S):
    """dW [F1, F2] = dY^T X over Ns samples through the transposed column-scaled slices + split-K"""
    dy = torch.randn(Ns, F1, device=dev, dtype=torch.float64) * torch.exp(torch.randn(Ns, 1, device=dev, dtype=torch.float64))
    x = torch.relu(torch.randn(Ns, F2, device=dev, dtype=torch.float64))
    a, ea = lib.oz_slice_colsT(dy, S, lib.oz_colmax(dy))
    b, eb = lib.oz_slice_colsT(x, S, lib.oz_colmax(x), ones_row=True)
    torch.cuda.synchronize()
    ra = recon(a[:, :, :Ns], ea)
    e_sl = ((ra - dy.t()).abs() / dy.abs().max(0).values[:, None]).max().item()
    c = lib.oz_gemm(a, ea, b, eb)
    torch.cuda.synchronize()
    true = dy.t() @ torch.cat([x, torch.ones(Ns, 1, device=dev, dtype=torch.float64)], 1)
    e_typ = ((c - true).abs().max() / true.abs().max()).item()
    ok = True
    if Ns <= 70000:
        ref = exact_ref(a, ea, b, eb)
        ok = torch.allclose(c, ref, rtol=1e-14, atol=0)        # split-K partial sums are added in float64
    print('wgrad Ns %8d F1 %3d F2 %3d S %d | slice resid %.2e | matches integer ref: %s | err/max|C| %.2e' % (Ns, F1, F2, S, e_sl, ok, e_typ), flush=True)
    return ok


def bench(M, N, K, S, iters=10):
    x = torch.randn(M, K, device=dev, dtype=torch.float64)
    w = torch.randn(N, K, device=dev, dtype=torch.float64)
    a, ea = lib.oz_slice_rows(x, S)
    b, eb = lib.oz_slice_rows(w, S)
    out = torch.empty(M, N, device=dev, dtype=torch.float64)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for _ in range(3):
        lib.oz_gemm(a, ea, b, eb, out=out)
    ev[0].record()
    for _ in range(iters):
        lib.oz_gemm(a, ea, b, eb, out=out)
    ev[1].record()
    out2 = torch.empty(M, N, device=dev, dtype=torch.float64)
    for _ in range(2):
        torch.mm(x, w.t(), out=out2)
    ev[2].record()
    for _ in range(iters):
        torch.mm(x, w.t(), out=out2)
    ev[3].record()
    torch.cuda.synchronize()
    t_oz, t_bl = ev[0].elapsed_time(ev[1]) / iters, ev[2].elapsed_time(ev[3]) / iters
    fl = 2.0 * M * N * K
    sl = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    sl[0].record()
    for _ in range(iters):
        lib.oz_slice_rows(x, S, out=(a, ea))
    sl[1].record()
    torch.cuda.synchronize()
    t_sl = sl[0].elapsed_time(sl[1]) / iters
    print('bench M %8d N %4d K %4d S %d | ozaki gemm %.3f ms (%.1f TFLOP/s f64-equivalent) | cuBLAS DGEMM %.3f ms (%.1f TFLOP/s) | slicing A %.3f ms (%.0f GB/s)'
          % (M, N, K, S, t_oz, fl / t_oz * 1e-9, t_bl, fl / t_bl * 1e-9, t_sl, M * K * (8 + S) / t_sl * 1e-6), flush=True)


def bench_wgrad(Ns, F1, F2, S, iters=20):
    dy = torch.randn(Ns, F1, device=dev, dtype=torch.float64)
    x = torch.relu(torch.randn(Ns, F2, device=dev, dtype=torch.float64))
    a, ea = lib.oz_slice_colsT(x, S, lib.oz_colmax(x), ones_row=True)
    b, eb = lib.oz_slice_colsT(dy, S, lib.oz_colmax(dy))
    out = torch.empty(F2 + 1, F1, device=dev, dtype=torch.float64)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    for _ in range(3):
        lib.oz_gemm(a, ea, b, eb, out=out)
    ev[0].record()
    for _ in range(iters):
        lib.oz_gemm(a, ea, b, eb, out=out)
    ev[1].record()
    out2 = torch.empty(F2, F1, device=dev, dtype=torch.float64)
    for _ in range(2):
        torch.mm(x.t(), dy, out=out2)
    ev[2].record()
    for _ in range(iters):
        torch.mm(x.t(), dy, out=out2)
    ev[3].record()
    cm = torch.zeros(F2, device=dev, dtype=torch.float64)
    ev[4].record()
    for _ in range(iters):
        lib.oz_colmax(x, cm)
        lib.oz_slice_colsT(x, S, cm, out=(a, ea), ones_row=True)
    ev[5].record()
    torch.cuda.synchronize()
    t_oz, t_bl, t_sl = ev[0].elapsed_time(ev[1]) / iters, ev[2].elapsed_time(ev[3]) / iters, ev[4].elapsed_time(ev[5]) / iters
    fl = 2.0 * Ns * F1 * F2
    print('bench wgrad Ns %8d F1 %3d F2 %3d S %d | ozaki %.3f ms (%.1f TFLOP/s f64-eq) | cuBLAS %.3f ms (%.1f TFLOP/s) | colmax+sliceT %.3f ms (%.0f GB/s)'
          % (Ns, F1, F2, S, t_oz, fl / t_oz * 1e-9, t_bl, fl / t_bl * 1e-9, t_sl, Ns * F2 * (16 + S) / t_sl * 1e-6), flush=True)


if __name__ == '__main__':
    quick = 'quick' in sys.argv
    ok = True
    ok &= check(128, 64, 64, 4, scale_rows=False)
    ok &= check(128, 64, 128, 4)
    ok &= check(100, 52, 200, 5, bias=True)
    ok &= check(1000, 300, 243, 6, bias=True, relu=True)
    ok &= check(4096 + 17, 304, 300, 3)
    ok &= check(513, 1, 300, 6, bias=True)
    ok &= check(2000, 300, 640, 6)
    ok &= check(18944, 300, 300, 6, bias=False, mask=True)
    ok &= check(3000, 244, 52, 6)
    ok &= check_wgrad(5000, 52, 300, 5)
    ok &= check_wgrad(65536 + 100, 300, 243, 6)
    print('ALL OK' if ok else 'FAILURES', flush=True)
    if not quick and ok and "full" not in sys.argv:
        for S in (4, 5, 6):
            bench(1228800, 300, 243, S)
        bench(18944, 300, 300, 6, iters=50)
        bench_wgrad(18944, 300, 300, 6)
    if "full" in sys.argv and ok:
        for S in (4, 5, 6):
            bench(1228800, 300, 243, S)
            bench(1228800, 300, 300, S)
        bench(18944, 300, 300, 6, iters=50)
        bench(18944, 52, 300, 6, iters=50)
        bench_wgrad(18944, 300, 300, 6)
        bench_wgrad(1228800, 300, 300, 6)
        ok &= check_wgrad(18944, 300, 300, 6)
        ok &= check_wgrad(1228800, 300, 300, 6)
        t0 = time.time()
