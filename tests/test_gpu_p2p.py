"""Peer-memory gradient all-reduce (csrc/p2p.cu, egp_allreduce_grads_f64) with all ranks inside ONE process
(egp_comm_connect_local): the kernels of the ranks run on separate streams of cuda:0 and meet at the flag barriers exactly as
they do across GPUs.  The multi-process / multi-GPU wiring (CUDA IPC) is covered by tests/test_gpu_scale.py
(two ranks on one GPU) and tools/check_multi_gpu.py (2 - 8 GPUs)."""
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


# (8 ranks on ONE device must all be resident at once to meet at the barrier: small grids there; the headline size with 2)
@pytest.mark.parametrize('world,n', [(2, 1000), (3, 1001), (8, 20_001), (2, 278_000 + 7), (1, 5)])
def test_peer_allreduce_is_the_rank_ordered_sum(world, n):
    from egopose_b200 import lib
    comms = lib.PeerComm.local_group(n, ['cuda:0'] * world)
    streams = [torch.cuda.Stream() for _ in range(world)]
    try:
        for rep in range(3):                              # successive epochs reuse the blocks
            g = torch.Generator(device='cuda').manual_seed(rep)
            srcs = [torch.randn(n, dtype=torch.float64, device='cuda', generator=g) * 10.0 ** (r - 2) for r in range(world)]
            for c, s in zip(comms, srcs):
                c.src.copy_(s)
            torch.cuda.synchronize()
            for c, st in zip(comms, streams):
                with torch.cuda.stream(st):
                    c.allreduce()
            torch.cuda.synchronize()
            want = srcs[0].clone()
            for s in srcs[1:]:
                want += s                                 # rank order, like the kernel
            for c in comms:
                assert c.error() == 0
                assert torch.equal(c.out, want)           # bit-identical on every rank
                assert torch.equal(c.src, srcs[c.rank])   # the exchange does not touch the local gradient
    finally:
        for c in comms:
            c.close()


def test_peer_allreduce_times_out_instead_of_hanging():
    """a rank that never arrives: the barrier gives up (error word 1) and the kernel returns; nothing hangs"""
    from egopose_b200 import lib
    L = lib.load()
    old = L.egp_comm_set_timeout_cycles(200_000_000)      # ~0.1 s
    comms = lib.PeerComm.local_group(64, ['cuda:0', 'cuda:0'])
    try:
        comms[0].src.fill_(1.0)
        comms[0].allreduce()                              # rank 1 never calls
        torch.cuda.synchronize()
        assert comms[0].error() & 1
        assert comms[1].error() == 0
    finally:
        L.egp_comm_set_timeout_cycles(old)
        for c in comms:
            c.close()
