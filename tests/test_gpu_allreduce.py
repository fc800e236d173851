"""gic_allreduce (csrc/allreduce.cu) on ONE GPU: W in-process communicators are each other's peers
(gic_comm_local_group) and all W ranks run as one launch (gic_allreduce_local_group; their CTAs are co-resident by
construction) -- the protocol of the multi-process call (start barrier, reduce-scatter + all-gather over "peer" pointers
in rank order, end barrier, fused square norm) with nothing but the transport taken out.  The multi-GPU run of the same
kernel over NVLink is profiles/dp_check.py (2 and 8 GPUs; logs under profiles/)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


class _Raw:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def _group(world, nfloats):
    from gic_b200 import _lib
    L = _lib.lib()
    comms = [L.gic_comm_create(r, world, nfloats * 4) for r in range(world)]
    assert all(comms), L.gic_last_error()
    arr = (C.c_void_p * world)(*comms)
    _lib.check(L.gic_comm_local_group(arr, world), "gic_comm_local_group")
    bufs = [torch.as_tensor(_Raw(L.gic_comm_buffer(c), nfloats), device="cuda:0") for c in comms]
    return L, comms, arr, bufs


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("n", [4, 1000, 4096 * 13 + 8, 3_000_000])
def test_allreduce_is_rank_order_sum_identical_on_all_ranks_with_fused_sqnorm(world, n):
    from gic_b200 import _lib
    L, comms, arr, bufs = _group(world, n)
    try:
        g = torch.Generator(device="cuda:0").manual_seed(n + world)
        xs = [torch.randn(n, generator=g, device="cuda:0") * (1.0 + r) for r in range(world)]
        want = xs[0].clone()
        for x in xs[1:]:
            want += x                                   # rank order, fp32
        sqs = [torch.zeros(1, device="cuda:0") for _ in range(world)]
        bp = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
        sp = (C.c_void_p * world)(*[s.data_ptr() for s in sqs])
        for rep in range(3):                            # the same channel three times: the flags' epochs advance
            for b, x in zip(bufs, xs):
                b.copy_(x)
            for s in sqs:
                s.zero_()
            with _lib.expect_kernels("allreduce_p2p_kernel"):
                _lib.check(L.gic_allreduce_local_group(arr, bp, sp, n, world, rep % 2, _lib.stream()), "gic_allreduce_local_group")
                torch.cuda.synchronize()
            for r in range(world):
                assert torch.equal(bufs[r], want), f"rank {r} rep {rep}: not the rank-order sum"
                assert torch.equal(sqs[r], sqs[0]), "square norms differ between ranks"
                assert L.gic_comm_error(comms[r]) == 0
            ref = float((want.double() ** 2).sum())
            assert abs(float(sqs[0]) - ref) <= 1e-5 * ref
        # sqnorm accumulates (several buffers may share one norm) and may be NULL
        for b, x in zip(bufs, xs):
            b.copy_(x)
        before = float(sqs[0])
        _lib.check(L.gic_allreduce_local_group(arr, bp, sp, n, world, 2, _lib.stream()), "gic_allreduce_local_group")
        torch.cuda.synchronize()
        assert abs(float(sqs[0]) - 2 * before) <= 1e-5 * before
        for b, x in zip(bufs, xs):
            b.copy_(x)
        _lib.check(L.gic_allreduce_local_group(arr, bp, None, n, world, 3, _lib.stream()), "gic_allreduce_local_group")
        torch.cuda.synchronize()
        assert torch.equal(bufs[world - 1], want)
    finally:
        del bufs
        for c in comms:
            L.gic_comm_destroy(c)


def test_allreduce_rejects_buffers_outside_the_symmetric_allocation_and_bad_sizes():
    from gic_b200 import _lib
    L, comms, arr, bufs = _group(2, 1024)
    try:
        stray = torch.zeros(1024, device="cuda:0")
        bp = (C.c_void_p * 2)(stray.data_ptr(), bufs[1].data_ptr())
        assert L.gic_allreduce_local_group(arr, bp, None, 1024, 2, 0, _lib.stream()) != 0
        bp = (C.c_void_p * 2)(bufs[0].data_ptr(), bufs[1].data_ptr())
        assert L.gic_allreduce_local_group(arr, bp, None, 1022, 2, 0, _lib.stream()) != 0      # n % 4
        assert L.gic_allreduce_local_group(arr, bp, None, 1024, 2, 9, _lib.stream()) != 0      # channel
        assert L.gic_allreduce_local_group(arr, bp, None, 2048, 2, 0, _lib.stream()) != 0      # past the end
        assert L.gic_allreduce(bufs[0].data_ptr(), 1024, comms[0], 0, None, _lib.stream()) != 0    # in-process group
    finally:
        del bufs
        for c in comms:
            L.gic_comm_destroy(c)
