"""World-size-2 data-parallel host logic on CPU (gloo): sharding by batch row + all-reduce(sum) + 1/world scaling
reproduces the single-process full-batch gradients and clip coefficient of the oracle step."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gic_b200  # noqa: F401
        from gic_b200 import parallel
        from oracle import ref_port as rp
        torch.set_num_threads(1)
        cfg = dict(rp.CONFIGS["c0"])                      # unconditional: no BatchNorm, every op is row-local
        inp = rp.make_inputs(cfg)
        B = inp["captions"].shape[0]
        R = inp["args"].disc_num_rep
        full = rp.adversarial_step(inp, 2.0, "standard")
        sl = parallel.shard_rows(B, rank, world)
        shard = dict(inp)
        shard["captions"] = inp["captions"][sl]
        shard["u"] = parallel.shard_uniforms(inp["u"], rank, world, 1)
        shard["keep"] = parallel.shard_uniforms(inp["keep"], rank, world, 1)         # rows b*R + r: contiguous per caption
        assert shard["keep"].shape[1] == (B // world) * R
        part = rp.adversarial_step(shard, 2.0, "standard")
        res = {}
        for net, key in (("d", "d_grads"), ("g", "g_grads")):
            names = sorted(full[key])
            flat = torch.cat([part[key][k].reshape(-1) for k in names]).clone()
            want = torch.cat([full[key][k].reshape(-1) for k in names])
            parallel.allreduce_sum_(flat)
            sq = float((flat.double() ** 2).sum())
            avg = flat * parallel.grad_scale()
            res[net] = (float((avg - want).abs().max()), float(want.abs().max()),
                        parallel.clip_coef(sq, 5.0, parallel.grad_scale()) / parallel.grad_scale(),
                        min(1.0, 5.0 / (float(want.double().norm()) + 1e-6)))
        # the losses are means over B*R logits: the mean of the shard means is the full-batch loss
        l = torch.stack([part["g_loss"], part["d_loss"]]).clone()
        dist.all_reduce(l)
        l /= world
        res["loss"] = (float((l - torch.stack([full["g_loss"], full["d_loss"]])).abs().max()), 1.0, 0.0, 0.0)
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradients_match_full_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = q.get()
    for k in ("d", "g"):
        err, scale, coef, coef_want = res[k]
        assert err <= 2e-5 * scale + 1e-10, (k, err, scale)
        assert abs(coef - coef_want) <= 1e-5
    assert res["loss"][0] <= 1e-6


def test_shard_helpers():
    import gic_b200  # noqa: F401
    from gic_b200 import parallel
    assert parallel.shard_rows(8, 1, 2) == slice(4, 8)
    with pytest.raises(ValueError):
        parallel.shard_rows(7, 0, 2)
    u = torch.arange(2 * 6 * 3).view(2, 6, 3)
    assert torch.equal(parallel.shard_uniforms(u, 2, 3, 1), u[:, 4:6])
    assert parallel.world_size() == 1 and parallel.grad_scale() == 1.0
