"""North-star EXTENSIONS (inverse-CDF sampler, Monte-Carlo rollouts, reward / baseline / policy-gradient loss) on the
GPU against oracle/ref_ext.py.  PARITY UNPINNED BY REFERENCE: the reference has none of these;
the oracle is this repo's own definition.  Token ids must be bit-exact except where the uniform draw lies within
1e-6 of a CDF boundary: such ties are counted and reported (gpurun_out/ext_report.json)."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import ref_ext as rx
from oracle import ref_port as rp

pytestmark = pytest.mark.gpu
REPORT = {}
RTOL = 1e-3


@pytest.fixture(scope="module", autouse=True)
def _report():
    yield
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "ext_report.json"), "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


def L_():
    import gic_b200
    from gic_b200 import _lib
    _lib.require_cuda()
    return _lib


def close(name, got, want, rtol=RTOL, atol=0.0):
    got = got.detach().double().cpu().reshape(-1)
    want = want.detach().double().cpu().reshape(-1)
    assert got.shape == want.shape, name
    scale = float(want.abs().max()) if want.numel() else 0.0
    err = float((got - want).abs().max()) if want.numel() else 0.0
    REPORT[name] = dict(err=err, scale=scale)
    assert err <= atol + rtol * max(scale, 1e-30), f"{name}: err {err:.3e} scale {scale:.3e}"


def instructor(inp):
    from gic_b200.training import GANInstructor
    a = inp["args"]; a.device = "cuda"
    inst = GANInstructor(a, device="cuda:0")
    sd = inst.gen.state_dict(); sd.update({k: v.clone() for k, v in inp["gen"].items()}); inst.gen.load_state_dict(sd)
    inst.disc.load_state_dict({k: v.clone() for k, v in inp["disc"].items()})
    inst.gen.train(); inst.disc.train()
    return inst


def count_ties(name, logits, u, got, want):
    """ids equal except draws within 1e-6 of a CDF boundary (counted)."""
    bad = (got != want).nonzero().flatten()
    ties = 0
    if bad.numel():
        gap = rx.cdf_boundary_gap(logits[bad], u[bad], got[bad], want[bad])
        assert float(gap.max()) < 1e-6, f"{name}: token mismatch away from a CDF boundary (gap {float(gap.max()):.3e})"
        ties = int(bad.numel())
    REPORT[name] = dict(ties=ties, total=int(want.numel()))
    return ties


@pytest.mark.parametrize("B,V,scale", [(7, 50, 1.0), (64, 1000, 3.0), (256, 10000, 1.0), (5, 60001, 2.0), (9, 130, 8.0)])
def test_sample_cdf_step(B, V, scale):
    L = L_()
    g = torch.Generator().manual_seed(B * 31 + V)
    logits = torch.randn(B, V, generator=g) * scale
    u = torch.rand(B, generator=g)
    u[0] = 0.0
    if B > 1:
        u[1] = 0.9999999
    E = 8
    emb = torch.randn(V, E, generator=g)
    want = rp.sample_inverse_cdf(logits, u)
    d = torch.device("cuda:0")
    lg, ud, embd = logits.to(d), u.to(d), emb.to(d)
    ids = torch.full((B, 3), -1, dtype=torch.int64, device=d)
    logp = torch.zeros(B, 3, device=d)
    out = torch.zeros(B, 3, V, device=d)
    xn = torch.zeros(B, E, device=d)
    L.check(L.lib().gic_sample_cdf_step(L.ptr(lg), L.ptr(ud), B, V, 3, 1, L.ptr(out), L.ptr(ids), L.ptr(logp), None,
                                        L.ptr(embd), E, L.ptr(xn), L.stream()), "gic_sample_cdf_step")
    torch.cuda.synchronize()
    got = ids[:, 1].cpu()
    count_ties(f"cdf_step/{B}x{V}", logits, u, got, want)
    assert int(got[0]) == 0                                       # u = 0 -> first token
    assert torch.equal(out[:, 1].cpu(), logits)                   # raw logits copied
    assert torch.equal(xn.cpu(), emb[got])                        # next-input gather
    close(f"cdf_step/{B}x{V}/logp", logp[:, 1], F.log_softmax(logits, -1).gather(1, got[:, None])[:, 0], rtol=1e-5, atol=1e-6)


def test_sample_cdf_frequencies_match_softmax():
    """Size-independent property: with many uniform draws the empirical token frequencies follow softmax(logits)."""
    L = L_()
    V, N = 37, 200000
    g = torch.Generator().manual_seed(11)
    row = torch.randn(V, generator=g) * 1.5
    d = torch.device("cuda:0")
    logits = row.to(d).expand(N, V).contiguous()
    u = torch.rand(N, generator=g).to(d)
    ids = torch.empty(N, 1, dtype=torch.int64, device=d)
    L.check(L.lib().gic_sample_cdf_step(L.ptr(logits), L.ptr(u), N, V, 1, 0, None, L.ptr(ids), None, None, None, 0, None,
                                        L.stream()), "gic_sample_cdf_step")
    freq = torch.bincount(ids[:, 0].cpu(), minlength=V).double() / N
    p = F.softmax(row.double(), -1)
    sigma = (p * (1 - p) / N).sqrt()
    assert float(((freq - p).abs() / sigma).max()) < 5.0


@pytest.mark.parametrize("cfg_name", ["c0", "c1"])
def test_cdf_decode_and_rollouts_vs_oracle(cfg_name):
    inp = rp.make_inputs(rp.CONFIGS[cfg_name])
    a = inp["args"]
    B, Lc = inp["captions"].shape
    n = 3
    g = torch.Generator().manual_seed(77)
    u = torch.rand(Lc, B, generator=g)
    u_roll = torch.rand(Lc, (Lc - 1) * B * n, generator=g)
    ref = rx.pg_step(inp, u, u_roll, n, baseline_mode=0)
    inst = instructor(inp)
    out = inst.pg_step(inp["captions"], pooled=inp["pooled"], u=u, u_roll=u_roll, n_roll=n, baseline_mode=0, update=False,
                       d_update=False)
    torch.cuda.synchronize()
    V = a.vocab_size
    # sampled captions: bit-exact (the decode is not teacher-forced here: a tie would diverge, report it)
    ids = out["ids"].cpu()
    same = torch.equal(ids, ref["ids"])
    REPORT[f"pg/{cfg_name}/ids_equal"] = bool(same)
    assert same, "sampled captions differ from the oracle (check gpurun_out/ext_report.json for ties)"
    close(f"pg/{cfg_name}/logits", out["logits"], ref["logits"])
    close(f"pg/{cfg_name}/logp", out["logp"], ref["logp"], rtol=1e-4, atol=1e-5)
    roll = out["roll_ids"].cpu()
    nbad = int((roll != ref["roll_ids"]).any(1).sum())
    REPORT[f"pg/{cfg_name}/rollout_rows_differing"] = dict(rows=nbad, total=int(roll.shape[0]))
    assert nbad == 0
    # structure: prefixes are copies of the sampled caption
    for t in range(1, Lc):
        r0 = (t - 1) * B * n
        assert torch.equal(roll[r0:r0 + B * n, :t], ids[:, :t].repeat_interleave(n, 0))
    close(f"pg/{cfg_name}/roll_logits", out["roll_logits"], ref["roll_logits"])
    close(f"pg/{cfg_name}/Q", out["Q"], ref["Q"])
    close(f"pg/{cfg_name}/loss", out["pg_loss"], ref["loss"])
    fg = inst._flat_g
    for k, p in inst.gen.named_parameters():
        if k in ref["g_grads"] and not k.startswith("encoder."):
            close(f"pg/{cfg_name}/g_grads/{k}", fg.g(p), ref["g_grads"][k], rtol=2e-3, atol=1e-9)


def test_pg_loss_baseline_and_backward():
    L = L_()
    B, Lc, V = 6, 5, 41
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn(B, Lc, V, generator=g) * 2).requires_grad_(True)
    ids = torch.randint(0, V, (B, Lc), generator=g)
    Q = torch.rand(B, Lc, generator=g)
    d = torch.device("cuda:0")
    lg_d, ids_d, Q_d = logits.detach().to(d), ids.to(d), Q.to(d)      # keep the device copies alive across the call
    for mode in (0, 1):
        loss, logp = rx.pg_loss(logits, ids, Q, mode)
        (gl,) = torch.autograd.grad(loss, logits)
        lo = torch.zeros(1, device=d); dl = torch.zeros(B, Lc, V, device=d); lp = torch.zeros(B, Lc, device=d)
        L.check(L.lib().gic_pg_loss_fwd_bwd(L.ptr(lg_d), L.ptr(ids_d), L.ptr(Q_d), mode, B, Lc, V,
                                            L.ptr(lo), L.ptr(dl), L.ptr(lp), L.stream()), "gic_pg_loss_fwd_bwd")
        close(f"pg_loss/mode{mode}/loss", lo[0], loss.detach(), rtol=1e-5, atol=1e-7)
        close(f"pg_loss/mode{mode}/dlogits", dl, gl, rtol=1e-4, atol=1e-8)
        close(f"pg_loss/mode{mode}/logp", lp, logp.detach(), rtol=1e-5, atol=1e-6)


def test_rollout_rewards():
    L = L_()
    B, Lc, n, R = 5, 6, 4, 8
    g = torch.Generator().manual_seed(9)
    rl = torch.randn((Lc - 1) * B * n * R, generator=g) * 2
    ml = torch.randn(B * R, generator=g) * 2
    want = rx.rollout_q(rl, ml, B, Lc, n, R)
    d = torch.device("cuda:0")
    Q = torch.zeros(B, Lc, device=d)
    rl_d, ml_d = rl.to(d), ml.to(d)
    L.check(L.lib().gic_rollout_rewards(L.ptr(rl_d), L.ptr(ml_d), B, Lc, n, R, L.ptr(Q), L.stream()), "rewards")
    close("rollout_rewards", Q, want, rtol=1e-5)


def test_pg_step_full_update_and_properties_c3_shape():
    """A reduced BASELINE.json configs[2] shape (rollouts x prefixes batched) through the whole PG step with updates:
    structural invariants that do not depend on the size."""
    cfg = dict(B=16, L=12, V=2000, E=64, H=128, layers=1, feat=0, filters=[300, 300, 300])
    inp = rp.make_inputs(cfg)
    inst = instructor(inp)
    n = 4
    w0 = inst.gen.decoder.linear.weight.detach().clone()
    d0 = inst.disc.highway.weight.detach().clone()
    out = inst.pg_step(inp["captions"], n_roll=n)
    torch.cuda.synchronize()
    B, Lc = inp["captions"].shape
    ids, roll, Q = out["ids"].cpu(), out["roll_ids"].cpu(), out["Q"].cpu()
    assert roll.shape == ((Lc - 1) * B * n, Lc)
    assert int(ids.min()) >= 0 and int(ids.max()) < cfg["V"] and int(roll.min()) >= 0 and int(roll.max()) < cfg["V"]
    for t in range(1, Lc):
        r0 = (t - 1) * B * n
        assert torch.equal(roll[r0:r0 + B * n, :t], ids[:, :t].repeat_interleave(n, 0))
    assert float(Q.min()) > 0.0 and float(Q.max()) < 1.0
    assert torch.isfinite(out["pg_loss"]).item() and torch.isfinite(out["d_loss"]).item()
    assert not torch.equal(inst.gen.decoder.linear.weight.detach(), w0)       # both networks stepped
    assert not torch.equal(inst.disc.highway.weight.detach(), d0)
    # log pi of the sampled tokens is a log-probability
    assert float(out["logp"].max()) <= 0.0
