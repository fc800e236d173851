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


# ------------------------------------------------------------------------------------------------------------
# B1: attention cell over the CNN feature grid (definition: oracle/ref_ext.py; parity unpinned by reference)
# ------------------------------------------------------------------------------------------------------------
def _attn_setup(B=4, Lc=6, V=40, E=16, H=32, Pn=9, Cf=24, Da=12, seed=5):
    from gic_b200.args import default_args
    import gic_b200.generator as G
    a = default_args(vocab_size=V, gen_embed_dim=E, gen_hidden_dim=H, gen_attention=1, attn_dim=Da, feature_channels=Cf,
                     device="cuda")
    torch.manual_seed(seed)
    gen = G.Generator(a).cuda()
    # larger attention weights than the U(-0.05, 0.05) init so that alpha is far from uniform
    with torch.no_grad():
        for p in gen.decoder.attn_params():
            p.mul_(12.0)
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, E, generator=g) * 0.5
    grid = torch.randn(B, Pn, Cf, generator=g)
    u = torch.rand(Lc, B, V, generator=g)
    return a, gen, feats, grid, u


@pytest.mark.parametrize("T", [1.0, 4.0])
def test_attention_decode_forward_backward_vs_oracle(T):
    a, gen, feats, grid, u = _attn_setup()
    Lc = u.shape[0]
    p_cpu = {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()}
    ref_p, ref_ids, ref_alpha = rx.decoder_sample_attn(p_cpu, feats, grid, u, T, Lc)
    gen.decoder.temperature = T
    fd, gd = feats.cuda().requires_grad_(True), grid.cuda()
    out, ids = gen.decoder.sample(fd, max_caption_len=Lc, u=u.cuda(), forced_ids=ref_ids.cuda(), grid=gd)
    torch.cuda.synchronize()
    assert torch.equal(ids.cpu(), ref_ids)
    close(f"attn/T{T}/probs", out, ref_p)
    # backward of a fixed linear functional of the soft captions: autograd on the oracle vs the CUDA backward
    g = torch.Generator().manual_seed(99)
    Rw = torch.randn(out.shape, generator=g)
    pr = {k: v.clone().requires_grad_(True) for k, v in p_cpu.items() if k.startswith("decoder.")}
    f_cpu = feats.clone().requires_grad_(True)
    pp, _, _ = rx.decoder_sample_attn(pr, f_cpu, grid, u, T, Lc, forced_ids=ref_ids)
    loss = (pp * Rw).sum()
    names = sorted(pr)
    grads = torch.autograd.grad(loss, [pr[k] for k in names] + [f_cpu], allow_unused=True)
    (out * Rw.cuda()).sum().backward()
    torch.cuda.synchronize()
    sd = dict(gen.named_parameters())
    for k, gref in zip(names, grads[:-1]):
        if gref is None:
            continue
        got = sd[k].grad
        assert got is not None, k
        close(f"attn/T{T}/grad/{k}", got.reshape(-1), gref.reshape(-1), rtol=2e-3, atol=1e-7)
    close(f"attn/T{T}/grad/features", fd.grad, grads[-1], rtol=2e-3, atol=1e-7)


def test_attention_weights_are_a_distribution():
    """Property at a larger shape (P = 49 locations, 2048 channels): the saved alphas sum to 1 per (step, caption)."""
    import ctypes as C
    L = L_()
    a, gen, feats, grid, u = _attn_setup(B=8, Lc=5, V=200, E=32, H=64, Pn=49, Cf=2048, Da=64, seed=7)
    lib = L.lib()
    d = torch.device("cuda:0")
    B, E = feats.shape
    Lc, V, H, Pn, Da = u.shape[0], 200, 64, 49, 64
    dec = gen.decoder
    saved = torch.empty(lib.gic_decode_saved_floats(B, Lc, E, H, 1), device=d)
    asaved = torch.zeros(lib.gic_attn_saved_floats(B, Lc, Pn, Da, E), device=d)
    ws = torch.empty(lib.gic_decode_fwd_workspace_floats(B, V, H), device=d)
    out = torch.empty(B, Lc, V, device=d); ids = torch.empty(B, Lc, dtype=torch.int64, device=d)
    gd, fd, ud = grid.to(d), feats.to(d), u.to(d)
    we = dec.attn_e.weight.detach().reshape(-1).contiguous()
    blk = L.attn_block(gd, dec.attn_k.weight.detach(), dec.attn_v.weight.detach(), dec.attn_q.weight.detach(), we, asaved)
    lp = [p.detach() for p in dec.lstm_params()]
    L.check(lib.gic_decode_sample_fwd_attn(C.byref(blk), 0, L.ptr(fd), L.ptr(dec.embed.weight.detach()), L.ptr_array([lp[0]]),
                                           L.ptr_array([lp[1]]), L.ptr_array([lp[2]]), L.ptr_array([lp[3]]),
                                           L.ptr(dec.linear.weight.detach()), L.ptr(dec.linear.bias.detach()), L.ptr(ud), 1.0, 0,
                                           None, B, Lc, V, E, H, 1, L.ptr(out), L.ptr(ids), L.ptr(saved), L.ptr(ws), L.stream()),
            "fwd_attn")
    torch.cuda.synchronize()
    a4 = lambda x: (x + 3) & ~3
    off = a4(B * Pn * Da) + a4(B * Pn * E) + a4(Lc * B * Da)
    alpha = asaved[off:off + Lc * B * Pn].view(Lc, B, Pn).cpu()
    assert float((alpha.sum(-1) - 1).abs().max()) < 1e-5 and float(alpha.min()) >= 0.0
    assert float((out.sum(-1).cpu() - 1).abs().max()) < 1e-4


def test_adv_step_with_attention_grid_runs_and_trains_attention_params():
    """Fused adversarial step with the attention cell: the step runs, every attention parameter receives a finite,
    non-zero gradient and is updated; the fused (factored) generator backward agrees with the autograd-driven one."""
    from gic_b200.args import default_args
    from gic_b200.training import GANInstructor
    B, Lc, V = 6, 8, 60
    a = default_args(vocab_size=V, gen_embed_dim=16, gen_hidden_dim=32, gen_attention=1, attn_dim=12, feature_channels=24,
                     conditional_gan=1, feature_dim=24, disc_num_filters=[20, 24, 28], device="cuda")
    torch.manual_seed(3)
    inst = GANInstructor(a, device="cuda:0")
    inst.gen.train(); inst.disc.train(); inst.gen.decoder.temperature = 2.0
    with torch.no_grad():
        for p in inst.gen.decoder.attn_params():
            p.mul_(12.0)
    g = torch.Generator().manual_seed(4)
    caps = torch.randint(4, V, (B, Lc), generator=g)
    grid = torch.randn(B, 9, 24, generator=g)
    u = torch.rand(Lc, B, V, generator=g)
    keep = (torch.rand(3, B * 64, 72, generator=g) >= 0.2)
    before = [p.detach().clone() for p in inst.gen.decoder.attn_params()]
    out = inst.adv_step(caps, grid=grid, u=u, keep=keep, update=False)
    torch.cuda.synchronize()
    fg = inst._flat_g
    fused = {k: fg.g(p).clone() for k, p in inst.gen.named_parameters() if k.startswith("decoder.")}
    for k in ("decoder.attn_k.weight", "decoder.attn_v.weight", "decoder.attn_q.weight", "decoder.attn_e.weight"):
        assert torch.isfinite(fused[k]).all() and float(fused[k].abs().max()) > 0, k
    # autograd-driven path on the same modules (dense d(probs) through Discriminator.forward)
    from gic_b200.utils import get_losses
    inst.gen.zero_grad(); inst.disc.zero_grad()
    feats = inst.gen.encoder(grid.cuda().mean(1))
    probs, ids = inst.gen.decoder.sample(feats, max_caption_len=Lc, u=u.cuda(), grid=grid.cuda())
    assert torch.equal(ids, out["ids"])
    g_out = inst.disc(probs, keep=keep[2].cuda())
    g_loss, _ = get_losses(out["d_real"], out["d_fake"], g_out, "standard")
    g_loss.backward()
    for k, p in inst.gen.named_parameters():
        if k.startswith("decoder.") and p.grad is not None:
            close(f"attn_step/{k}", fused[k], p.grad.detach().cpu(), rtol=2e-3, atol=1e-9)
    inst.adv_step(caps, grid=grid, u=u, keep=keep)
    torch.cuda.synchronize()
    for b0, p in zip(before, inst.gen.decoder.attn_params()):
        assert not torch.equal(b0, p.detach())


# ---- library-side random draws (Philox4x32-10): production steps pass no uniforms / masks ---------------------------------
def test_philox_uniform_statistics_and_determinism():
    from gic_b200 import _lib as L
    L.require_cuda(); lib = L.lib()
    n = 4_000_003                                                   # not a multiple of 4: the tail group is clipped
    a, b, c = (torch.empty(n, device="cuda:0") for _ in range(3))
    lib.gic_set_rng(1008, 7, None); L.check(lib.gic_philox_uniform(L.RNG_TAG_GUMBEL, n, L.ptr(a), L.stream()), "u")
    lib.gic_set_rng(1008, 7, None); L.check(lib.gic_philox_uniform(L.RNG_TAG_GUMBEL, n, L.ptr(b), L.stream()), "u")
    lib.gic_set_rng(1008, 8, None); L.check(lib.gic_philox_uniform(L.RNG_TAG_GUMBEL, n, L.ptr(c), L.stream()), "u")
    lib.gic_set_rng(0, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert float(a.min()) >= 0.0 and float(a.max()) < 1.0
    assert abs(float(a.double().mean()) - 0.5) < 1e-3 and abs(float(a.double().var()) - 1 / 12) < 1e-3
    hist = torch.histc(a, bins=64, min=0.0, max=1.0)
    assert float((hist / n - 1 / 64).abs().max()) < 5e-4               # 64 equiprobable bins, 62 500 expected each
    assert abs(float(torch.corrcoef(torch.stack([a[:-1], a[1:]]))[0, 1])) < 2e-3     # neighbouring draws uncorrelated
    # the state can also live in device memory (CUDA-graph replay)
    st = torch.tensor([1008, 7], dtype=torch.int64, device="cuda:0")
    lib.gic_set_rng(0, 0, L.ptr(st)); L.check(lib.gic_philox_uniform(L.RNG_TAG_GUMBEL, n, L.ptr(c), L.stream()), "u")
    lib.gic_set_rng(0, 0, None)
    assert torch.equal(a, c)
    # keep masks: the same stream thresholded at p
    k = torch.empty(n - 3, dtype=torch.uint8, device="cuda:0")
    u = torch.empty(n - 3, device="cuda:0")
    lib.gic_set_rng(5, 1, None)
    L.check(lib.gic_philox_keep_mask(L.RNG_TAG_DROPOUT, n - 3, 0.2, L.ptr(k), L.stream()), "k")
    L.check(lib.gic_philox_uniform(L.RNG_TAG_DROPOUT, n - 3, L.ptr(u), L.stream()), "u")
    lib.gic_set_rng(0, 0, None)
    assert torch.equal(k.bool(), u >= 0.2) and abs(float(k.float().mean()) - 0.8) < 1e-3


@pytest.mark.parametrize("mode_name", ["GEMM_BF16", "GEMM_FP32"])
def test_step_with_library_draws_equals_step_with_the_same_draws_supplied(mode_name):
    """adv_step(u=None, keep=None) -- uniforms generated inside the fused decode kernel (tensor-core modes) or slice by
    slice (exact-fp32 mode), masks by gic_philox_keep_mask -- against the same step with exactly those draws passed in."""
    import gic_b200
    from gic_b200 import _lib as L
    from gic_b200.args import default_args
    from gic_b200.training import GANInstructor
    B, Lc, V = 40, 7, 1200
    a = default_args(vocab_size=V, gen_embed_dim=64, gen_hidden_dim=128, gen_num_layers=1, conditional_gan=0, device="cuda")
    old = gic_b200.get_gemm_mode()
    gic_b200.set_gemm_mode(getattr(gic_b200, mode_name))
    try:
        def build():
            torch.manual_seed(3)
            inst = GANInstructor(a, device="cuda:0")
            inst.gen.train(); inst.disc.train(); inst.gen.decoder.temperature = 2.0
            return inst
        g = torch.Generator(device="cuda:0").manual_seed(9)
        caps = torch.randint(4, V, (B, Lc), generator=g, device="cuda:0")
        lib = L.lib()
        seed, off = 424242, 17
        U = torch.empty(Lc, B, V, device="cuda:0")
        K = torch.empty(3, B * 64, 900, dtype=torch.uint8, device="cuda:0")
        lib.gic_set_rng(seed, off, None)
        L.check(lib.gic_philox_uniform(L.RNG_TAG_GUMBEL, U.numel(), L.ptr(U), L.stream()), "u")
        L.check(lib.gic_philox_keep_mask(L.RNG_TAG_DROPOUT, K.numel(), 0.2, L.ptr(K), L.stream()), "k")
        lib.gic_set_rng(0, 0, None)
        i1 = build(); i1._rng_seed, i1._rng_offset = seed, off - 1
        r1 = i1.adv_step(caps, update=False)
        ids1, p1, g1 = r1["ids"].clone(), r1["probs"].clone(), i1._flat_g.grad.clone()
        i2 = build()
        r2 = i2.adv_step(caps, u=U, keep=K, update=False)
        torch.cuda.synchronize()
        assert torch.equal(ids1, r2["ids"])
        assert torch.equal(p1, r2["probs"])                           # same uniforms -> bit-identical soft captions
        close(f"philox/{mode_name}/d_loss", r1["d_loss"], r2["d_loss"], rtol=1e-6)
        close(f"philox/{mode_name}/g_grad", g1, i2._flat_g.grad, rtol=2e-2, atol=1e-12)     # run-to-run noise of the step (atomics, pool ties)
        # the next step draws a different stream
        r3 = i1.adv_step(caps, update=False)
        assert not torch.equal(r3["ids"], ids1)
    finally:
        gic_b200.set_gemm_mode(old)


def test_graph_replay_with_library_draws_advances_the_stream():
    import gic_b200
    from gic_b200.args import default_args
    from gic_b200.training import GANInstructor
    B, Lc, V = 32, 6, 1000
    a = default_args(vocab_size=V, gen_embed_dim=32, gen_hidden_dim=64, gen_num_layers=1, conditional_gan=0, device="cuda")
    old = gic_b200.get_gemm_mode()
    gic_b200.set_gemm_mode(gic_b200.GEMM_BF16)
    try:
        torch.manual_seed(3)
        inst = GANInstructor(a, device="cuda:0")
        inst.gen.train(); inst.disc.train()
        caps = torch.randint(4, V, (B, Lc), device="cuda:0")
        seen = []
        for _ in range(4):                                          # step 1 eager + capture, steps 2-4 replayed
            r = inst.adv_step(caps, graph=True)
            torch.cuda.synchronize()
            seen.append(r["ids"].clone())
        assert not torch.equal(seen[1], seen[2]) and not torch.equal(seen[2], seen[3])
        assert all(int(s.min()) >= 0 and int(s.max()) < V for s in seen)
        assert inst._rng_offset >= 4
    finally:
        gic_b200.set_gemm_mode(old)
