"""CPU-side checks: the C-ABI library loads and exports every declared symbol, host logic (flat buffers,
argument surface, temperature schedule) works, and the product path fails loudly without CUDA."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    import gic_b200
    return gic_b200


def test_library_exports_every_header_symbol(built):
    from gic_b200 import _lib
    L = _lib.lib()
    syms = _lib.header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), s
    assert L.gic_version() >= 100
    # size helpers are pure host functions
    assert L.gic_decode_saved_floats(4, 8, 16, 32, 1) > 0
    assert L.gic_disc_saved_floats(4, 8, 64, 64, 72) > 0


def test_sass_is_sm100a_only(built):
    from gic_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback(built):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gic_b200 import _lib
    from gic_b200.utils import get_losses
    x = torch.zeros(4)
    with pytest.raises(_lib.GicError):
        get_losses(x, x, x, "standard")
    import gic_b200.discriminator as D
    from gic_b200.args import default_args
    disc = D.Discriminator(default_args(vocab_size=20, disc_num_filters=[4, 4, 4]))
    with pytest.raises(_lib.GicError):
        disc(torch.rand(2, 8, 20))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gan-image-captioning_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("oracle/", ""), f


def test_args_surface_matches_reference_defaults(built):
    from gic_b200.args import default_args, get_args
    a = get_args([])
    assert (a.gen_hidden_dim, a.gen_embed_dim, a.gen_num_layers) == (512, 32, 1)
    assert (a.disc_embed_dim, a.disc_num_rep) == (64, 64)
    assert a.disc_filter_sizes == [3, 4, 5] and a.disc_num_filters == [300, 300, 300]
    assert a.temperature == 100 and a.temp_adpt == "exp" and a.clip_norm == 5.0
    assert a.adv_loss_type == "standard" and a.gen_lr == 1e-4 and a.disc_lr == 1e-4
    b = get_args(["--gen-embed-dim", "512", "--gen-num-layers", "2", "--disc-filter-sizes", "2,3,4"])
    assert b.gen_embed_dim == 512 and b.gen_num_layers == 2 and b.disc_filter_sizes == [2, 3, 4]
    with pytest.raises(AttributeError):
        default_args(nope=1)


def test_state_dict_keys_match_reference(built):
    import gic_b200.discriminator as D
    import gic_b200.generator as G
    from gic_b200.args import default_args
    from oracle import ref_port as rp
    a = default_args(vocab_size=40, gen_num_layers=2, disc_num_filters=[5, 6, 7], conditional_gan=1)
    gen, disc = G.Generator(a), D.Discriminator(a)
    gk = set(gen.state_dict())
    for k, shape in rp.gen_param_shapes(a):
        assert k in gk and tuple(gen.state_dict()[k].shape) == shape, k
    assert {k for k, _ in rp.disc_param_shapes(a)} == set(disc.state_dict())
    for p in list(gen.parameters()) + list(disc.parameters()):
        if p.dim() > 0:
            assert float(p.detach().min()) >= -0.05 - 1e-7 and float(p.detach().max()) <= 0.05 + 1e-7   # init_params, Q6 (fp32(0.05) rounds outward)


def test_temperature_schedule_matches_oracle(built):
    from gic_b200.utils import get_fixed_temperature
    from oracle import ref_port as rp
    for adapt in ("no", "lin", "exp", "log", "sigmoid", "quad", "sqrt"):
        for i in (0, 1, 7.5, 29):
            assert get_fixed_temperature(100, i, 30, adapt) == rp.get_fixed_temperature(100, i, 30, adapt)
    with pytest.raises(Exception):
        get_fixed_temperature(100, 1, 30, "bogus")


def test_flat_params_views(built):
    from gic_b200.training import FlatParams
    ps = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2, 2))]
    vals = [p.detach().clone() for p in ps]
    fp = FlatParams(ps, "cpu")
    assert fp.homed() and fp.n % 4 == 0
    for p, v in zip(ps, vals):
        assert torch.equal(p.detach(), v)
    fp.flat.mul_(2.0)
    for p, v in zip(ps, vals):
        assert torch.equal(p.detach(), 2 * v)          # params are views of the flat buffer
    fp.g(ps[1]).fill_(1.0)
    assert float(fp.grad.sum()) == 7.0
