"""RelGAN-style multi-representation CNN discriminator with the reference's interface
(src/discriminator.py:9-86): same constructor, parameter names/shapes and ``forward(inp[B,L,V]) ->
logits[B*num_rep]``.  Compute runs in libgic_b200.so (gic_disc_fwd / gic_disc_bwd)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

import gic_b200
from . import _lib


def _f32c(t):
    return t.detach().contiguous().float()


def disc_weights(disc):
    """Flat tuple of the discriminator's tensors in the order the C ABI takes them."""
    cw = [c.weight for c in disc.convs]
    cb = [c.bias for c in disc.convs]
    return (disc.embeddings.weight, disc.highway.weight, disc.highway.bias, disc.feature2out.weight,
            disc.feature2out.bias, disc.out2logits.weight, disc.out2logits.bias, *cw, *cb)


def disc_fwd_raw(lib, mode, inp_soft, ids, N, L, V, De, R, fsz, nfl, W_e, cw, cb, W_h, b_h, W_f, b_f, W_o, b_o, keeps,
                 drop_p, dev):
    """Thin call of gic_disc_fwd: returns (list of logits per head, saved)."""
    Fd = sum(nfl)
    logits = [torch.empty(N * R, device=dev) for _ in keeps]
    saved = torch.empty(lib.gic_disc_saved_floats(N, L, De, R, Fd), device=dev)
    ws = torch.empty(lib.gic_disc_fwd_workspace_floats(Fd), device=dev)
    _lib.check(lib.gic_disc_fwd(mode, _lib.ptr(inp_soft), _lib.ptr(ids), N, L, V, De, R, len(fsz),
                                _lib.int_array(fsz), _lib.int_array(nfl), _lib.ptr(W_e), _lib.ptr_array(cw),
                                _lib.ptr_array(cb), _lib.ptr(W_h), _lib.ptr(b_h), _lib.ptr(W_f), _lib.ptr(b_f),
                                W_f.shape[0], _lib.ptr(W_o), _lib.ptr(b_o), len(keeps), _lib.ptr_array(keeps),
                                float(drop_p), _lib.ptr_array(logits), _lib.ptr(saved), _lib.ptr(ws), _lib.stream()),
               "gic_disc_fwd")
    return logits, saved


class _DiscForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, ids, keep, drop_p, mode, fsz, nfl, R, W_e, W_h, b_h, W_f, b_f, W_o, b_o, *convs):
        _lib.require_cuda()
        lib = _lib.lib()
        ng = len(fsz)
        W_e, W_h, b_h, W_f, b_f, W_o, b_o = map(_f32c, (W_e, W_h, b_h, W_f, b_f, W_o, b_o))
        cw = [_f32c(c) for c in convs[:ng]]
        cb = [_f32c(c) for c in convs[ng:]]
        De, V = W_e.shape
        if inp is not None:
            inp = _f32c(inp)
            N, L, Vin = inp.shape
            if Vin != V:
                raise ValueError("discriminator input has vocab %d, expected %d" % (Vin, V))
            dev = inp.device
        else:
            ids = ids.detach().contiguous().long()
            N, L = ids.shape
            dev = ids.device
        if keep is not None:
            keep = keep.detach().to(torch.uint8).contiguous()
        logits, saved = disc_fwd_raw(lib, mode, inp, ids, N, L, V, De, R, fsz, nfl, W_e, cw, cb, W_h, b_h, W_f, b_f,
                                     W_o, b_o, [keep], drop_p, dev)
        ctx.save_for_backward(saved, inp, ids, keep, W_e, W_h, b_h, W_f, b_f, W_o, b_o, *cw, *cb)
        ctx.meta = (N, L, V, De, R, list(fsz), list(nfl), float(drop_p), mode)
        return logits[0]

    @staticmethod
    def backward(ctx, dlogits):
        lib = _lib.lib()
        N, L, V, De, R, fsz, nfl, drop_p, mode = ctx.meta
        ng = len(fsz)
        saved, inp, ids, keep, W_e, W_h, b_h, W_f, b_f, W_o, b_o = ctx.saved_tensors[:11]
        cw = list(ctx.saved_tensors[11:11 + ng])
        cb = list(ctx.saved_tensors[11 + ng:11 + 2 * ng])
        dev = saved.device
        Fd = sum(nfl)
        dlogits = _f32c(dlogits)
        ws = torch.empty(lib.gic_disc_bwd_workspace_floats(N, L, De, R, Fd), device=dev)
        want_inp = inp is not None and ctx.needs_input_grad[0]
        want_param = any(ctx.needs_input_grad[8:])
        dW_e, dW_h, db_h = torch.empty_like(W_e), torch.empty_like(W_h), torch.empty_like(b_h)
        dW_f, db_f, dW_o, db_o = torch.empty_like(W_f), torch.empty_like(b_f), torch.empty_like(W_o), torch.empty_like(b_o)
        dcw = [torch.empty_like(c) for c in cw]
        dcb = [torch.empty_like(c) for c in cb]
        dinp = torch.empty(N, L, V, device=dev) if want_inp else None
        _lib.check(lib.gic_disc_bwd(mode, _lib.ptr(dlogits), _lib.ptr(keep), drop_p, _lib.ptr(inp), _lib.ptr(ids), N, L,
                                    V, De, R, ng, _lib.int_array(fsz), _lib.int_array(nfl), _lib.ptr(W_e),
                                    _lib.ptr_array(cw), _lib.ptr_array(cb), _lib.ptr(W_h), _lib.ptr(W_f), _lib.ptr(b_f),
                                    W_f.shape[0], _lib.ptr(W_o), _lib.ptr(b_o), _lib.ptr(saved), _lib.ptr(ws),
                                    _lib.ptr(dW_e), _lib.ptr_array(dcw), _lib.ptr_array(dcb), _lib.ptr(dW_h),
                                    _lib.ptr(db_h), _lib.ptr(dW_f), _lib.ptr(db_f), _lib.ptr(dW_o), _lib.ptr(db_o),
                                    _lib.ptr(dinp), int(want_param), 0, _lib.stream()), "gic_disc_bwd")
        if not want_param:
            return (dinp,) + (None,) * (14 + 2 * ng)
        return (dinp, None, None, None, None, None, None, None, dW_e, dW_h, db_h, dW_f, db_f, dW_o, db_o, *dcw, *dcb)


class Discriminator(nn.Module):
    def __init__(self, args, gpu=False, dropout=0.2):
        super().__init__()
        self.vocab_size = args.vocab_size
        self.embed_dim = args.disc_embed_dim
        self.padding_idx = args.padding_idx
        self.feature_dim = sum(args.disc_num_filters)
        self.emb_dim_single = int(args.disc_embed_dim / args.disc_num_rep)
        self.num_rep = args.disc_num_rep
        self.filter_sizes = list(args.disc_filter_sizes)
        self.num_filters = list(args.disc_num_filters)
        self.gpu = gpu
        self.embeddings = nn.Linear(self.vocab_size, self.embed_dim, bias=False)
        self.convs = nn.ModuleList([
            nn.Conv2d(1, n, (f, self.emb_dim_single), stride=(1, self.emb_dim_single))
            for (n, f) in zip(self.num_filters, self.filter_sizes)])
        self.highway = nn.Linear(self.feature_dim, self.feature_dim)
        self.feature2out = nn.Linear(self.feature_dim, 100)
        self.out2logits = nn.Linear(100, 1)
        self.dropout = nn.Dropout(dropout)
        self.args = args
        self.init_params()

    def _run(self, inp, ids, keep):
        N = (inp if inp is not None else ids).shape[0]
        if keep is None and self.training and self.dropout.p > 0:
            # same behaviour as nn.Dropout in train mode: a fresh Bernoulli(1-p) keep mask per call
            dev = (inp if inp is not None else ids).device
            keep = torch.rand(N * self.num_rep, self.feature_dim, device=dev) >= self.dropout.p
        w = disc_weights(self)
        ng = len(self.filter_sizes)
        return _DiscForward.apply(inp, ids, keep, self.dropout.p, gic_b200.get_gemm_mode(), self.filter_sizes,
                                  self.num_filters, self.num_rep, *w[:7], *w[7:7 + ng], *w[7 + ng:])

    def forward(self, inp, keep=None):
        """Get logits of discriminator.  inp: batch_size * seq_len * vocab_size (soft or one-hot) ->
        logits [batch_size * num_rep].  ``keep`` optionally supplies the dropout keep-mask [B*R, F]."""
        return self._run(inp, None, keep)

    def forward_ids(self, ids, keep=None):
        """Hard-token path: equals forward(F.one_hot(ids, V).float()) without materialising the one-hot
        (src/training.py:158)."""
        return self._run(None, ids, keep)

    def get_feature(self, inp):
        raise NotImplementedError("Discriminator.get_feature is never called by the reference and is shape-invalid "
                                  "for num_rep > 1 (SURVEY.md section 2 row 4)")

    def init_params(self):
        for param in self.parameters():
            if param.requires_grad and len(param.shape) > 0:
                stddev = 1 / math.sqrt(param.shape[0])
                if self.args.disc_init == "uniform":
                    torch.nn.init.uniform_(param, a=-0.05, b=0.05)
                elif self.args.disc_init == "normal":
                    torch.nn.init.normal_(param, std=stddev)
