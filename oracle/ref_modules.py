"""CPU baseline port, TEST/BENCH INFRASTRUCTURE ONLY (see oracle/ref_port.py header).

Same algorithm as oracle/ref_port.py but expressed with the torch *library* modules the reference itself
calls on this path (nn.LSTM stepped one token at a time, nn.Conv2d + F.max_pool2d, nn.Dropout,
F.one_hot, BCEWithLogitsLoss, clip_grad_norm_, optim.Adam), so that timing it on the host cores is a
faithful stand-in for "the reference's own PyTorch CPU path" (cpu_baseline.kind = "port"): oneDNN RNN /
conv and MKL GEMM are what the reference reaches on CPU.  Used only by
bench.py's cpu_baseline / --impl reference legs and by tests that cross-check it against ref_port.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class PortDecoder(nn.Module):
    """src/generator.py:27-96 (embed, lstm, linear; sample loop with Gumbel-softmax)."""

    def __init__(self, a):
        super().__init__()
        self.embed = nn.Embedding(a.vocab_size, a.gen_embed_dim)
        self.lstm = nn.LSTM(a.gen_embed_dim, a.gen_hidden_dim, a.gen_num_layers, batch_first=True)
        self.linear = nn.Linear(a.gen_hidden_dim, a.vocab_size)
        self.temperature = float(a.temperature)

    def sample(self, features, L, u=None):
        x, state, outs, ids = features.unsqueeze(1), None, [], []
        for t in range(L):
            h, state = self.lstm(x, state)
            logits = self.linear(h.squeeze(1))
            uu = torch.rand_like(logits) if u is None else u[t]
            g = -torch.log(-torch.log(uu + 1e-10) + 1e-10)
            p = F.softmax((logits + g) * self.temperature, dim=-1)
            outs.append(p)
            tok = p.max(1)[1]
            ids.append(tok)
            x = self.embed(tok.detach()).unsqueeze(1)
        return torch.stack(outs, 1), torch.stack(ids, 1)


class PortEncoder(nn.Module):
    """Encoder.linear + Encoder.bn on pooled CNN features (src/generator.py:15-16,23-24; the frozen ResNet trunk is out of
    the path: the configs feed pooled features)."""

    def __init__(self, a):
        super().__init__()
        self.linear = nn.Linear(a.feature_dim, a.gen_embed_dim)
        self.bn = nn.BatchNorm1d(a.gen_embed_dim, momentum=0.01)

    def forward(self, pooled):
        return self.bn(self.linear(pooled))


class PortDiscriminator(nn.Module):
    """src/discriminator.py:9-62."""

    def __init__(self, a, dropout=0.2):
        super().__init__()
        es = a.disc_embed_dim // a.disc_num_rep
        self.fd = sum(a.disc_num_filters)
        self.embeddings = nn.Linear(a.vocab_size, a.disc_embed_dim, bias=False)
        self.convs = nn.ModuleList([nn.Conv2d(1, n, (f, es), stride=(1, es))
                                    for n, f in zip(a.disc_num_filters, a.disc_filter_sizes)])
        self.highway = nn.Linear(self.fd, self.fd)
        self.feature2out = nn.Linear(self.fd, 100)
        self.out2logits = nn.Linear(100, 1)
        self.dropout = nn.Dropout(dropout)

    def forward(self, inp, keep=None):
        emb = self.embeddings(inp).unsqueeze(1)
        pools = [F.max_pool2d(F.relu(c(emb)), (emb.size(2) - c.kernel_size[0] + 1, 1)).squeeze(2) for c in self.convs]
        x = torch.cat(pools, 1).permute(0, 2, 1).contiguous().view(-1, self.fd)
        hw = self.highway(x)
        s = torch.sigmoid(hw)
        y = s * F.relu(hw) + (1.0 - s) * x
        y = self.dropout(y) if keep is None else y * keep / (1.0 - self.dropout.p)
        return self.out2logits(self.feature2out(y)).squeeze(1)


def load_port(a, gen_params, disc_params, with_encoder=False):
    dec, disc = PortDecoder(a), PortDiscriminator(a)
    dec.load_state_dict({k[len("decoder."):]: v.clone() for k, v in gen_params.items() if k.startswith("decoder.")})
    disc.load_state_dict({k: v.clone() for k, v in disc_params.items()})
    if with_encoder:
        enc = PortEncoder(a)
        enc.load_state_dict({k[len("encoder."):]: v.clone() for k, v in gen_params.items() if k.startswith("encoder.")}, strict=False)
        return enc.train(), dec.train(), disc.train()
    return dec.train(), disc.train()


def port_adv_step(a, dec, disc, g_opt, d_opt, captions, feats, u=None, keep=None, enc=None, pooled=None):
    """src/training.py:144-169 + optimize (:194-199) with the Q1 ordering fix.  enc + pooled: conditional mode, the features
    come from Encoder.linear + bn (:144-145) and the encoder's parameters are trained with the generator's."""
    bce = nn.BCEWithLogitsLoss()
    B, L = captions.shape
    if enc is not None:
        feats = enc(pooled)
    gen_caps, _ = dec.sample(feats, L, u)
    fake = gen_caps.detach()
    real = F.one_hot(captions, a.vocab_size).float()
    k = keep if keep is not None else [None, None, None]
    d_real, d_fake, g_out = disc(real, k[0]), disc(fake, k[1]), disc(gen_caps, k[2])
    d_loss = bce(d_real, torch.ones_like(d_real)) + bce(d_fake, torch.zeros_like(d_fake))
    g_loss = bce(g_out, torch.ones_like(g_out))
    dps, gps = list(disc.parameters()), list(dec.parameters()) + (list(enc.parameters()) if enc is not None else [])
    dg = torch.autograd.grad(d_loss, dps, retain_graph=True)
    gg = torch.autograd.grad(g_loss, gps, allow_unused=True)
    for p, g in zip(dps, dg):
        p.grad = g
    for p, g in zip(gps, gg):
        p.grad = g
    torch.nn.utils.clip_grad_norm_(dps, a.clip_norm)
    d_opt.step()
    torch.nn.utils.clip_grad_norm_(gps, a.clip_norm)
    g_opt.step()
    return float(g_loss), float(d_loss)
