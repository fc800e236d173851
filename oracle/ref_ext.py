"""CPU oracle for the north-star EXTENSIONS of the hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED BY REFERENCE: kawshik8/GAN-Image-Captioning contains no inverse-CDF sampler, no Monte-Carlo
rollouts, no reward/baseline and no policy-gradient loss (its
generator update is RL-free, src/training.py:150-169).  BASELINE.json's north_star asks for them, so they are
defined HERE first (plain PyTorch fp32 on the CPU, composed from the reference-pinned pieces of
``oracle/ref_port.py``: the same LSTM cell, vocab projection and discriminator) and the CUDA path is tested
against this file.  Only ``tests/`` may import it.

Definitions (mirrored in include/gic_b200.h):
  * categorical sampling by inverse CDF: token = first index whose cumulative softmax probability exceeds u;
  * rollouts: for every prefix length t = 1..L-1 of a sampled caption, n continuations with the same policy;
    row (t-1)*B*n + b*n + j, uniforms u_roll[L, Mmax];
  * reward = mean over the R representations of sigmoid(D logit) (D in eval mode: no dropout);
    Q[b,t-1] = mean reward of the rollouts of prefix t; Q[b,L-1] = reward of the caption itself;
  * loss = -(1/(B L)) sum log pi(y_bt) (Q_bt - base_t), base_t = mean_b Q_bt (baseline_mode 1) or 0.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ref_port as rp


def decode_cdf(p, features, u, L, forced_ids=None):
    """Single-layer Decoder.sample loop (src/generator.py:55-81) with inverse-CDF sampling from u[L,B].
    Returns logits[B,L,V], ids[B,L], logp[B,L], hs[L+1][B,H], cs[L+1][B,H]."""
    B = features.shape[0]
    H = p["decoder.lstm.weight_hh_l0"].shape[1]
    h = features.new_zeros(B, H)
    c = features.new_zeros(B, H)
    hs, cs = [h], [c]
    x = features
    logits_l, ids, logp = [], [], []
    for t in range(L):
        h, c = rp.lstm_cell(x, h, c, p["decoder.lstm.weight_ih_l0"], p["decoder.lstm.weight_hh_l0"],
                            p["decoder.lstm.bias_ih_l0"], p["decoder.lstm.bias_hh_l0"])
        hs.append(h); cs.append(c)
        logits = F.linear(h, p["decoder.linear.weight"], p["decoder.linear.bias"])
        tok = rp.sample_inverse_cdf(logits, u[t])
        logits_l.append(logits); ids.append(tok)
        logp.append(F.log_softmax(logits, -1).gather(1, tok[:, None])[:, 0])
        fed = tok if forced_ids is None else forced_ids[:, t]
        x = p["decoder.embed.weight"][fed]
    return torch.stack(logits_l, 1), torch.stack(ids, 1), torch.stack(logp, 1), hs, cs


def cdf_boundary_gap(logits, u_row, tok_a, tok_b):
    """|u - CDF boundary| for rows where two samplers disagree (a draw within 1e-6 of a boundary is a tie)."""
    cdf = torch.cumsum(F.softmax(logits.double(), -1), -1)
    lo = torch.minimum(tok_a, tok_b)
    return (cdf.gather(1, lo[:, None])[:, 0] - u_row.double()).abs()


def rollouts(p, main_ids, hs, cs, u_roll, n):
    """roll_ids[Mmax, L]; row (t-1)*B*n + b*n + j."""
    B, L = main_ids.shape
    G = B * n
    Mmax = (L - 1) * G
    roll = torch.zeros(Mmax, L, dtype=torch.long)
    H = hs[0].shape[1]
    h = torch.zeros(Mmax, H); c = torch.zeros(Mmax, H)
    x = torch.zeros(Mmax, p["decoder.embed.weight"].shape[1])
    for t in range(1, L):
        r0 = (t - 1) * G
        h[r0:r0 + G] = hs[t].repeat_interleave(n, 0)
        c[r0:r0 + G] = cs[t].repeat_interleave(n, 0)
        x[r0:r0 + G] = p["decoder.embed.weight"][main_ids[:, t - 1]].repeat_interleave(n, 0)
        roll[r0:r0 + G, :t] = main_ids[:, :t].repeat_interleave(n, 0)
        M = t * G
        hn, cn = rp.lstm_cell(x[:M], h[:M], c[:M], p["decoder.lstm.weight_ih_l0"], p["decoder.lstm.weight_hh_l0"],
                              p["decoder.lstm.bias_ih_l0"], p["decoder.lstm.bias_hh_l0"])
        h[:M], c[:M] = hn, cn
        logits = F.linear(hn, p["decoder.linear.weight"], p["decoder.linear.bias"])
        tok = rp.sample_inverse_cdf(logits, u_roll[t, :M])
        roll[:M, t] = tok
        x[:M] = p["decoder.embed.weight"][tok]
    return roll


def rollout_q(roll_logits, main_logits, B, L, n, R):
    Q = torch.zeros(B, L)
    if L > 1:
        r = torch.sigmoid(roll_logits).view(L - 1, B, n * R).mean(-1)      # [L-1, B]
        Q[:, :L - 1] = r.t()
    Q[:, L - 1] = torch.sigmoid(main_logits).view(B, R).mean(-1)
    return Q


def pg_loss(logits, ids, Q, baseline_mode=1):
    B, L, V = logits.shape
    logp = F.log_softmax(logits, -1).gather(2, ids[:, :, None])[:, :, 0]
    base = Q.mean(0, keepdim=True) if baseline_mode == 1 else torch.zeros(1, L)
    return -(logp * (Q - base)).sum() / (B * L), logp


def pg_step(inp, u_main, u_roll, n, baseline_mode=1):
    """Sampled captions, rollouts, rewards, loss and d loss / d theta_G (teacher-forced on the sampled ids)."""
    a = inp["args"]
    gp, dp = inp["gen"], inp["disc"]
    B, L = inp["captions"].shape
    feats = rp.encoder_project(gp, inp["pooled"]) if a.conditional_gan else rp.start_features(gp, B)
    logits, ids, logp, hs, cs = decode_cdf(gp, feats, u_main, L)
    roll = rollouts(gp, ids, hs, cs, u_roll, n)
    R = a.disc_num_rep
    roll_logits = rp.disc_forward_ids(dp, roll, None, a.disc_filter_sizes)
    main_logits = rp.disc_forward_ids(dp, ids, None, a.disc_filter_sizes)
    Q = rollout_q(roll_logits, main_logits, B, L, n, R)
    # gradient: autograd through the teacher-forced decode
    gpar = {k: v.clone().requires_grad_(True) for k, v in gp.items()}
    feats_g = rp.encoder_project(gpar, inp["pooled"]) if a.conditional_gan else rp.start_features(gpar, B)
    logits_g, _, _, _, _ = decode_cdf(gpar, feats_g, u_main, L, forced_ids=ids)
    loss, _ = pg_loss(logits_g, ids, Q, baseline_mode)
    names = [k for k in gpar]
    grads = torch.autograd.grad(loss, [gpar[k] for k in names], allow_unused=True)
    g_grads = {k: g for k, g in zip(names, grads) if g is not None}
    return dict(logits=logits, ids=ids, logp=logp, roll_ids=roll, roll_logits=roll_logits, main_logits=main_logits, Q=Q,
                loss=loss.detach(), g_grads=g_grads, features=feats)


# --------------------------------------------------------------------------------------------------------------
# B1 (EXTENSION, parity unpinned by reference): additive attention over the CNN feature grid.
# The reference feeds ONE pooled vector at step 0 and has no attention (src/generator.py:19-25,58); its report
# describes grid cross-attention only for a Transformer variant whose code is not in the repository.  north_star
# asks for "the per-step RNN/attention cell over the CNN image-feature grid", so the cell is defined here:
#   once per image:  Ak = grid W_k^T [B,P,Da],  Av = grid W_v^T [B,P,E]            (P = 49 locations, 2048 channels)
#   every step t:    q = h_{t-1} W_q^T;  s_l = w_e . tanh(Ak_l + q);  alpha = softmax_l(s);  ctx = sum_l alpha_l Av_l
#                    LSTM input x'_t = x_t + ctx   (x_0 = projected pooled feature, x_t = embed(token_{t-1}))
# Everything after the LSTM input (cell, vocab projection, Gumbel-softmax sampling) is the reference-pinned path.
# --------------------------------------------------------------------------------------------------------------
ATTN_KEYS = ("decoder.attn_k.weight", "decoder.attn_v.weight", "decoder.attn_q.weight", "decoder.attn_e.weight")


def attn_param_shapes(a, Cf, Da):
    return [("decoder.attn_k.weight", (Da, Cf)), ("decoder.attn_v.weight", (a.gen_embed_dim, Cf)),
            ("decoder.attn_q.weight", (Da, a.gen_hidden_dim)), ("decoder.attn_e.weight", (Da,))]


def attention_context(p, Ak, Av, h_prev):
    q = F.linear(h_prev, p["decoder.attn_q.weight"])                            # [B,Da]
    s = (torch.tanh(Ak + q[:, None, :]) * p["decoder.attn_e.weight"].reshape(-1)).sum(-1)    # [B,P]
    alpha = F.softmax(s, dim=-1)
    return (alpha[:, :, None] * Av).sum(1), alpha


def decoder_sample_attn(p, features, grid, u, temperature, L, forced_ids=None):
    """Single-layer Decoder.sample with the attention cell.  Returns probs[B,L,V], ids[B,L], alphas[B,L,P]."""
    B = features.shape[0]
    H = p["decoder.lstm.weight_hh_l0"].shape[1]
    Ak = F.linear(grid, p["decoder.attn_k.weight"])
    Av = F.linear(grid, p["decoder.attn_v.weight"])
    h = features.new_zeros(B, H)
    c = features.new_zeros(B, H)
    x = features
    outs, ids, alphas = [], [], []
    for t in range(L):
        ctx, alpha = attention_context(p, Ak, Av, h)
        h, c = rp.lstm_cell(x + ctx, h, c, p["decoder.lstm.weight_ih_l0"], p["decoder.lstm.weight_hh_l0"],
                            p["decoder.lstm.bias_ih_l0"], p["decoder.lstm.bias_hh_l0"])
        logits = F.linear(h, p["decoder.linear.weight"], p["decoder.linear.bias"])
        pred = F.softmax((logits + rp.gumbel_noise(u[t])) * temperature, dim=-1)
        tok = pred.max(1)[1]
        outs.append(pred); ids.append(tok); alphas.append(alpha)
        fed = tok if forced_ids is None else forced_ids[:, t]
        x = p["decoder.embed.weight"][fed.detach()]
    return torch.stack(outs, 1), torch.stack(ids, 1), torch.stack(alphas, 1)
