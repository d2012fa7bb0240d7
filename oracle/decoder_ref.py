"""CPU ORACLE (test infrastructure, NOT product code) -- PARITY UNPINNED.

A PyTorch-CPU restatement of the reference's adaptive-attention LSTM caption
decoder, following /root/reference/ImageCaptioning/model/model_adaAttention_aic.py
op for op.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import this module; the product path
(myimagecaptioningmodel_b200/) never does.

"Parity unpinned": the reference ships no tests, golden vectors or fixtures, and its
arithmetic lives in PaddlePaddle 1.8 (README.md:15), which cannot be installed here
(no wheel for CPython 3.12, no network).  The pins we do have are (a) a second,
independently written literal NumPy restatement (oracle/numpy_ref.py) that must agree
with this file to 1e-12 in fp64, (b) finite-difference gradient checks and (c) the
structural known-answer tests in tests/test_oracle.py.

Paddle 1.8 operator semantics relied on:
  * layers.fc(x, size, num_flatten_dims, act): x.W[in,out] + b then activation;
    ``act='softmax'`` is appended with axis=-1  ->  at model:107 the soft-max runs over
    the size-1 axis and alpha == 1.0 exactly (softmax_axis='feature').  The
    mathematically intended form (Lu et al. 2017) soft-maxes over the k+1 positions
    (softmax_axis='position').  Both are implemented; nothing is silently "fixed".
  * layers.lstm_unit(x, h, c): concat([x, h]).lstm_w + lstm_b, gate column blocks in
    the order i, f, o, g (forget_bias = 0); c' = sig(f) c + sig(i) tanh(g);
    h' = sig(o) tanh(c').
  * fluid.embedding(padding_idx=0): row 0 reads as zeros and receives no gradient.
  * layers.argmax: first maximal index.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional, Tuple

import torch

# ----------------------------------------------------------------------------------
# configuration: config.py:22-24 (ids) and config.py:51-60 (dims)
# ----------------------------------------------------------------------------------


@dataclasses.dataclass(frozen=True)
class DecoderConfig:
    hidden_dim: int = 1024          # config.py:59
    embedding_size: int = 256       # config.py:56
    vocab_size: int = 12295         # config.py:55
    num_positions: int = 49         # config.py:51 (7x7 encoder map)
    sentence_length: int = 35       # config.py:57
    infer_max_length: int = 35      # config.py:60
    padding_idx: int = 0            # config.py:24
    start_idx: int = 2              # config.py:22
    stop_idx: int = 3               # config.py:23
    softmax_axis: str = "feature"   # 'feature' = reference as written (F2); 'position' = intended

    @property
    def train_steps(self) -> int:   # model:66
        return self.sentence_length - 1


# Parameter names and Paddle layouts ([in, out] for every fc), in construction order
# (model:52, 53, 87-88, 89, 90, 99, 102, 104, 107, 115, 24, 18, 23).
def param_shapes(cfg: DecoderConfig) -> Dict[str, Tuple[int, ...]]:
    H, E, V = cfg.hidden_dim, cfg.embedding_size, cfg.vocab_size
    X = E + H
    return {
        "fc_v_w": (H, H), "fc_v_b": (H,),          # model:52  values  = tanh(fc(P))
        "fc_k_w": (H, H), "fc_k_b": (H,),          # model:53  keys    = fc(P)
        "lstm_w": (X + H, 4 * H), "lstm_b": (4 * H,),  # model:87-88
        "fc_xs_w": (X, H), "fc_xs_b": (H,),        # model:89  p_word_emb
        "fc_hs_w": (H, H), "fc_hs_b": (H,),        # model:90  p_hidden
        "fc_p_w": (H, H), "fc_p_b": (H,),          # model:99  p_hid
        "fc_g_w": (H, H), "fc_g_b": (H,),          # model:102 hid_emb
        "fc_s_w": (H, H), "fc_s_b": (H,),          # model:104 sentinel_emb
        "fc_a_w": (H, 1), "fc_a_b": (1,),          # model:107 alpha score
        "fc_o_w": (H, H), "fc_o_b": (H,),          # model:115 out
        "fc_e_w": (H, E), "fc_e_b": (E,),          # model:24  proj_x
        "word_embedding": (V, E),                  # model:16-19 / model:29-32 (tied)
        "out_fc_bias": (V,),                       # model:20-23
    }


def init_params(cfg: DecoderConfig, seed: int = 0, dtype=torch.float64,
                bias_scale: float = 0.1) -> Dict[str, torch.Tensor]:
    """Xavier-uniform fc weights (Paddle fc default), U(-1,1) embedding (model:19).
    Biases are U(-bias_scale, bias_scale) (Paddle zero-inits them; non-zero biases are
    used in tests so that bias bugs are visible; pass bias_scale=0 for the reference init)."""
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        if name == "word_embedding":
            t = torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1
        elif name.endswith("_w"):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim
        else:
            t = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bias_scale
        out[name] = t.to(dtype)
    return out


# ----------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------


def embedding_function(params, ids: torch.Tensor, cfg: DecoderConfig) -> torch.Tensor:
    """model:28-32 -- lookup with padding_idx (zero row, no gradient)."""
    return torch.nn.functional.embedding(ids, params["word_embedding"], padding_idx=cfg.padding_idx) \
        * (ids != cfg.padding_idx).unsqueeze(-1).to(params["word_embedding"].dtype)


def weight_tying_fc(params, x: torch.Tensor) -> torch.Tensor:
    """model:15-25 -- (x.W_e + b_e).Emb^T + out_fc_bias."""
    proj = x @ params["fc_e_w"] + params["fc_e_b"]
    return proj @ params["word_embedding"].t() + params["out_fc_bias"]


def image_projections(params, p_img_feat: torch.Tensor):
    """model:52-53 -- attention values (tanh) and keys, once per image."""
    img_feat = torch.tanh(p_img_feat @ params["fc_v_w"] + params["fc_v_b"])       # [B,k,H]
    img_feat_emb = p_img_feat @ params["fc_k_w"] + params["fc_k_b"]               # [B,k,H]
    return img_feat, img_feat_emb


def decoder_step(params, cfg: DecoderConfig, word: torch.Tensor, global_img_feat, img_feat,
                 img_feat_emb, hid, cell):
    """One While-loop body, model:84-117.  Returns (logits, h, c, alpha, aux)."""
    H = cfg.hidden_dim
    word_emb = embedding_function(params, word, cfg)                                # model:84
    xt = torch.cat([word_emb, global_img_feat], dim=-1)                            # model:86
    gates = torch.cat([xt, hid], dim=-1) @ params["lstm_w"] + params["lstm_b"]     # model:87-88
    i, f, o, g = gates[:, 0:H], gates[:, H:2 * H], gates[:, 2 * H:3 * H], gates[:, 3 * H:4 * H]
    c = torch.sigmoid(f) * cell + torch.sigmoid(i) * torch.tanh(g)
    h = torch.sigmoid(o) * torch.tanh(c)
    p_word_emb = xt @ params["fc_xs_w"] + params["fc_xs_b"]                        # model:89
    p_hidden = hid @ params["fc_hs_w"] + params["fc_hs_b"]                         # model:90 (previous hid)
    sentinel = torch.sigmoid(p_word_emb + p_hidden) * torch.tanh(c)                # model:91-92
    p_hid = torch.tanh(h @ params["fc_p_w"] + params["fc_p_b"])                    # model:99
    hid_emb = p_hid @ params["fc_g_w"] + params["fc_g_b"]                          # model:102
    sentinel_emb = sentinel @ params["fc_s_w"] + params["fc_s_b"]                  # model:104
    feat_emb = torch.cat([img_feat_emb, sentinel_emb.unsqueeze(1)], dim=1)         # model:105
    z = torch.tanh(feat_emb + hid_emb.unsqueeze(1))                                # model:106
    e = z @ params["fc_a_w"] + params["fc_a_b"]                                    # model:107 [B,k+1,1]
    if cfg.softmax_axis == "feature":
        alpha = torch.softmax(e, dim=-1)          # size-1 axis -> exactly 1 (F2)
    elif cfg.softmax_axis == "position":
        alpha = torch.softmax(e, dim=1)
    else:
        raise ValueError("softmax_axis must be 'feature' or 'position'")
    context = torch.cat([img_feat, sentinel.unsqueeze(1)], dim=1)                  # model:111
    context = (context * alpha).mean(dim=1)                                        # model:112-113
    out = torch.tanh((context + p_hid) @ params["fc_o_w"] + params["fc_o_b"])      # model:115
    logits = weight_tying_fc(params, out)                                          # model:117
    aux = {"sentinel": sentinel, "p_hid": p_hid, "hid_emb": hid_emb, "sentinel_emb": sentinel_emb,
           "context": context, "out": out, "e": e.squeeze(-1)}
    return logits, h, c, alpha.squeeze(-1), aux


def decoder_train(params, cfg: DecoderConfig, global_img_feat, p_img_feat, words,
                  return_aux: bool = False):
    """Decoder(mode='train').call -- model:50-130.  words [B,T] int64 -> logits [B,T,V]."""
    B, T = words.shape
    dt = global_img_feat.dtype
    img_feat, img_feat_emb = image_projections(params, p_img_feat)
    hid = torch.zeros(B, cfg.hidden_dim, dtype=dt)                                 # model:63
    cell = torch.zeros(B, cfg.hidden_dim, dtype=dt)
    outs, alphas, auxs = [], [], []
    for t in range(T):                                                             # model:72-127
        logits, hid, cell, alpha, aux = decoder_step(params, cfg, words[:, t], global_img_feat,
                                                     img_feat, img_feat_emb, hid, cell)
        outs.append(logits)
        alphas.append(alpha)
        if return_aux:
            aux = dict(aux, h=hid, c=cell)
            auxs.append(aux)
    logits = torch.stack(outs, dim=1)                                              # model:129-130
    if return_aux:
        keys = auxs[0].keys()
        return logits, torch.stack(alphas, 1), {k: torch.stack([a[k] for a in auxs], 1) for k in keys}
    return logits, torch.stack(alphas, 1)


def masked_loss(cfg: DecoderConfig, logits, caption):
    """training_network wiring, model:163-171 + 180-183 + 205-212.
    Returns (loss, loss_sum, ntok)."""
    target = caption[:, 1:]
    mask = (target != cfg.padding_idx).to(logits.dtype)
    ce = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), target.reshape(-1),
                                           reduction="none").reshape(target.shape)
    loss_sum = (ce * mask).sum()
    ntok = mask.sum()
    return loss_sum / ntok, loss_sum, ntok


def training_network(params, cfg: DecoderConfig, global_img_feat, p_img_feat, caption):
    """model:161-183 from the decoder boundary onwards: caption [B,35] -> scalar loss."""
    source = caption[:, :-1]
    logits, alpha = decoder_train(params, cfg, global_img_feat, p_img_feat, source)
    loss, _, _ = masked_loss(cfg, logits, caption)
    return loss, logits, alpha


def loss_and_grads(params, cfg, global_img_feat, p_img_feat, caption):
    """Forward + autograd backward.  Returns loss, logits, alpha and a dict of gradients for every
    parameter plus 'global_img_feat' and 'p_img_feat' (the gradient handed to the encoder side)."""
    ps = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    g = global_img_feat.detach().clone().requires_grad_(True)
    P = p_img_feat.detach().clone().requires_grad_(True)
    loss, logits, alpha = training_network(ps, cfg, g, P, caption)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in ps.items()}
    grads["global_img_feat"] = g.grad
    grads["p_img_feat"] = P.grad
    return loss.detach(), logits.detach(), alpha.detach(), grads


# ----------------------------------------------------------------------------------
# inference: greedy (reference, model:119-123) and beam (our spec)
# ----------------------------------------------------------------------------------


@torch.no_grad()
def greedy_decode(params, cfg: DecoderConfig, global_img_feat, p_img_feat, return_margin=False):
    """Decoder(mode='eval').call -- fixed infer_max_length steps, never stops early (model:66-68),
    returns float32 ids [B,T] (model:122, F9).  margin[b,t] = top1 - top2 logit gap."""
    B = global_img_feat.shape[0]
    dt = global_img_feat.dtype
    img_feat, img_feat_emb = image_projections(params, p_img_feat)
    word = torch.full((B,), cfg.start_idx, dtype=torch.int64)                      # model:56-58
    hid = torch.zeros(B, cfg.hidden_dim, dtype=dt)
    cell = torch.zeros(B, cfg.hidden_dim, dtype=dt)
    ids, margins = [], []
    for _ in range(cfg.infer_max_length):
        logits, hid, cell, _, _ = decoder_step(params, cfg, word, global_img_feat, img_feat,
                                               img_feat_emb, hid, cell)
        word = first_argmax(logits)                                                # model:120-121
        ids.append(word.to(torch.float32))                                         # model:122
        if return_margin:
            top2 = logits.topk(2, dim=-1).values
            margins.append(top2[:, 0] - top2[:, 1])
    out = torch.stack(ids, dim=1)                                                  # model:132-133
    if return_margin:
        return out, torch.stack(margins, 1)
    return out


def first_argmax(x: torch.Tensor) -> torch.Tensor:
    """Paddle arg_max: lowest index among maxima."""
    V = x.shape[-1]
    mx = x.max(dim=-1, keepdim=True).values
    idx = torch.arange(V).expand_as(x)
    return torch.where(x == mx, idx, torch.full_like(idx, V)).min(dim=-1).values


@torch.no_grad()
def beam_select(cfg: DecoderConfig, score, logp, finished, beam: int, finish_on_stop: bool = True):
    """One selection step of the beam spec for every image.

    score [B,beam] prefix scores (-inf = dead prefix), logp [B,beam,V] log-probabilities of the next token,
    finished [B,beam] bool.  A finished prefix offers exactly one candidate (itself, score unchanged, emitting
    padding_idx); the top-`beam` of the <= beam*V candidates win, ties -> lower flat index = (lower prefix, lower
    token) -- a stable descending sort.  Returns (top_val [B,beam], src [B,beam], tok [B,beam], new_finished,
    margin [B] = gap between the last winner and the first loser, adjacent [B] = smallest gap between neighbours
    among the best beam+1 candidates: below it the ORDER of the winners may change too)."""
    B, _, V = logp.shape
    NEG = -float("inf")
    cand = score.unsqueeze(-1) + logp                   # [B,beam,V]
    if finish_on_stop:
        fin = finished.unsqueeze(-1)
        only_pad = torch.full_like(cand, NEG)
        only_pad[:, :, cfg.padding_idx] = score         # the single "stay finished" candidate
        cand = torch.where(fin, only_pad, cand)
    flat = cand.reshape(B, beam * V)
    # stable top-k: ties -> lower flat index = (lower prefix, lower token)
    order = torch.sort(flat, dim=-1, descending=True, stable=True)
    top_val, top_idx = order.values[:, :beam], order.indices[:, :beam]
    if flat.shape[1] > beam:
        margin = order.values[:, beam - 1] - order.values[:, beam]
        gaps = order.values[:, :beam] - order.values[:, 1:beam + 1]
        gaps = torch.where(torch.isnan(gaps), torch.full_like(gaps, float("inf")), gaps)   # -inf - -inf
        adjacent = gaps.min(dim=1).values
    else:
        margin = torch.full((B,), float("inf"), dtype=score.dtype)
        adjacent = margin.clone()
    src = top_idx // V                                   # backpointer [B,beam]
    tok = top_idx % V
    was_fin = torch.gather(finished, 1, src)
    new_fin = was_fin | ((tok == cfg.stop_idx) if finish_on_stop else torch.zeros_like(was_fin))
    return top_val, src, tok, new_fin, margin, adjacent


@torch.no_grad()
def beam_search(params, cfg: DecoderConfig, global_img_feat, p_img_feat, beam: int,
                finish_on_stop: bool = True, return_margin: bool = False, return_trace: bool = False):
    """Beam search as specified in SURVEY.md 8c row C5 (the reference has none, F1).

    rows = image x beam; score = sum of log-probs, no length normalisation; a prefix that ended in
    stop_idx contributes exactly one candidate (itself, score unchanged) and emits padding_idx
    afterwards; per image the top-`beam` candidates are taken with ties broken by (lower prefix
    index, lower token id); at step 0 only prefix 0 is live; exactly infer_max_length steps.
    Returns ids int64 [B,beam,T] sorted by final score (desc, stable) and scores [B,beam]
    (+ margin [B,T] with return_margin; + a dict of per-step records with return_trace: tok, src [T,B,beam],
    adjacent [B,T] -- see beam_select).
    With beam=1 and finish_on_stop=False this is bit-identical to greedy_decode."""
    B = global_img_feat.shape[0]
    dt = global_img_feat.dtype
    T, V, H = cfg.infer_max_length, cfg.vocab_size, cfg.hidden_dim
    NEG = -float("inf")
    img_feat, img_feat_emb = image_projections(params, p_img_feat)
    rep = lambda x: x.repeat_interleave(beam, dim=0)
    g_r, v_r, k_r = rep(global_img_feat), rep(img_feat), rep(img_feat_emb)
    word = torch.full((B * beam,), cfg.start_idx, dtype=torch.int64)
    hid = torch.zeros(B * beam, H, dtype=dt)
    cell = torch.zeros(B * beam, H, dtype=dt)
    score = torch.zeros(B, beam, dtype=dt)
    score[:, 1:] = NEG                                      # only prefix 0 live at step 0
    finished = torch.zeros(B, beam, dtype=torch.bool)
    hist = torch.zeros(B, beam, T, dtype=torch.int64)
    margins, adjacents, toks, srcs = [], [], [], []
    for t in range(T):
        logits, h_new, c_new, _, _ = decoder_step(params, cfg, word, g_r, v_r, k_r, hid, cell)
        logp = torch.log_softmax(logits, dim=-1).reshape(B, beam, V)
        score, src, tok, finished, margin, adjacent = beam_select(cfg, score, logp, finished, beam, finish_on_stop)
        margins.append(margin)
        adjacents.append(adjacent)
        toks.append(tok)
        srcs.append(src)
        base = (torch.arange(B) * beam).unsqueeze(-1)
        rows = (base + src).reshape(-1)
        hid, cell = h_new[rows], c_new[rows]
        hist = torch.gather(hist, 1, src.unsqueeze(-1).expand(B, beam, T)).clone()
        hist[:, :, t] = tok
        word = tok.reshape(-1)
    # beams are already in descending score order (stable sort each step)
    out = (hist, score)
    if return_margin:
        out = out + (torch.stack(margins, 1),)
    if return_trace:
        out = out + ({"tok": torch.stack(toks, 0), "src": torch.stack(srcs, 0), "adjacent": torch.stack(adjacents, 1)},)
    return out


def filter_caption(ids, cfg: DecoderConfig):
    """evaluate.py:15-25 without the id->word map: cut at stop, drop padding."""
    out = []
    for i in ids:
        i = int(round(float(i)))
        if i == cfg.stop_idx:
            break
        if i == cfg.padding_idx:
            continue
        out.append(i)
    return out
