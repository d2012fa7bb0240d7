"""Deterministic synthetic inputs of the decoder hot path, used by bench.py and the tests.

  P = relu(N(0,1)) [B,k,H], g = relu(N(0,1)) [B,H]     post-ReLU like model:196-198
  captions [B,S] = [start, w_1..w_L, stop, pad...], L ~ U{5..S-2}, w ~ U{4..V-1}
  (preprocess/ai_challenge_tokenizer.py:47, 81-86: <pad>=0 <unk>=1 <start>=2 <stop>=3; config.py:9: <= 33 words)
Only torch-CPU generators are used so a seed gives the same bytes on every machine.  `cfg` is any object with the
attributes hidden_dim, num_positions, sentence_length, vocab_size, start_idx, stop_idx.
"""
from __future__ import annotations

import torch


def make_features(cfg, B: int, seed: int = 20260, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    P = torch.relu(torch.randn(B, cfg.num_positions, cfg.hidden_dim, generator=gen, dtype=torch.float32))
    g = torch.relu(torch.randn(B, cfg.hidden_dim, generator=gen, dtype=torch.float32))
    return g.to(dtype), P.to(dtype)


def make_captions(cfg, B: int, seed: int = 20260, min_len: int = 5, max_len=None):
    gen = torch.Generator().manual_seed(seed + 7919)
    S = cfg.sentence_length
    max_len = S - 2 if max_len is None else max_len
    min_len = min(min_len, max_len)
    lens = torch.randint(min_len, max_len + 1, (B,), generator=gen)
    words = torch.randint(4, cfg.vocab_size, (B, S), generator=gen)
    cap = torch.zeros(B, S, dtype=torch.int64)
    pos = torch.arange(S).unsqueeze(0)
    L = lens.unsqueeze(1)
    cap = torch.where((pos >= 1) & (pos <= L), torch.roll(words, 1, dims=1), cap)
    cap[:, 0] = cfg.start_idx
    cap[torch.arange(B), lens + 1] = cfg.stop_idx
    return cap


def make_encoder_map(cfg, B: int, encoder_channel: int = 1280, seed: int = 20260, dtype=torch.float32):
    """Synthetic MobileNetV2 output [B, X, k] (model:193-195): the encoder ends in ReLU6, so values lie in [0, 6]."""
    gen = torch.Generator().manual_seed(seed + 31)
    x = torch.randn(B, encoder_channel, cfg.num_positions, generator=gen, dtype=torch.float32)
    return torch.clamp(x, 0.0, 6.0).to(dtype)


def make_batch(cfg, B: int, seed: int = 20260, dtype=torch.float32):
    g, P = make_features(cfg, B, seed, dtype)
    return g, P, make_captions(cfg, B, seed)
