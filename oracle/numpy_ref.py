"""CPU ORACLE, second restatement (test infrastructure, NOT product code) -- PARITY UNPINNED.

A literal NumPy forward of /root/reference/ImageCaptioning/model/model_adaAttention_aic.py
written independently of oracle/decoder_ref.py: every Paddle op of the While body is spelled out
in the reference's order (unsqueeze / expand / concat / elementwise_mul(axis=0) / reduce_mean),
intermediates are materialised exactly as the Paddle graph would.  It exists to pin
decoder_ref.py (tests/test_oracle.py requires agreement to 1e-12 in fp64).
"""
from __future__ import annotations

import numpy as np


def _fc(x, w, b, num_flatten_dims=1, act=None):
    """Paddle layers.fc: flatten leading `num_flatten_dims` dims, x.W + b, activation (axis=-1)."""
    lead = x.shape[:num_flatten_dims]
    y = x.reshape(int(np.prod(lead)), -1) @ w + b
    y = y.reshape(*lead, w.shape[1])
    if act == "tanh":
        y = np.tanh(y)
    elif act == "softmax":                       # append_activation -> softmax over the LAST axis
        m = y.max(axis=-1, keepdims=True)
        ex = np.exp(y - m)
        y = ex / ex.sum(axis=-1, keepdims=True)
    elif act is not None:
        raise ValueError(act)
    return y


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _softmax(y, axis):
    m = y.max(axis=axis, keepdims=True)
    ex = np.exp(y - m)
    return ex / ex.sum(axis=axis, keepdims=True)


def _embedding(ids, table, padding_idx):
    out = table[ids]
    out[ids == padding_idx] = 0.0                # fluid.embedding(padding_idx): zero row
    return out


def _lstm_unit(x_t, h_prev, c_prev, w, b, forget_bias=0.0):
    """Paddle layers.lstm_unit: fc(concat([x,h])) then lstm_unit op, gate order i,f,o,g."""
    D = h_prev.shape[1]
    fc_out = np.concatenate([x_t, h_prev], axis=1) @ w + b
    i = _sigmoid(fc_out[:, 0 * D:1 * D])
    f = _sigmoid(fc_out[:, 1 * D:2 * D] + forget_bias)
    o = _sigmoid(fc_out[:, 2 * D:3 * D])
    g = np.tanh(fc_out[:, 3 * D:4 * D])
    c = f * c_prev + i * g
    h = o * np.tanh(c)
    return h, c


def decoder_call(p, global_img_feat, p_img_feat, words=None, mode="train", hid_size=1024,
                 infer_max_length=35, start_idx=2, padding_idx=0, softmax_axis="feature"):
    """Decoder.call, model:50-135.  `p` maps the names of decoder_ref.param_shapes to ndarrays."""
    if mode not in ("train", "eval"):
        raise ValueError("unsupported mode {}".format(mode))                       # model:43-44
    B = global_img_feat.shape[0]
    img_feat = _fc(p_img_feat, p["fc_v_w"], p["fc_v_b"], 2, "tanh")                # model:52
    img_feat_emb = _fc(p_img_feat, p["fc_k_w"], p["fc_k_b"], 2)                    # model:53
    if mode == "eval":
        word = np.full((B,), start_idx, dtype=np.int64)                            # model:56-58
        mx = infer_max_length
    else:
        words_tm = words.T                                                         # model:60
        mx = words.shape[1]
    hid = np.zeros((B, hid_size), dtype=global_img_feat.dtype)                     # model:63
    cell = np.zeros((B, hid_size), dtype=global_img_feat.dtype)
    outputs = []
    step = 0
    while step < mx:                                                               # model:72-77
        if mode == "train":
            word = words_tm[step]                                                  # model:79-82
        word_emb = _embedding(word, p["word_embedding"], padding_idx)              # model:84
        xt = np.concatenate([word_emb, global_img_feat], axis=-1)                  # model:86
        h, c = _lstm_unit(xt, hid, cell, p["lstm_w"], p["lstm_b"])                 # model:87-88
        p_word_emb = _fc(xt, p["fc_xs_w"], p["fc_xs_b"])                           # model:89
        p_hidden = _fc(hid, p["fc_hs_w"], p["fc_hs_b"])                            # model:90
        sentinel_gate = _sigmoid(p_word_emb + p_hidden)                            # model:91
        sentinel = sentinel_gate * np.tanh(c)                                      # model:92
        hid, cell = h, c                                                           # model:94-95
        k = p_img_feat.shape[1]                                                    # model:97
        p_hid = _fc(h, p["fc_p_w"], p["fc_p_b"], act="tanh")                       # model:99
        hid_emb = _fc(p_hid, p["fc_g_w"], p["fc_g_b"])                             # model:102
        exp_hid_emb = np.tile(hid_emb[:, None, :], (1, k + 1, 1))                  # model:103
        sentinel_emb = _fc(sentinel, p["fc_s_w"], p["fc_s_b"])[:, None, :]         # model:104
        feat_emb = np.concatenate([img_feat_emb, sentinel_emb], axis=1)            # model:105
        z = np.tanh(feat_emb + exp_hid_emb)                                        # model:106
        if softmax_axis == "feature":
            alpha = _fc(z, p["fc_a_w"], p["fc_a_b"], 2, "softmax")                 # model:107 as written
        else:
            alpha = _softmax(_fc(z, p["fc_a_w"], p["fc_a_b"], 2), axis=1)          # intended
        context = np.concatenate([img_feat, sentinel[:, None, :]], axis=1)         # model:111
        context = context * alpha                                                  # model:112 (axis=0 bcast)
        context = context.mean(axis=1)                                             # model:113
        out = _fc(context + p_hid, p["fc_o_w"], p["fc_o_b"], act="tanh")           # model:115
        proj_x = _fc(out, p["fc_e_w"], p["fc_e_b"])                                # model:24
        word_pred = proj_x @ p["word_embedding"].T + p["out_fc_bias"]              # model:25
        if mode == "eval":
            next_word = np.argmax(word_pred, axis=-1)                              # model:120 (first max)
            word = next_word                                                       # model:121
            outputs.append(next_word.astype(np.float32))                           # model:122-123
        else:
            outputs.append(word_pred)                                              # model:125
        step += 1                                                                  # model:126
    if mode == "train":
        return np.transpose(np.stack(outputs, 0), (1, 0, 2))                       # model:129-130
    return np.transpose(np.stack(outputs, 0), (1, 0))                              # model:132-133


def training_loss(p, global_img_feat, p_img_feat, caption, padding_idx=0, **kw):
    """training_network, model:163-183 + loss model:205-212 (decoder boundary onwards)."""
    target = caption[:, 1:]
    source = caption[:, :-1]
    mask = np.logical_not(target == padding_idx).astype(global_img_feat.dtype)
    scale_factor = mask.sum()
    seq_out = decoder_call(p, global_img_feat, p_img_feat, words=source, mode="train",
                           padding_idx=padding_idx, **kw)
    m = seq_out.max(axis=-1, keepdims=True)
    lse = np.log(np.exp(seq_out - m).sum(axis=-1, keepdims=True)) + m
    los = (lse - np.take_along_axis(seq_out, target[..., None], axis=-1))[..., 0]
    return (los * mask).sum() / scale_factor, seq_out
