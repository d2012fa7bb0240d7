"""Host-side mirror of the reference decoder interface over the C-ABI (include/capdec.h).

Mirrors /root/reference/ImageCaptioning/model/model_adaAttention_aic.py:
  * ``Decoder(hidden_size, mode='train'|'eval', rnn_layer=1).call(global_img_feat, p_img_feat, embedding_fn, words)``
    (model:40-135): train -> logits [B,T,V] fp32 batch-major; eval -> float32 ids [B,infer_max_length];
    an unsupported mode raises ValueError exactly like model:43-44.
  * ``ImageCaptionModel.training_network / eval_network / loss`` from the decoder boundary on (model:161-189, 205-212),
    and ``_img2feature`` / ``training_network_from_map`` / ``eval_network_from_map`` from the encoder output on
    (model:193-199: the feature adaptor in front of the decoder).
PyTorch is used for tensor ownership, streams and torch.distributed only; all arithmetic is in libcapdec.so.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from typing import Dict, Optional

import torch

from . import _capi
from ._capi import (ADAPTOR_NAMES, CapdecAdaptorGrads, CapdecAdaptorParams, CapdecConfig, CapdecGrads, CapdecParams,
                    PARAM_NAMES, check, ptr, stream_ptr)


@dataclasses.dataclass(frozen=True)
class DecoderConfig:
    """config.py:22-24 and config.py:51-60 of the reference, plus the two build-time choices of this library."""
    hidden_dim: int = 1024
    embedding_size: int = 256
    vocab_size: int = 12295
    num_positions: int = 49
    sentence_length: int = 35
    infer_max_length: int = 35
    padding_idx: int = 0
    start_idx: int = 2
    stop_idx: int = 3
    dtype: str = "bf16"             # 'bf16' (tcgen05 path) | 'fp32' (validation precision)
    softmax_axis: str = "feature"   # 'feature' = reference as written (alpha == 1) | 'position' = intended
    stepwise_recurrence: bool = False   # validation: one GEMM + cell launch per step instead of the persistent kernels

    @property
    def train_steps(self) -> int:
        return self.sentence_length - 1

    def to_c(self) -> CapdecConfig:
        if self.dtype not in ("bf16", "fp32"):
            raise ValueError("unsupported dtype {}".format(self.dtype))
        if self.softmax_axis not in ("feature", "position"):
            raise ValueError("unsupported softmax_axis {}".format(self.softmax_axis))
        return CapdecConfig(self.hidden_dim, self.embedding_size, self.vocab_size, self.num_positions,
                            self.sentence_length, self.infer_max_length, self.padding_idx, self.start_idx,
                            self.stop_idx, _capi.DTYPE_BF16 if self.dtype == "bf16" else _capi.DTYPE_F32,
                            _capi.SOFTMAX_POSITION if self.softmax_axis == "position" else _capi.SOFTMAX_FEATURE,
                            _capi.FLAG_STEPWISE_RECURRENCE if self.stepwise_recurrence else 0)


def param_shapes(cfg: DecoderConfig) -> Dict[str, tuple]:
    """PaddlePaddle layouts ([in, out] fc weights), construction order of model:52..115, 24, 18, 23."""
    H, E, V = cfg.hidden_dim, cfg.embedding_size, cfg.vocab_size
    X = E + H
    return {
        "fc_v_w": (H, H), "fc_v_b": (H,), "fc_k_w": (H, H), "fc_k_b": (H,),
        "lstm_w": (X + H, 4 * H), "lstm_b": (4 * H,),
        "fc_xs_w": (X, H), "fc_xs_b": (H,), "fc_hs_w": (H, H), "fc_hs_b": (H,),
        "fc_p_w": (H, H), "fc_p_b": (H,), "fc_g_w": (H, H), "fc_g_b": (H,), "fc_s_w": (H, H), "fc_s_b": (H,),
        "fc_a_w": (H, 1), "fc_a_b": (1,), "fc_o_w": (H, H), "fc_o_b": (H,), "fc_e_w": (H, E), "fc_e_b": (E,),
        "word_embedding": (V, E), "out_fc_bias": (V,),
    }


# Order of the tensors inside the flat buffers (the C structs address every tensor by name, so any order works): the
# order in which capdec_train_bwd finalises the gradients, so that each group of capdec_set_grad_ready_events is ONE
# contiguous range = one all-reduce call; lstm_w comes last because its word rows (the first rows) are final last.
FLAT_ORDER = ["fc_hs_w", "fc_hs_b", "fc_p_w", "fc_p_b", "fc_g_w", "fc_g_b", "fc_s_w", "fc_s_b", "fc_a_w", "fc_a_b",
              "fc_o_w", "fc_o_b", "fc_e_w", "fc_e_b", "fc_v_w", "fc_v_b", "fc_k_w", "fc_k_b",
              "word_embedding", "out_fc_bias", "lstm_b", "fc_xs_w", "fc_xs_b", "lstm_w"]
assert sorted(FLAT_ORDER) == sorted(PARAM_NAMES)


class ParameterStore:
    """All decoder parameters (and their gradients) as views into two flat fp32 device buffers, so that a
    data-parallel gradient all-reduce is a single NCCL call.  Every view starts on a 256-byte boundary."""

    def __init__(self, cfg: DecoderConfig, device):
        self.cfg = cfg
        self.device = torch.device(device)
        self.shapes = param_shapes(cfg)
        self.offsets = {}
        off = 0
        for n in FLAT_ORDER:
            self.offsets[n] = off
            off += (math.prod(self.shapes[n]) + 63) // 64 * 64
        self.numel_padded = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.flat_grad = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.params = {n: self._view(self.flat, n) for n in PARAM_NAMES}
        self.grads = {n: self._view(self.flat_grad, n) for n in PARAM_NAMES}
        self.c_params = CapdecParams(*[C.c_void_p(self.params[n].data_ptr()) for n in PARAM_NAMES])
        self.c_grads = CapdecGrads(*[C.c_void_p(self.grads[n].data_ptr()) for n in PARAM_NAMES])

    def _view(self, flat, name):
        o = self.offsets[name]
        return flat[o:o + math.prod(self.shapes[name])].view(self.shapes[name])

    @property
    def num_parameters(self) -> int:
        return sum(math.prod(s) for s in self.shapes.values())

    def init_reference(self, seed: int = 0, bias_scale: float = 0.0):
        """PaddlePaddle default initialisers of the reference graph: Xavier-uniform fc weights, zero biases
        (bias_scale > 0 draws U(-s, s) instead, used by tests), U(-1, 1) tied embedding (model:19)."""
        gen = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for n in PARAM_NAMES:
                shape = self.shapes[n]
                if n == "word_embedding":
                    t = torch.rand(shape, generator=gen) * 2 - 1
                elif n.endswith("_w"):
                    lim = math.sqrt(6.0 / (shape[0] + shape[1]))
                    t = (torch.rand(shape, generator=gen) * 2 - 1) * lim
                else:
                    t = (torch.rand(shape, generator=gen) * 2 - 1) * bias_scale
                self.params[n].copy_(t)
        _capi.bump_version(self.flat)
        return self

    def load(self, state: Dict[str, "torch.Tensor"]):
        """Copy parameters given by name (torch tensors or numpy arrays, Paddle layouts)."""
        for n in PARAM_NAMES:
            if n not in state:
                raise KeyError("missing parameter {}".format(n))
            t = torch.as_tensor(state[n]).to(torch.float32)
            if tuple(t.shape) != tuple(self.shapes[n]):
                raise ValueError("parameter {} has shape {}, expected {}".format(n, tuple(t.shape), self.shapes[n]))
            with torch.no_grad():
                self.params[n].copy_(t)
        _capi.bump_version(self.flat)
        return self

    def state_dict(self):
        return {n: self.params[n].detach().cpu().clone() for n in PARAM_NAMES}


class AdaptorStore:
    """Parameters and gradients of the feature adaptor `_img2feature` (model:197, 199) as views into two flat fp32
    buffers of their own (the decoder's ParameterStore layout stays what capdec_set_grad_ready_events describes)."""

    def __init__(self, cfg: DecoderConfig, encoder_channel: int, device):
        X, H = int(encoder_channel), cfg.hidden_dim
        if X <= 0 or X % 8:
            raise ValueError("encoder_channel must be a positive multiple of 8, got {}".format(X))
        self.encoder_channel = X
        self.device = torch.device(device)
        self.shapes = {"fc_img_w": (X, H), "fc_img_b": (H,), "fc_glob_w": (X, H), "fc_glob_b": (H,)}
        self.offsets, off = {}, 0
        for n in ADAPTOR_NAMES:
            self.offsets[n] = off
            off += (math.prod(self.shapes[n]) + 63) // 64 * 64
        self.numel_padded = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.flat_grad = torch.zeros(off, dtype=torch.float32, device=self.device)
        view = lambda flat, n: flat[self.offsets[n]:self.offsets[n] + math.prod(self.shapes[n])].view(self.shapes[n])
        self.params = {n: view(self.flat, n) for n in ADAPTOR_NAMES}
        self.grads = {n: view(self.flat_grad, n) for n in ADAPTOR_NAMES}
        self.c_params = CapdecAdaptorParams(*[C.c_void_p(self.params[n].data_ptr()) for n in ADAPTOR_NAMES])
        self.c_grads = CapdecAdaptorGrads(*[C.c_void_p(self.grads[n].data_ptr()) for n in ADAPTOR_NAMES])

    @property
    def num_parameters(self) -> int:
        return sum(math.prod(s) for s in self.shapes.values())

    def init_reference(self, seed: int = 0, bias_scale: float = 0.0):
        """Paddle fc defaults: Xavier-uniform weights, zero biases (bias_scale > 0: U(-s, s), used by tests)."""
        gen = torch.Generator().manual_seed(seed + 104729)
        with torch.no_grad():
            for n in ADAPTOR_NAMES:
                shape = self.shapes[n]
                if n.endswith("_w"):
                    t = (torch.rand(shape, generator=gen) * 2 - 1) * math.sqrt(6.0 / (shape[0] + shape[1]))
                else:
                    t = (torch.rand(shape, generator=gen) * 2 - 1) * bias_scale
                self.params[n].copy_(t)
        _capi.bump_version(self.flat)
        return self

    def load(self, state: Dict[str, "torch.Tensor"]):
        for n in ADAPTOR_NAMES:
            if n not in state:
                raise KeyError("missing parameter {}".format(n))
            t = torch.as_tensor(state[n]).to(torch.float32)
            if tuple(t.shape) != tuple(self.shapes[n]):
                raise ValueError("parameter {} has shape {}, expected {}".format(n, tuple(t.shape), self.shapes[n]))
            with torch.no_grad():
                self.params[n].copy_(t)
        _capi.bump_version(self.flat)
        return self

    def state_dict(self):
        return {n: self.params[n].detach().cpu().clone() for n in ADAPTOR_NAMES}


class CaptionDecoderEngine:
    """One handle per (process, GPU): owns the C handle and the ParameterStore; thin, sync-free wrappers of the
    C-ABI entry points.  This is what bench.py and the reference-shaped classes below drive."""

    def __init__(self, cfg: DecoderConfig = DecoderConfig(), device=None, seed: Optional[int] = 0,
                 bias_scale: float = 0.0):
        if not torch.cuda.is_available():
            raise _capi.CapdecError("no CUDA device: the caption decoder has no CPU fallback")
        self.lib = _capi.load_library()
        self.cfg = cfg
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = C.c_void_p()
        ccfg = cfg.to_c()
        check(self.lib.capdec_create(C.byref(ccfg), self.device.index or 0, C.byref(h)), "capdec_create")
        self.handle = h
        self.store = ParameterStore(cfg, self.device)
        if seed is not None:
            self.store.init_reference(seed, bias_scale)
        self._loss_buf = torch.zeros(4, dtype=torch.float32, device=self.device)
        self._caption_keepalive = None
        self._bound = False
        self._bound_version = -1
        self._adaptor_version = -1
        self.adaptor: Optional[AdaptorStore] = None      # attach_adaptor() creates the feature-adaptor parameters
        self._adaptor_bound = False
        self._staged = None

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.capdec_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    # ---- C-ABI wrappers --------------------------------------------------------------------------------------
    def bind_params(self):
        """capdec_bind_params: re-layout of the fp32 parameters into the handle's operand copies."""
        check(self.lib.capdec_bind_params(self.handle, C.byref(self.store.c_params), stream_ptr()), "capdec_bind_params")
        self._bound = True
        self._bound_version = _capi.version_of(self.store.flat)

    def _ensure_bound(self):
        """(Re-)bind when the parameters were written since the last bind: optim.Adam, ParameterStore.load /
        init_reference and checkpoint.load_params bump the buffer's version (_capi.bump_version).  A caller that writes
        store.flat by other means calls bind_params() itself (include/capdec.h)."""
        if not self._bound or self._bound_version != _capi.version_of(self.store.flat):
            self.bind_params()

    def _ensure_adaptor_bound(self):
        if not self._adaptor_bound or self._adaptor_version != _capi.version_of(self.adaptor.flat):
            self.bind_adaptor()

    def _check_inputs(self, g, P, caption=None):
        cfg = self.cfg
        B = g.shape[0]
        if tuple(g.shape) != (B, cfg.hidden_dim) or tuple(P.shape) != (B, cfg.num_positions, cfg.hidden_dim):
            raise ValueError("expected global_img_feat [B,{H}] and p_img_feat [B,{k},{H}], got {} and {}".format(
                tuple(g.shape), tuple(P.shape), H=cfg.hidden_dim, k=cfg.num_positions))
        for t in (g, P):
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                raise ValueError("features must be contiguous float32 CUDA tensors")
        if caption is not None:
            if tuple(caption.shape) != (B, cfg.sentence_length) or caption.dtype != torch.int64 \
                    or not caption.is_cuda or not caption.is_contiguous():
                raise ValueError("caption must be a contiguous int64 CUDA tensor [B,{}]".format(cfg.sentence_length))
        return B

    def _host_lengths(self, lengths, B):
        """`lengths` (HOST: list, numpy array or CPU tensor of B token counts, see caption_lengths) as the int32 array
        capdec_train_fwd reads before it returns; None stays None (no packing)."""
        if lengths is None:
            return None, None
        if isinstance(lengths, torch.Tensor):
            if lengths.is_cuda:
                raise ValueError("lengths must live on the host (it sizes the launches); pass caption_lengths(caption_cpu)")
            lengths = lengths.numpy()
        import numpy as np
        arr = np.ascontiguousarray(lengths, dtype=np.int32)
        if arr.shape != (B,):
            raise ValueError("lengths must have shape [{}], got {}".format(B, arr.shape))
        return arr, arr.ctypes.data_as(C.c_void_p)

    def train_forward(self, g, P, caption, want_logits: bool = False, lengths=None):
        """capdec_train_fwd.  Returns (loss_buf[4] device tensor, logits or None).
        lengths: optional HOST array [B] of non-padding token counts per caption (caption_lengths): the time-batched
        head then runs on the rows that carry a target only -- same results, less work."""
        B = self._check_inputs(g, P, caption)
        self._ensure_bound()
        logits = None
        if want_logits:
            logits = torch.empty(B, self.cfg.train_steps, self.cfg.vocab_size, dtype=torch.float32, device=self.device)
        self._caption_keepalive = (g, P, caption)
        arr, lp = self._host_lengths(lengths, B)
        check(self.lib.capdec_train_fwd(self.handle, ptr(g), ptr(P), ptr(caption), lp, B, ptr(logits),
                                        ptr(self._loss_buf), stream_ptr()), "capdec_train_fwd")
        return self._loss_buf, logits

    def train_backward(self, loss_scale: float = 1.0, inv_ntok: Optional[torch.Tensor] = None, want_dg: bool = True,
                       want_dP: bool = True):
        """capdec_train_bwd.  Fills store.flat_grad; returns (d_global_img_feat, d_p_img_feat)."""
        g, P, _ = self._caption_keepalive
        dg = torch.empty_like(g) if want_dg else None
        dP = torch.empty_like(P) if want_dP else None
        check(self.lib.capdec_train_bwd(self.handle, C.c_float(loss_scale), ptr(inv_ntok), ptr(dg), ptr(dP),
                                        C.byref(self.store.c_grads), stream_ptr()), "capdec_train_bwd")
        return dg, dP

    # ---- feature adaptor: _img2feature from the encoder map on (model:193-199) -----------------------------------
    def attach_adaptor(self, encoder_channel: int = 1280, seed: Optional[int] = 0, bias_scale: float = 0.0):
        """Create the adaptor parameters (config.py:52 encoder_channel) next to the decoder's; returns the store."""
        self.adaptor = AdaptorStore(self.cfg, encoder_channel, self.device)
        if seed is not None:
            self.adaptor.init_reference(seed, bias_scale)
        self._adaptor_bound = False
        return self.adaptor

    def bind_adaptor(self):
        if getattr(self, "adaptor", None) is None:
            raise _capi.CapdecError("no adaptor attached: call attach_adaptor(encoder_channel) first")
        check(self.lib.capdec_bind_adaptor(self.handle, self.adaptor.encoder_channel, C.byref(self.adaptor.c_params),
                                           stream_ptr()), "capdec_bind_adaptor")
        self._adaptor_bound = True
        self._adaptor_version = _capi.version_of(self.adaptor.flat)

    def _check_map(self, enc_map):
        ad = getattr(self, "adaptor", None)
        if ad is None:
            raise _capi.CapdecError("no adaptor attached: call attach_adaptor(encoder_channel) first")
        B = enc_map.shape[0]
        if enc_map.dim() == 4:                                     # [B,1280,7,7] as the encoder returns it (model:193)
            enc_map = enc_map.reshape(B, enc_map.shape[1], -1)     # layers.reshape(img, [0, 0, -1]), model:195
        if tuple(enc_map.shape) != (B, ad.encoder_channel, self.cfg.num_positions):
            raise ValueError("expected an encoder map [B,{},{}], got {}".format(ad.encoder_channel,
                                                                                self.cfg.num_positions,
                                                                                tuple(enc_map.shape)))
        if enc_map.dtype != torch.float32 or not enc_map.is_cuda or not enc_map.is_contiguous():
            raise ValueError("the encoder map must be a contiguous float32 CUDA tensor")
        return enc_map, B

    def img2feature(self, enc_map, want_outputs: bool = False):
        """capdec_img2feature: stages (p_img_feat, global_img_feat) in the handle; returns fp32 copies if asked."""
        enc_map, B = self._check_map(enc_map)
        self._ensure_adaptor_bound()
        P = g = None
        if want_outputs:
            P = torch.empty(B, self.cfg.num_positions, self.cfg.hidden_dim, dtype=torch.float32, device=self.device)
            g = torch.empty(B, self.cfg.hidden_dim, dtype=torch.float32, device=self.device)
        check(self.lib.capdec_img2feature(self.handle, ptr(enc_map), B, ptr(P), ptr(g), stream_ptr()),
              "capdec_img2feature")
        self._staged = (enc_map, B)
        return P, g

    def train_forward_map(self, enc_map, caption, want_logits: bool = False, lengths=None):
        """_img2feature + teacher-forced forward from the encoder map (model:177-183 after the encoder)."""
        self.img2feature(enc_map)
        enc_map, B = self._staged
        cfg = self.cfg
        if tuple(caption.shape) != (B, cfg.sentence_length) or caption.dtype != torch.int64 or not caption.is_cuda \
                or not caption.is_contiguous():
            raise ValueError("caption must be a contiguous int64 CUDA tensor [B,{}]".format(cfg.sentence_length))
        self._ensure_bound()
        logits = None
        if want_logits:
            logits = torch.empty(B, cfg.train_steps, cfg.vocab_size, dtype=torch.float32, device=self.device)
        self._caption_keepalive = (None, None, caption)
        arr, lp = self._host_lengths(lengths, B)
        check(self.lib.capdec_train_fwd(self.handle, None, None, ptr(caption), lp, B, ptr(logits), ptr(self._loss_buf),
                                        stream_ptr()), "capdec_train_fwd")
        return self._loss_buf, logits

    def train_backward_map(self, loss_scale: float = 1.0, inv_ntok: Optional[torch.Tensor] = None,
                           want_dmap: bool = True):
        """capdec_train_bwd + capdec_img2feature_bwd.  Fills store.flat_grad and adaptor.flat_grad; returns the
        gradient w.r.t. the encoder map [B,X,k] (or None)."""
        enc_map, _ = self._staged
        check(self.lib.capdec_train_bwd(self.handle, C.c_float(loss_scale), ptr(inv_ntok), None, None,
                                        C.byref(self.store.c_grads), stream_ptr()), "capdec_train_bwd")
        dmap = torch.empty_like(enc_map) if want_dmap else None
        check(self.lib.capdec_img2feature_bwd(self.handle, ptr(dmap), C.byref(self.adaptor.c_grads), stream_ptr()),
              "capdec_img2feature_bwd")
        return dmap

    def decode_map(self, enc_map, beam: int = 1, finish_on_stop: bool = False, want_f32: bool = True):
        """eval_network after the encoder (model:185-189): _img2feature + beam / greedy decode."""
        self.img2feature(enc_map)
        return self._decode(None, None, self._staged[1], beam, finish_on_stop, want_f32)

    def decode(self, g, P, beam: int = 1, finish_on_stop: bool = False, want_f32: bool = True):
        """capdec_decode.  Returns (ids int32 [B,beam,T], scores f32 [B,beam], ids_f32 [B,T] or None)."""
        B = self._check_inputs(g, P)
        return self._decode(g, P, B, beam, finish_on_stop, want_f32)

    def _decode(self, g, P, B, beam, finish_on_stop, want_f32):
        self._ensure_bound()
        T = self.cfg.infer_max_length
        ids = torch.empty(B, beam, T, dtype=torch.int32, device=self.device)
        scores = torch.empty(B, beam, dtype=torch.float32, device=self.device)
        ids_f32 = torch.empty(B, T, dtype=torch.float32, device=self.device) if want_f32 else None
        check(self.lib.capdec_decode(self.handle, ptr(g), ptr(P), B, beam, 1 if finish_on_stop else 0, ptr(ids),
                                     ptr(scores), ptr(ids_f32), stream_ptr()), "capdec_decode")
        return ids, scores, ids_f32

    def check_status(self):
        """capdec_status: synchronise the current stream and raise CapdecError if a kernel of an earlier call found bad
        input (token ids outside the vocabulary, lengths that cut off non-padding tokens)."""
        check(self.lib.capdec_status(self.handle, stream_ptr()), "capdec_status")

    def decode_trace(self, B: int, beam: int):
        """capdec_decode_trace: (token, back-pointer) of every selection of the last decode, int32 [T,B,beam,2]."""
        out = torch.empty(self.cfg.infer_max_length, B, beam, 2, dtype=torch.int32, device=self.device)
        check(self.lib.capdec_decode_trace(self.handle, ptr(out), stream_ptr()), "capdec_decode_trace")
        return out

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.capdec_workspace_bytes(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self.lib.capdec_launch_count(self.handle))

    N_GRAD_EVENTS = 6

    def grad_group_slices(self):
        """[(event index, lo, hi)]: element ranges of store.flat_grad in the order capdec_train_bwd finalises them
        (include/capdec.h, capdec_set_grad_ready_events); together they cover every parameter exactly once."""
        st, cfg = self.store, self.cfg
        E, H = cfg.embedding_size, cfg.hidden_dim

        def span(first, last):
            return st.offsets[first], st.offsets[last] + (math.prod(st.shapes[last]) + 63) // 64 * 64

        lw = st.offsets["lstm_w"]
        row = 4 * H
        out = [(0,) + span("fc_hs_w", "fc_e_b"),                         # head parameters: final before the BPTT kernel
               (2, lw + (E + H) * row, lw + (E + 2 * H) * row),          # lstm_w hidden-state rows
               (3, lw + E * row, lw + (E + H) * row),                    # lstm_w global-image rows
               (4, st.offsets["lstm_b"], lw + E * row),                  # lstm_b, fc_xs_* and the lstm_w word rows
               (1,) + span("word_embedding", "out_fc_bias"),
               (5,) + span("fc_v_w", "fc_k_b")]                          # per-image projections: last work of the pass
        return out

    def set_grad_ready_events(self, events) -> None:
        """events: N_GRAD_EVENTS recorded torch.cuda.Event, or None to switch the marks off"""
        if events is None:
            check(self.lib.capdec_set_grad_ready_events(self.handle, None, 0), "capdec_set_grad_ready_events")
            self._grad_events = None
            return
        arr = (C.c_void_p * len(events))(*[C.c_void_p(int(e.cuda_event)) for e in events])
        check(self.lib.capdec_set_grad_ready_events(self.handle, arr, len(events)), "capdec_set_grad_ready_events")
        self._grad_events = list(events)   # keep them alive

    def profile(self, enable: bool = True) -> None:
        """Switch the device-side phase timeline on or off (capdec_profile)."""
        check(self.lib.capdec_profile(self.handle, 1 if enable else 0), "capdec_profile")

    def profile_report(self) -> dict:
        """{phase: milliseconds} summed since the previous report (synchronises on the last recorded event)."""
        import ctypes
        buf = ctypes.create_string_buffer(8192)
        check(self.lib.capdec_profile_report(self.handle, buf, len(buf)), "capdec_profile_report")
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms = line.rsplit(" ", 1)
            out[name] = float(ms)
        return out


# --------------------------------------------------------------------------------------------------------------
# reference-shaped interface
# --------------------------------------------------------------------------------------------------------------
_default_engine: Optional[CaptionDecoderEngine] = None


def default_engine(cfg: DecoderConfig = DecoderConfig()) -> CaptionDecoderEngine:
    """The reference keeps its parameters as global named variables of the Paddle scope; the
    equivalent here is one process-wide engine shared by the train- and eval-mode Decoder objects
    (cf. share_vars_from, train.py:127)."""
    global _default_engine
    if _default_engine is None or _default_engine.cfg != cfg:
        _default_engine = CaptionDecoderEngine(cfg)
    return _default_engine


class Decoder:
    """Drop-in for model_adaAttention_aic.Decoder (model:40-135)."""

    def __init__(self, hidden_size, mode="train", rnn_layer=1, engine: Optional[CaptionDecoderEngine] = None):
        if mode not in ["train", "eval"]:
            raise ValueError("不支持{}".format(mode))                                  # model:43-44
        self.engine = engine if engine is not None else default_engine()
        if hidden_size != self.engine.cfg.hidden_dim:
            raise ValueError("hidden_size {} does not match the engine ({})".format(hidden_size,
                                                                                    self.engine.cfg.hidden_dim))
        if rnn_layer != 1:
            raise ValueError("rnn_layer must be 1 (the reference never uses another value)")
        self.hid_size = hidden_size
        self.rnn_layer = rnn_layer
        self.mode = mode
        self.start_tag = self.engine.cfg.start_idx

    def call(self, global_img_feat, p_img_feat, embedding_fn=None, words=None):
        """train: words [B,T] int64 (the source tokens caption[:, :-1]) -> logits [B,T,V] float32;
        eval: -> float32 token ids [B,infer_max_length] (model:122, 132-133).  `embedding_fn` is accepted for
        signature compatibility; the tied `word_embedding` table of the engine is always used (model:28-32)."""
        eng = self.engine
        if self.mode == "eval":
            _, _, ids_f32 = eng.decode(global_img_feat, p_img_feat, beam=1, finish_on_stop=False)
            return ids_f32
        if words is None:
            raise ValueError("train mode needs `words`")
        cfg = eng.cfg
        B = words.shape[0]
        if tuple(words.shape) != (B, cfg.train_steps):
            raise ValueError("words must be [B,{}]".format(cfg.train_steps))
        # the C-ABI takes the full caption [B,S]; only its first S-1 columns are read as source tokens when the
        # logits are requested, so pad one column (its value only matters for the loss, which is not returned)
        caption = torch.cat([words.to(torch.int64), torch.full((B, 1), cfg.padding_idx, dtype=torch.int64,
                                                               device=words.device)], dim=1).contiguous()
        _, logits = eng.train_forward(global_img_feat, p_img_feat, caption, want_logits=True)
        return logits


class _TrainingLoss(torch.autograd.Function):
    """loss = training_network(...) with the hand-written backward behind torch.autograd, so that a PyTorch
    encoder in front and a torch optimizer behind can be used unchanged."""

    @staticmethod
    def forward(ctx, engine, g, P, caption, flat_params):
        loss_buf, _ = engine.train_forward(g, P, caption)
        ctx.engine = engine
        ctx.need = (g.requires_grad, P.requires_grad)
        return loss_buf[2].clone()

    @staticmethod
    def backward(ctx, grad_out):
        eng = ctx.engine
        scale = (eng._loss_buf[3] * grad_out.to(torch.float32)).reshape(1).contiguous()   # no host sync
        dg, dP = eng.train_backward(1.0, inv_ntok=scale, want_dg=True, want_dP=True)
        return None, dg, dP, None, eng.store.flat_grad.clone()


class _TrainingLossFromMap(torch.autograd.Function):
    """loss = training_network(...) from the encoder output on (adaptor + decoder + masked CE); the backward hands the
    gradient of the encoder map back to autograd (config.py:50 encoder_trainable)."""

    @staticmethod
    def forward(ctx, engine, enc_map, caption, flat_params, flat_adaptor):
        loss_buf, _ = engine.train_forward_map(enc_map, caption)
        ctx.engine = engine
        ctx.shape = enc_map.shape
        return loss_buf[2].clone()

    @staticmethod
    def backward(ctx, grad_out):
        eng = ctx.engine
        scale = (eng._loss_buf[3] * grad_out.to(torch.float32)).reshape(1).contiguous()   # no host sync
        dmap = eng.train_backward_map(1.0, inv_ntok=scale, want_dmap=True)
        return None, dmap.view(ctx.shape), None, eng.store.flat_grad.clone(), eng.adaptor.flat_grad.clone()


class ImageCaptionModel:
    """model_adaAttention_aic.ImageCaptionModel (model:138-212) after the MobileNetV2 encoder: the feature adaptor
    `_img2feature` (model:193-199, from the encoder output on), the decoder and the loss wiring.  The encoder itself is
    library CNN work outside the hot path."""

    def __init__(self, engine: Optional[CaptionDecoderEngine] = None):
        self.engine = engine if engine is not None else default_engine()
        self._flat = self.engine.store.flat.requires_grad_(True)

    def build_network(self, mode="train", **kwargs):
        if mode not in ["train", "eval"]:
            raise ValueError("不支持{}".format(mode))                                  # model:154-155
        if mode == "train":
            return self.training_network(**kwargs)
        return self.eval_network(**kwargs)

    def training_network(self, global_image_feat, image_embed, caption):
        """model:161-183: shift, mask, decoder, masked CE -> scalar 'loss' (autograd-enabled)."""
        self.engine.bind_params()
        return _TrainingLoss.apply(self.engine, global_image_feat, image_embed, caption, self._flat)

    def eval_network(self, global_image_feat, image_embed):
        """model:185-189: greedy decode, float32 ids [B,35]."""
        return Decoder(self.engine.cfg.hidden_dim, mode="eval", engine=self.engine).call(global_image_feat, image_embed)

    # ---- the same two networks from the encoder output on (what model:177 / model:186 hand to the decoder) ------
    def _img2feature(self, encoder_map):
        """model:193-199 after `self.encoder.net(img)`: encoder map [B,1280,7,7] or [B,1280,49] ->
        (img_embed [B,49,H], img_global [B,H]) float32; the features also stay staged inside the engine."""
        return self.engine.img2feature(encoder_map, want_outputs=True)

    def training_network_from_map(self, encoder_map, caption):
        """model:161-183 with `img` replaced by the encoder output: scalar loss, autograd-enabled w.r.t. the map."""
        eng = self.engine
        eng.bind_params()
        eng.bind_adaptor()
        if getattr(self, "_flat_adaptor", None) is None:
            self._flat_adaptor = eng.adaptor.flat.requires_grad_(True)
        return _TrainingLossFromMap.apply(eng, encoder_map, caption, self._flat, self._flat_adaptor)

    def eval_network_from_map(self, encoder_map):
        """model:185-189 with `img` replaced by the encoder output: float32 ids [B,35]."""
        return self.engine.decode_map(encoder_map, beam=1, finish_on_stop=False)[2]

    @staticmethod
    def loss(ground_true, prediction):
        """model:205-212: per-position soft-max cross entropy [B,T,1] of given logits (plain torch; the fused
        loss used for training lives inside capdec_train_fwd)."""
        lp = torch.log_softmax(prediction, dim=-1)
        return -torch.gather(lp, -1, ground_true.unsqueeze(-1))


def caption_lengths(caption, padding_idx: int = 0):
    """Number of leading non-padding tokens of every caption [B,S] = [start, w.., stop, pad..]
    (preprocess/ai_challenge_tokenizer.py:81-86), as the HOST int32 array `lengths` of capdec_train_fwd.  `caption` is the
    host copy of the batch (CPU tensor or numpy array) -- the data loader has it before the upload."""
    import numpy as np
    c = caption.numpy() if isinstance(caption, torch.Tensor) else np.asarray(caption)
    nz = c != padding_idx
    # position after the last non-padding token (holes inside a caption are kept)
    S = c.shape[1]
    last = S - np.argmax(nz[:, ::-1], axis=1)
    return np.where(nz.any(axis=1), last, 0).astype(np.int32)


def filter_caption(ids, stop_idx=3, padding_idx=0):
    """evaluate.py:15-25 without the id->word table: cut at <stop>, drop <pad>."""
    out = []
    for i in ids:
        i = int(round(float(i)))
        if i == stop_idx:
            break
        if i == padding_idx:
            continue
        out.append(i)
    return out
