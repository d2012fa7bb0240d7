"""Import / export of the hot-path parameters in the reference's on-disk format.

The reference checkpoints with ``fluid.io.save_persistables`` / ``save_params`` (train.py:68-91) and resumes with
``load_persistables`` (train.py:94-107): PaddlePaddle 1.8 writes ONE FILE PER VARIABLE, named after the variable, each
holding a serialised LoDTensor (paddle/fluid/framework/lod_tensor.cc SerializeToStream + tensor_util.cc TensorToStream):

    uint32  LoDTensor version (0)
    uint64  number of LoD levels (0 for parameters); per level: uint64 byte size + that many bytes of size_t offsets
    uint32  Tensor version (0)
    int32   byte size of the TensorDesc protobuf
    bytes   VarType.TensorDesc { required Type data_type = 1; repeated int64 dims = 2; }   (proto2, dims not packed)
    bytes   the data, row-major, little endian

Variable names: `word_embedding`, `out_fc_bias`, `lstm_w`, `lstm_b` are explicit (model:18, 23, 87-88);
every other fc is auto-named `fc_N.w_0` / `fc_N.b_0` in construction order under unique_name.guard() (train.py:38):
model:197, 199 (adaptor), 52, 53, the fc inside lstm_unit (explicit parameter names, but it consumes fc_4), 89, 90, 99,
102, 104, 107, 115, 24.  The map is a plain dict (`PADDLE_NAMES`) so a checkpoint with other numbering can be loaded by
passing another one.  Adam state follows Paddle's accumulator names: `<param>_moment1_0`, `<param>_moment2_0`,
`<param>_beta1_pow_acc_0`, `<param>_beta2_pow_acc_0`.

No fixture of the reference exists in its tree: the format is pinned by a byte-level known answer written out
by hand from the description above (tests/test_checkpoint_cpu.py), not by a file produced by Paddle.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Optional

import numpy as np

# VarType.Type (framework.proto)
_DTYPES = {0: np.bool_, 1: np.int16, 2: np.int32, 3: np.int64, 4: np.float16, 5: np.float32, 6: np.float64,
           20: np.uint8, 21: np.int8}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}

PADDLE_NAMES: Dict[str, str] = {
    "fc_img_w": "fc_0.w_0", "fc_img_b": "fc_0.b_0",          # model:197
    "fc_glob_w": "fc_1.w_0", "fc_glob_b": "fc_1.b_0",        # model:199
    "fc_v_w": "fc_2.w_0", "fc_v_b": "fc_2.b_0",              # model:52
    "fc_k_w": "fc_3.w_0", "fc_k_b": "fc_3.b_0",              # model:53
    "lstm_w": "lstm_w", "lstm_b": "lstm_b",                  # model:87-88 (the helper still consumes the name fc_4)
    "fc_xs_w": "fc_5.w_0", "fc_xs_b": "fc_5.b_0",            # model:89
    "fc_hs_w": "fc_6.w_0", "fc_hs_b": "fc_6.b_0",            # model:90
    "fc_p_w": "fc_7.w_0", "fc_p_b": "fc_7.b_0",              # model:99
    "fc_g_w": "fc_8.w_0", "fc_g_b": "fc_8.b_0",              # model:102
    "fc_s_w": "fc_9.w_0", "fc_s_b": "fc_9.b_0",              # model:104
    "fc_a_w": "fc_10.w_0", "fc_a_b": "fc_10.b_0",            # model:107
    "fc_o_w": "fc_11.w_0", "fc_o_b": "fc_11.b_0",            # model:115
    "fc_e_w": "fc_12.w_0", "fc_e_b": "fc_12.b_0",            # model:24
    "word_embedding": "word_embedding",                      # model:18, 31
    "out_fc_bias": "out_fc_bias",                            # model:23
}


def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _read_varint(buf: bytes, pos: int):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            break
        shift += 7
    if val >= 1 << 63:
        val -= 1 << 64
    return val, pos


def encode_tensor_desc(dtype, dims) -> bytes:
    out = b"\x08" + _varint(_DTYPE_IDS[np.dtype(dtype)])
    for d in dims:
        out += b"\x10" + _varint(int(d))
    return out


def decode_tensor_desc(buf: bytes):
    pos, dtype, dims = 0, None, []
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if field == 1 and wire == 0:
            v, pos = _read_varint(buf, pos)
            dtype = _DTYPES[v]
        elif field == 2 and wire == 0:
            v, pos = _read_varint(buf, pos)
            dims.append(v)
        elif field == 2 and wire == 2:                       # packed encoding (written by some proto3 tool chains)
            n, pos = _read_varint(buf, pos)
            end = pos + n
            while pos < end:
                v, pos = _read_varint(buf, pos)
                dims.append(v)
        else:
            raise ValueError("unexpected field {} (wire type {}) in TensorDesc".format(field, wire))
    if dtype is None:
        raise ValueError("TensorDesc without data_type")
    return dtype, dims


def write_lod_tensor(path: str, array) -> None:
    a = np.ascontiguousarray(array)
    desc = encode_tensor_desc(a.dtype, a.shape)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", 0))            # LoDTensor version
        f.write(struct.pack("<Q", 0))            # no LoD levels
        f.write(struct.pack("<I", 0))            # Tensor version
        f.write(struct.pack("<i", len(desc)))
        f.write(desc)
        f.write(a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes())


def read_lod_tensor(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0
    (ver,) = struct.unpack_from("<I", buf, pos); pos += 4
    if ver != 0:
        raise ValueError("{}: unsupported LoDTensor version {}".format(path, ver))
    (levels,) = struct.unpack_from("<Q", buf, pos); pos += 8
    for _ in range(levels):
        (nbytes,) = struct.unpack_from("<Q", buf, pos); pos += 8 + nbytes
    (tver,) = struct.unpack_from("<I", buf, pos); pos += 4
    if tver != 0:
        raise ValueError("{}: unsupported Tensor version {}".format(path, tver))
    (dlen,) = struct.unpack_from("<i", buf, pos); pos += 4
    dtype, dims = decode_tensor_desc(buf[pos:pos + dlen]); pos += dlen
    n = int(np.prod(dims)) if dims else 1
    need = n * np.dtype(dtype).itemsize
    if len(buf) - pos != need:
        raise ValueError("{}: {} data bytes for shape {} {}".format(path, len(buf) - pos, dims, np.dtype(dtype)))
    return np.frombuffer(buf, dtype=np.dtype(dtype).newbyteorder("<"), count=n, offset=pos).reshape(dims).copy()


# ----------------------------------------------------------------------------------------------------------------
# engine <-> directory of variable files
# ----------------------------------------------------------------------------------------------------------------
def _stores(engine):
    stores = [engine.store]
    if getattr(engine, "adaptor", None) is not None:
        stores.append(engine.adaptor)
    return stores


def save_params(dirname: str, engine, names: Optional[Dict[str, str]] = None) -> None:
    """fluid.io.save_params (train.py:78-79) for the decoder (and, when attached, adaptor) parameters."""
    names = names or PADDLE_NAMES
    os.makedirs(dirname, exist_ok=True)
    for st in _stores(engine):
        for n, t in st.params.items():
            write_lod_tensor(os.path.join(dirname, names[n]), t.detach().cpu().numpy())


def load_params(dirname: str, engine, names: Optional[Dict[str, str]] = None, strict: bool = True):
    """fluid.io.load_vars / load_persistables for the hot-path parameters.  Returns the names that were missing
    (strict=False; cf. util.get_predicate, util.py:122-138, which skips variables without a file)."""
    import torch
    names = names or PADDLE_NAMES
    missing = []
    for st in _stores(engine):
        for n, t in st.params.items():
            path = os.path.join(dirname, names[n])
            if not os.path.isfile(path):
                if strict:
                    raise FileNotFoundError("ERROR: %s not found!" % names[n])           # util.py:133
                missing.append(n)
                continue
            a = read_lod_tensor(path)
            if tuple(a.shape) != tuple(t.shape):
                raise ValueError("variable {} has shape {}, expected {}".format(names[n], a.shape, tuple(t.shape)))
            with torch.no_grad():
                t.copy_(torch.from_numpy(a.astype(np.float32)))
    engine._bound = False                # the operand copies inside the handle are stale
    if hasattr(engine, "_adaptor_bound"):
        engine._adaptor_bound = False
    from . import _capi
    for st in _stores(engine):
        _capi.bump_version(st.flat)
    return missing


def save_persistables(dirname: str, engine, optimizer=None, names: Optional[Dict[str, str]] = None) -> None:
    """fluid.io.save_persistables (train.py:73): the parameters plus, per parameter, Adam's accumulators."""
    names = names or PADDLE_NAMES
    save_params(dirname, engine, names)
    if optimizer is None:
        return
    for st, m1, m2 in zip(_stores(engine), optimizer.moment1, optimizer.moment2):
        for n in st.params:
            o, cnt = st.offsets[n], int(np.prod(st.shapes[n]))
            base = names[n]
            write_lod_tensor(os.path.join(dirname, base + "_moment1_0"), m1[o:o + cnt].view(st.shapes[n]).cpu().numpy())
            write_lod_tensor(os.path.join(dirname, base + "_moment2_0"), m2[o:o + cnt].view(st.shapes[n]).cpu().numpy())
            write_lod_tensor(os.path.join(dirname, base + "_beta1_pow_acc_0"), np.array([optimizer.beta1_pow], np.float32))
            write_lod_tensor(os.path.join(dirname, base + "_beta2_pow_acc_0"), np.array([optimizer.beta2_pow], np.float32))
    # The graph's step counter (util.py:48-52, layers.autoincreased_step_counter(begin)) starts at begin - 1 and is
    # incremented BEFORE it is read: after n executor runs it holds n - 1 + begin.  begin is 1 for the two warm-up
    # schedules and 0 for cosine_decay / cosine_decay_restart; with lr_decay_strategy None the variable does not exist.
    sched = getattr(optimizer, "lr", None)
    begin = getattr(sched, "counter_begin", None)
    if begin is not None:
        write_lod_tensor(os.path.join(dirname, "@LR_DECAY_COUNTER@"), np.array([optimizer.steps - 1 + begin], np.int64))
        if sched.strategy == "cosine_decay_restart_warmup":            # the persistable of util.py:92-93
            write_lod_tensor(os.path.join(dirname, "cur_epoch"),
                             np.array([sched.cur_epoch_state(optimizer.steps)], np.float32))


def load_persistables(dirname: str, engine, optimizer=None, names: Optional[Dict[str, str]] = None) -> None:
    """fluid.io.load_persistables (train.py:101-102): resume parameters and optimizer state."""
    import torch
    names = names or PADDLE_NAMES
    load_params(dirname, engine, names, strict=True)
    if optimizer is None:
        return
    first = True
    for st, m1, m2 in zip(_stores(engine), optimizer.moment1, optimizer.moment2):
        for n in st.params:
            o, cnt = st.offsets[n], int(np.prod(st.shapes[n]))
            base = os.path.join(dirname, names[n])
            m1[o:o + cnt].copy_(torch.from_numpy(read_lod_tensor(base + "_moment1_0").reshape(-1)))
            m2[o:o + cnt].copy_(torch.from_numpy(read_lod_tensor(base + "_moment2_0").reshape(-1)))
            if first:
                optimizer.beta1_pow = np.float32(read_lod_tensor(base + "_beta1_pow_acc_0")[0])
                optimizer.beta2_pow = np.float32(read_lod_tensor(base + "_beta2_pow_acc_0")[0])
                first = False
    ctr = os.path.join(dirname, "@LR_DECAY_COUNTER@")
    sched = getattr(optimizer, "lr", None)
    begin = getattr(sched, "counter_begin", None)
    if begin is not None and os.path.isfile(ctr):
        optimizer.steps = int(read_lod_tensor(ctr)[0]) + 1 - begin     # counter = steps - 1 + begin (see save_persistables)
        ce = os.path.join(dirname, "cur_epoch")
        if sched.strategy == "cosine_decay_restart_warmup" and os.path.isfile(ce):
            sched._epoch_counter = float(read_lod_tensor(ce)[0])
            sched._last_n = optimizer.steps - 1
    elif begin is None and float(optimizer.beta1_pow) > 0.0 and hasattr(optimizer, "beta1"):
        # constant learning rate (util.py:44-45): the graph has no step counter; the number of updates taken is what the
        # Beta1Pow accumulator encodes (beta1^(n+1) after n updates) -- only our own bookkeeping needs it
        import math
        optimizer.steps = max(0, int(round(math.log(float(optimizer.beta1_pow)) / math.log(optimizer.beta1))) - 1)
