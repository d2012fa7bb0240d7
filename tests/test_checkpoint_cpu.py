"""CPU: the reference's on-disk parameter format (one serialised LoDTensor per variable, train.py:68-107)."""
import os
import struct
import types

import numpy as np
import pytest
import torch

import myimagecaptioningmodel_b200 as M
from myimagecaptioningmodel_b200 import checkpoint as CK

DIMS = dict(hidden_dim=32, embedding_size=16, vocab_size=53, num_positions=7, sentence_length=7, infer_max_length=7)


def _engine():
    cfg = M.DecoderConfig(**DIMS)
    return types.SimpleNamespace(store=M.ParameterStore(cfg, "cpu").init_reference(1, 0.1),
                                 adaptor=M.AdaptorStore(cfg, 24, "cpu").init_reference(1, 0.1), _bound=True,
                                 _adaptor_bound=True)


def test_lod_tensor_byte_level_known_answer(tmp_path):
    """written out by hand from lod_tensor.cc / tensor_util.cc: a [2,3] float32 tensor"""
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    p = str(tmp_path / "fc_0.w_0")
    CK.write_lod_tensor(p, a)
    want = (struct.pack("<I", 0) + struct.pack("<Q", 0) + struct.pack("<I", 0) + struct.pack("<i", 6)
            + bytes([0x08, 0x05, 0x10, 0x02, 0x10, 0x03])       # data_type = FP32 (5); dims = 2, 3 (unpacked int64)
            + a.tobytes())
    assert open(p, "rb").read() == want
    assert np.array_equal(CK.read_lod_tensor(p), a)
    # a file with one LoD level and packed dims (as other writers produce) is read too
    lod = struct.pack("<Q", 16) + struct.pack("<QQ", 0, 2)
    desc = bytes([0x08, 0x03, 0x12, 0x02, 0x02, 0x01])          # INT64, packed dims [2, 1]
    blob = (struct.pack("<I", 0) + struct.pack("<Q", 1) + lod + struct.pack("<I", 0) + struct.pack("<i", len(desc)) + desc
            + np.array([[7], [9]], dtype=np.int64).tobytes())
    q = str(tmp_path / "ids")
    open(q, "wb").write(blob)
    assert CK.read_lod_tensor(q).tolist() == [[7], [9]]
    open(q, "wb").write(blob[:-3])
    with pytest.raises(ValueError):
        CK.read_lod_tensor(q)
    # large and negative dims survive the varint coding (-1 = Paddle's "batch" dimension in descs)
    assert CK.decode_tensor_desc(CK.encode_tensor_desc(np.float32, [12295, 256, -1])) == (np.float32, [12295, 256, -1])


def test_variable_names_follow_the_construction_order():
    names = CK.PADDLE_NAMES
    assert set(names) == set(M.PARAM_NAMES) | {"fc_img_w", "fc_img_b", "fc_glob_w", "fc_glob_b"}
    assert names["word_embedding"] == "word_embedding" and names["lstm_w"] == "lstm_w"        # model:18, 87
    assert names["fc_img_w"] == "fc_0.w_0" and names["fc_v_w"] == "fc_2.w_0" and names["fc_e_b"] == "fc_12.b_0"
    assert "fc_4.w_0" not in names.values()                      # consumed by the fc inside lstm_unit
    assert len(set(names.values())) == len(names)


def test_save_load_roundtrip_with_optimizer_state(tmp_path):
    eng = _engine()
    opt = types.SimpleNamespace(
        moment1=[torch.randn_like(eng.store.flat), torch.randn_like(eng.adaptor.flat)],
        moment2=[torch.rand_like(eng.store.flat), torch.rand_like(eng.adaptor.flat)],
        beta1_pow=np.float32(0.9 ** 5), beta2_pow=np.float32(0.999 ** 5), steps=4,
        lr=M.get_lr("cosine_decay_warmup", 1e-3, 64, 16))
    d = str(tmp_path / "checkpoint")
    CK.save_persistables(d, eng, opt)
    files = set(os.listdir(d))
    assert {"lstm_w", "lstm_w_moment1_0", "lstm_w_moment2_0", "lstm_w_beta1_pow_acc_0", "fc_0.w_0", "fc_12.b_0",
            "word_embedding", "@LR_DECAY_COUNTER@"} <= files
    assert CK.read_lod_tensor(os.path.join(d, "lstm_w")).shape == (16 + 2 * 32, 4 * 32)
    eng2 = _engine()
    eng2.store.flat.zero_()
    eng2.adaptor.flat.zero_()
    opt2 = types.SimpleNamespace(moment1=[torch.zeros_like(eng.store.flat), torch.zeros_like(eng.adaptor.flat)],
                                 moment2=[torch.zeros_like(eng.store.flat), torch.zeros_like(eng.adaptor.flat)],
                                 beta1_pow=np.float32(0.9), beta2_pow=np.float32(0.999), steps=0,
                                 lr=M.get_lr("cosine_decay_warmup", 1e-3, 64, 16))
    CK.load_persistables(d, eng2, opt2)
    for n in M.PARAM_NAMES:
        assert torch.equal(eng2.store.params[n], eng.store.params[n]), n
    assert torch.equal(eng2.adaptor.flat, eng.adaptor.flat)
    for st, a, b in zip((eng.store, eng.adaptor), opt.moment1, opt2.moment1):
        for n in st.params:                                      # the padding between views is not part of the state
            o, c = st.offsets[n], int(np.prod(st.shapes[n]))
            assert torch.equal(a[o:o + c], b[o:o + c])
    assert opt2.beta1_pow == opt.beta1_pow and opt2.steps == 4
    assert eng2._bound is False and eng2._adaptor_bound is False  # the handle's operand copies must be rebuilt
    # missing files: strict raises like util.get_predicate's message, non-strict reports them
    os.remove(os.path.join(d, "fc_7.w_0"))
    with pytest.raises(FileNotFoundError):
        CK.load_params(d, eng2)
    assert CK.load_params(d, eng2, strict=False) == ["fc_p_w"]
    CK.write_lod_tensor(os.path.join(d, "fc_7.w_0"), np.zeros((3, 3), np.float32))
    with pytest.raises(ValueError):
        CK.load_params(d, eng2)


@pytest.mark.parametrize("strategy,begin", [(None, None), ("cosine_decay", 0), ("cosine_decay_restart", 0),
                                            ("cosine_decay_warmup", 1), ("cosine_decay_restart_warmup", 1)])
def test_lr_decay_counter_follows_the_graphs_begin(tmp_path, strategy, begin):
    """@LR_DECAY_COUNTER@ (util.py:48-52, autoincreased_step_counter(begin)): the variable starts at begin - 1 and is
    incremented before it is read, so after n executor runs it holds n - 1 + begin: n for the two warm-up schedules
    (begin = 1), n - 1 for cosine_decay / cosine_decay_restart (begin = 0); with strategy None it does not exist.  The
    restart-warm-up schedule also carries the persistable `cur_epoch` (util.py:92-99)."""
    eng = _engine()
    n = 9
    sched = M.get_lr(strategy, 1e-3, 64, 16, decay_epoch=2)          # 4 steps per epoch
    mk = lambda steps, s: types.SimpleNamespace(
        moment1=[torch.zeros_like(eng.store.flat), torch.zeros_like(eng.adaptor.flat)],
        moment2=[torch.zeros_like(eng.store.flat), torch.zeros_like(eng.adaptor.flat)],
        beta1_pow=np.float32(0.9), beta2_pow=np.float32(0.999), steps=steps, lr=s)
    d = str(tmp_path / "ck")
    CK.save_persistables(d, eng, mk(n, sched))
    ctr = os.path.join(d, "@LR_DECAY_COUNTER@")
    if begin is None:
        assert not os.path.exists(ctr)
        return
    assert int(CK.read_lod_tensor(ctr)[0]) == n - 1 + begin
    # the graph itself: simulate n runs of the counter op and compare
    var = begin - 1
    for _ in range(n):
        var += 1
    assert var == int(CK.read_lod_tensor(ctr)[0])
    sched2 = M.get_lr(strategy, 1e-3, 64, 16, decay_epoch=2)
    opt2 = mk(0, sched2)
    CK.load_persistables(d, eng, opt2)
    assert opt2.steps == n
    if strategy == "cosine_decay_restart_warmup":
        ce = float(CK.read_lod_tensor(os.path.join(d, "cur_epoch"))[0])
        assert ce == 2.0                                              # global_step 4 and 8 hit `cur_part <= 0`
        # resumed and uninterrupted schedules give the same learning rates from here on
        full = M.get_lr(strategy, 1e-3, 64, 16, decay_epoch=2)
        want = [full(i) for i in range(n + 6)][n:]
        assert [sched2(i) for i in range(n, n + 6)] == want
    else:
        assert not os.path.exists(os.path.join(d, "cur_epoch"))
