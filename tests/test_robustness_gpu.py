"""-m gpu: host-layer guarantees around the C-ABI:
  * the handle's operand copies are re-bound after every optimizer update without the caller's help;
  * token ids outside the vocabulary never index out of bounds: they read as padding and are reported
    (PaddlePaddle's lookup_table_v2 / softmax_with_cross_entropy enforce the same range);
  * the loss a trainer returns is not a view of the engine's live buffer;
  * data-parallel result == single-process result at the same global batch on real GPUs,
    whenever two GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

import myimagecaptioningmodel_b200 as M
from gpu_util import rel_err
from myimagecaptioningmodel_b200 import _capi
from oracle import decoder_ref as R
from oracle import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MID = dict(hidden_dim=256, embedding_size=64, vocab_size=1000, num_positions=49, sentence_length=12, infer_max_length=12)


@pytest.mark.parametrize("from_map", [False, True])
def test_optimizer_updates_rebind_the_handle(from_map):
    """Three forward_backward(optimizer=...) steps WITHOUT a manual bind must equal three steps WITH a bind before each:
    train.py:138-139 simply runs the executor in a loop; stale bf16 weight copies would mix updated biases with
    old weights silently."""
    cfg = M.DecoderConfig(dtype="bf16", softmax_axis="position", **MID)
    g, P, cap = synth.make_batch(cfg, 12, 5)
    x = M.synth.make_encoder_map(cfg, 12, 64, 6).cuda()
    gd, Pd, capd = g.cuda(), P.cuda(), cap.cuda()
    losses = {}
    for manual in (False, True):
        eng = M.CaptionDecoderEngine(cfg, seed=4)
        if from_map:
            eng.attach_adaptor(64, seed=4)
        opt, _ = M.get_optimizer(eng, {"learning_rate": 1e-3})
        tr = M.DataParallelTrainer(eng)
        out = []
        for _ in range(3):
            if manual:
                eng.bind_params()
                if from_map:
                    eng.bind_adaptor()
            loss = tr.forward_backward_map(x, capd, optimizer=opt)[0] if from_map else \
                tr.forward_backward(gd, Pd, capd, optimizer=opt)
            out.append(loss)                      # device tensors kept across steps (deferred logging)
        torch.cuda.synchronize()
        losses[manual] = [float(v) for v in out]
        assert opt.steps == 3
    a, b = losses[False], losses[True]
    assert a[0] > a[1] > a[2], a                  # the parameters the forward pass sees really change
    # ... and as with a bind before every step (fp32 atomics make the sums order-dependent in the last bits)
    assert all(abs(x - y) < 1e-4 * abs(y) for x, y in zip(a, b)), (a, b)
    # Adam driven directly (no trainer) invalidates the binding too
    eng = M.CaptionDecoderEngine(cfg, seed=4)
    opt = M.Adam([(eng.store.flat, eng.store.flat_grad)], 1e-3)
    l0 = float(eng.train_forward(gd, Pd, capd)[0][2])
    eng.train_backward()
    opt.step()
    l1 = float(eng.train_forward(gd, Pd, capd)[0][2])
    assert l1 < l0
    # a store.load() after the first bind is picked up as well
    eng.store.load({k: v for k, v in M.CaptionDecoderEngine(cfg, seed=4).store.state_dict().items()})
    assert abs(float(eng.train_forward(gd, Pd, capd)[0][2]) - l0) < 1e-5 * l0


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_token_ids_outside_the_vocabulary_are_reported_not_dereferenced(dtype):
    ocfg = R.DecoderConfig(softmax_axis="position", **MID)
    params = R.init_params(ocfg, 7, torch.float32, bias_scale=0.1)
    g, P, cap = synth.make_batch(ocfg, 6, 3)
    eng = M.CaptionDecoderEngine(M.DecoderConfig(dtype=dtype, softmax_axis="position", **MID), seed=None)
    eng.store.load(params)
    gd, Pd = g.cuda(), P.cuda()
    # clean input: no report
    eng.train_forward(gd, Pd, cap.cuda())
    eng.train_backward()
    eng.check_status()
    ref_grad = eng.store.flat_grad.clone()
    # ids past the table (a vocabulary that does not match the checkpoint) and a negative id
    bad = cap.clone()
    bad[1, 2] = ocfg.vocab_size          # read as source token at step 2 and as target at step 1
    bad[4, 3] = -5
    eng.train_forward(gd, Pd, bad.cuda())
    with pytest.raises(_capi.CapdecError) as ei:
        eng.check_status()                # synchronises; without it the NEXT call on the handle reports (below)
    assert "vocab_size" in str(ei.value) and "target" in str(ei.value)
    eng.train_backward()
    eng.check_status()
    assert torch.isfinite(eng.store.flat_grad).all()
    # the offending entries were treated as padding: same result as the caption with those entries zeroed
    asp = cap.clone()
    asp[1, 2] = 0
    asp[4, 3] = 0
    eng.check_status()                    # the report was cleared by being returned
    lb = eng.train_forward(gd, Pd, bad.cuda())[0].clone()
    torch.cuda.synchronize()
    with pytest.raises(_capi.CapdecError):        # ... and the NEXT call on the handle reports it without a status call
        eng.train_backward()
    eng.train_backward()                          # reported once; the step itself is intact
    torch.cuda.synchronize()
    gb = eng.store.flat_grad.clone()
    la = eng.train_forward(gd, Pd, asp.cuda())[0].clone()
    eng.train_backward()
    eng.check_status()
    assert torch.equal(lb, la) or rel_err(lb, la) < 1e-6
    assert rel_err(gb, eng.store.flat_grad) < 1e-5
    assert rel_err(gb, ref_grad) > 1e-4           # (the batch did change)


def test_trainer_loss_is_not_a_view_of_the_live_buffer():
    cfg = M.DecoderConfig(dtype="bf16", softmax_axis="position", **MID)
    eng = M.CaptionDecoderEngine(cfg, seed=1)
    tr = M.DataParallelTrainer(eng)
    a = tr.forward_backward(*[t.cuda() for t in synth.make_batch(cfg, 4, 1)])
    va = float(a)
    b = tr.forward_backward(*[t.cuda() for t in synth.make_batch(cfg, 4, 2)])
    assert float(a) == va and float(b) != va


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("dtype,exchange", [("fp32", "fp32"), ("bf16", "fp32"), ("bf16", "bf16")])
def test_two_gpu_data_parallel_equals_single_process(dtype, exchange):
    """scripts/dp_check.py under torch.distributed.run on 2 ranks (NCCL): all-reduced shard gradients == a local
    full-batch step (fp32: every tensor to 1e-4; bf16: the flat gradient to 1e-2), features and encoder-map entry points;
    gradient ranges all-reduced in fp32 or rounded to bf16 for the wire."""
    env = dict(os.environ, DP_CHECK_DTYPE=dtype, DP_CHECK_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", {"fp32fp32": "29541", "bf16fp32": "29542", "bf16bf16": "29543"}[dtype + exchange],
                        os.path.join(ROOT, "scripts", "dp_check.py")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DP CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
