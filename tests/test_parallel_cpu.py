"""world_size-2 gloo test of the data-parallel host logic: two ranks, each holding half of the
batch, must reproduce the single-process loss and gradients at the global batch.  The engine behind
DataParallelTrainer is the CPU oracle here (the C-ABI engine needs a GPU); the trainer code is the product's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from myimagecaptioningmodel_b200.parallel import DataParallelTrainer, shard_range
from oracle import decoder_ref as R
from oracle import synth

SMALL = dict(hidden_dim=16, embedding_size=8, vocab_size=37, num_positions=3, sentence_length=6, infer_max_length=6)


class _Store:
    def __init__(self, params):
        self.names = list(params)
        self.sizes = [params[n].numel() for n in self.names]
        self.flat_grad = torch.zeros(sum(self.sizes), dtype=torch.float64)


class OracleEngine:
    """Implements the engine protocol of DataParallelTrainer with the oracle's autograd."""

    def __init__(self, cfg, params):
        self.cfg, self.params = cfg, params
        self.store = _Store(params)
        self._saved = None

    def train_forward(self, g, P, caption):
        ps = {k: v.detach().clone().requires_grad_(True) for k, v in self.params.items()}
        logits, _ = R.decoder_train(ps, self.cfg, g, P, caption[:, :-1])
        loss, loss_sum, ntok = R.masked_loss(self.cfg, logits, caption)
        self._saved = (ps, loss_sum, ntok)
        return torch.stack([loss_sum.detach(), ntok.detach(), loss.detach(), 1.0 / ntok.detach()]), None

    def train_backward(self, loss_scale=1.0, inv_ntok=None, want_dg=False, want_dP=False):
        ps, loss_sum, ntok = self._saved
        inv = (1.0 / ntok) if inv_ntok is None else inv_ntok.reshape(())
        (loss_sum * inv * loss_scale).backward()
        flat = torch.cat([(ps[n].grad if ps[n].grad is not None else torch.zeros_like(ps[n])).reshape(-1)
                          for n in self.store.names])
        self.store.flat_grad.copy_(flat)
        return None, None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, pe, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = R.DecoderConfig(softmax_axis="position", **SMALL)
        params = R.init_params(cfg, 2, torch.float64)
        g, P, cap = synth.make_batch(cfg, 6, 9, torch.float64)
        lo, hi = shard_range(6, rank, world)
        eng = OracleEngine(cfg, params)
        tr = DataParallelTrainer(eng, pe_semantics=pe)
        loss = tr.forward_backward(g[lo:hi], P[lo:hi], cap[lo:hi])
        if rank == 0:
            torch.save({"loss": loss, "grad": eng.store.flat_grad.clone()}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("pe", [False, True])
def test_two_ranks_match_single_process(tmp_path, pe):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), pe, out), nprocs=2, join=True)
    got = torch.load(out)
    cfg = R.DecoderConfig(softmax_axis="position", **SMALL)
    params = R.init_params(cfg, 2, torch.float64)
    g, P, cap = synth.make_batch(cfg, 6, 9, torch.float64)
    if not pe:
        # global-token normalisation == the single-device reference at the global batch (model:169, 182)
        loss, _, _, grads = R.loss_and_grads(params, cfg, g, P, cap)
    else:
        # ParallelExecutor default: mean over devices of the per-device token means
        parts = [R.loss_and_grads(params, cfg, g[a:b], P[a:b], cap[a:b]) for a, b in (shard_range(6, 0, 2),
                                                                                     shard_range(6, 1, 2))]
        loss = (parts[0][0] + parts[1][0]) / 2
        grads = {k: (parts[0][3][k] + parts[1][3][k]) / 2 for k in params}
    flat = torch.cat([grads[n].reshape(-1) for n in params])
    assert abs(float(got["loss"]) - float(loss)) < 1e-12
    assert float((got["grad"] - flat).abs().max()) < 1e-12


# ---- the same through the feature adaptor (forward_backward_map) ------------------------------------------------------
class OracleMapEngine(OracleEngine):
    """engine protocol of DataParallelTrainer.forward_backward_map: adaptor + decoder, oracle autograd"""

    def __init__(self, cfg, params, ap):
        super().__init__(cfg, params)
        self.ap = ap
        self.adaptor = _Store(ap)

    def train_forward_map(self, enc_map, caption):
        from oracle import train_ref as TR
        ps = {k: v.detach().clone().requires_grad_(True) for k, v in self.params.items()}
        aps = {k: v.detach().clone().requires_grad_(True) for k, v in self.ap.items()}
        x = enc_map.detach().clone().requires_grad_(True)
        P, g = TR.img2feature(aps, x)
        logits, _ = R.decoder_train(ps, self.cfg, g, P, caption[:, :-1])
        loss, loss_sum, ntok = R.masked_loss(self.cfg, logits, caption)
        self._saved = (ps, aps, x, loss_sum, ntok)
        return torch.stack([loss_sum.detach(), ntok.detach(), loss.detach(), 1.0 / ntok.detach()]), None

    def train_backward_map(self, loss_scale=1.0, inv_ntok=None, want_dmap=True):
        ps, aps, x, loss_sum, ntok = self._saved
        inv = (1.0 / ntok) if inv_ntok is None else inv_ntok.reshape(())
        (loss_sum * inv * loss_scale).backward()
        z = lambda d, n: (d[n].grad if d[n].grad is not None else torch.zeros_like(d[n])).reshape(-1)
        self.store.flat_grad.copy_(torch.cat([z(ps, n) for n in self.store.names]))
        self.adaptor.flat_grad.copy_(torch.cat([z(aps, n) for n in self.adaptor.names]))
        return x.grad if want_dmap else None


def _worker_map(rank, world, port, out):
    from oracle import train_ref as TR
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = R.DecoderConfig(softmax_axis="position", **SMALL)
        params = R.init_params(cfg, 2, torch.float64)
        ap = TR.init_adaptor_params(cfg, 8, 2, torch.float64)
        x = TR.make_encoder_map(cfg, 6, 8, 9, torch.float64)
        cap = synth.make_captions(cfg, 6, 9)
        lo, hi = shard_range(6, rank, world)
        eng = OracleMapEngine(cfg, params, ap)
        tr = DataParallelTrainer(eng)
        loss, dmap = tr.forward_backward_map(x[lo:hi], cap[lo:hi], want_dmap=True)
        torch.save({"loss": loss, "grad": eng.store.flat_grad.clone(), "agrad": eng.adaptor.flat_grad.clone(),
                    "dmap": dmap, "span": (lo, hi)}, out + str(rank))
    finally:
        dist.destroy_process_group()


def test_two_ranks_match_single_process_from_encoder_map(tmp_path):
    from oracle import train_ref as TR
    out = str(tmp_path / "dpmap.pt")
    mp.spawn(_worker_map, args=(2, _free_port(), out), nprocs=2, join=True)
    cfg = R.DecoderConfig(softmax_axis="position", **SMALL)
    params = R.init_params(cfg, 2, torch.float64)
    ap = TR.init_adaptor_params(cfg, 8, 2, torch.float64)
    x = TR.make_encoder_map(cfg, 6, 8, 9, torch.float64)
    cap = synth.make_captions(cfg, 6, 9)
    loss, _, grads, _ = TR.loss_and_grads_from_map(params, ap, cfg, x, cap)
    for r in range(2):
        got = torch.load(out + str(r))
        assert abs(float(got["loss"]) - float(loss)) < 1e-12
        assert float((got["grad"] - torch.cat([grads[n].reshape(-1) for n in params])).abs().max()) < 1e-12
        assert float((got["agrad"] - torch.cat([grads[n].reshape(-1) for n in ap])).abs().max()) < 1e-12
        lo, hi = got["span"]
        assert float((got["dmap"] - grads["enc_map"][lo:hi]).abs().max()) < 1e-12     # normalised by the GLOBAL count
