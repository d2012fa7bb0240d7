"""CPU: the bench.py contract that can be checked without a GPU -- the reference arm (`--impl reference`) prints ONE JSON
line with the keys the driver reads, and rank != 0 exits without work; our arm refuses to run without a GPU (no fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--reference-batch", "96"])   # > one autograd chunk
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decoder_train_tokens_per_sec" and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"].startswith("teacher_forced_train_step_B256")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "Adam" in cb["sample"]
    assert "batch 96" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1"], env={"RANK": "1", "WORLD_SIZE": "2"}, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_our_arm_needs_a_gpu():
    r = _run(["--steps", "1", "--warmup", "3", "--no-cpu-baseline"], timeout=300)
    assert r.returncode != 0                                  # no CPU fallback: the product path fails loudly
    assert "cuda" in (r.stderr + r.stdout).lower() or "gpu" in (r.stderr + r.stdout).lower()


def _count_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    # local step times that straddle an integer boundary of budget / ms: 5500 / 2.6000 = 2115.4, 5500 / 2.5995 = 2115.8 ..
    # and one pair that does not round alike locally: 5500 / 2.59990 = 2115.47, 5500 / 2.59930 = 2115.96 vs 2116.04
    counts = [bench.agreed_step_count(5500.0, ms[rank], world, "cpu") for ms in ((2.6000, 2.5995), (2.59925, 2.59935), (3.0, 2.0))]
    local = [int(5500.0 / ms[rank]) + 1 for ms in ((2.6000, 2.5995), (2.59925, 2.59935), (3.0, 2.0))]
    if rank == 0:
        gathered = [None, None]
    else:
        gathered = None
    dist.gather_object((counts, local), gathered, dst=0)
    if rank == 0:
        import torch
        torch.save(gathered, out)
    dist.destroy_process_group()


def test_sustained_leg_step_count_is_agreed_over_ranks(tmp_path):
    """Every data-parallel step is a set of collectives: ranks that derive the number of steps of a timed leg from their
    LOCAL step time can disagree by one and then wait for each other forever (the intermittent multi-GPU bench hang of
    round 2).  bench.agreed_step_count must return the same count on every rank -- the slowest rank's."""
    import socket

    import torch
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "counts.pt")
    mp.spawn(_count_worker, args=(2, port, out), nprocs=2, join=True)
    (c0, l0), (c1, l1) = torch.load(out, weights_only=False)
    assert c0 == c1                                   # agreed ...
    assert l0 != l1                                   # ... where the local counts would not have been
    assert c0 == [min(a, b) for a, b in zip(l0, l1)]  # the slowest rank decides
