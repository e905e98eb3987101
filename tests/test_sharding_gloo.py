"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path.

Frames are independent (fiery/models/fiery.py:231): rank r trains on the samples ``rank_shard`` gives it, there is no collective
on the lift's data path, and a training step averages the gradients exactly once -- ``FlatGradBucket.all_reduce_mean``
(fiery_b200/train.py), ONE all-reduce of one flat buffer.  Here the product's bucket, sharding and synthetic batch run over gloo on
CPU tensors; the lift itself has no CPU path, so between the product's ``depth_layer`` and ``BevHead`` the oracle's lift stands in
(the CUDA lift in the same harness is covered by tests/test_train_gpu.py).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fiery_b200.synthetic import CONFIGS, LiftConfig
from fiery_b200.train import BevHead, FlatGradBucket, StandInEncoder, rank_shard, synthetic_batch
from oracle import lift_oracle as O

CFG = LiftConfig(**{**CONFIGS["cfg1_tiny"].__dict__})
GLOBAL_BATCH, SEQ = 4, 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _CpuStandIn(torch.nn.Module):
    """depth_layer -> (oracle lift) -> BevHead: the product's modules around the oracle's CPU lift."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.encoder = StandInEncoder(CFG.depth_bins, CFG.out_channels)
        self.head = BevHead(CFG.out_channels, width=8)

    def forward(self, batch):
        b, s, n = batch["image"].shape[:3]
        feats = batch["image"].reshape(b * s * n, *batch["image"].shape[3:])
        head = self.encoder.depth_layer(feats)
        oracle = O.LiftOracle.from_config(LiftConfig(**{**CFG.__dict__, "frames": b * s}))
        bev = oracle.lift(head, batch["intrinsics"].reshape(b * s, n, 3, 3), batch["extrinsics"].reshape(b * s, n, 4, 4))
        out = self.head(bev)
        return (out["instance_center"] - batch["centerness"].reshape(b * s, 1, *bev.shape[2:])).pow(2).mean()


def _grads(first, count):
    model = _CpuStandIn()
    bucket = FlatGradBucket(model.parameters())
    batch = synthetic_batch(CFG, count, SEQ, torch.device("cpu"), seed=5, feature_input=True, first_sample=first)
    model(batch).backward()
    return bucket, batch


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    first, count = rank_shard(GLOBAL_BATCH, world, rank)
    bucket, batch = _grads(first, count)
    n_calls = {"n": 0}
    real = dist.all_reduce

    def counting(*a, **k):
        n_calls["n"] += 1
        return real(*a, **k)
    dist.all_reduce = counting
    bucket.all_reduce_mean()                                     # the step's only collective
    dist.all_reduce = real
    dist.barrier()
    if rank == 0:
        # numpy arrays travel by value; torch tensors would be shared through the producer's fd server, which dies with it
        out_q.put((bucket.flat.numpy().copy(), n_calls["n"], batch["image"].numpy().copy()))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_flat_bucket_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat, calls, image0 = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert calls == 1                                            # ONE all-reduce for all parameters
    # single process on the whole global batch: mean over 4 samples == mean of the two ranks' means over 2 samples each
    bucket, batch = _grads(0, GLOBAL_BATCH)
    assert torch.allclose(torch.from_numpy(flat), bucket.flat, rtol=1e-4, atol=1e-7)
    assert float(bucket.flat.abs().max()) > 0
    # a rank's shard is the same rows of the global batch
    first, count = rank_shard(GLOBAL_BATCH, world, 0)
    assert torch.equal(torch.from_numpy(image0), batch["image"][first:first + count])
    assert [rank_shard(7, 3, r) for r in range(3)] == [(0, 3), (3, 2), (5, 2)]


def test_reference_arm_non_zero_ranks_exit_quietly():
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_line_keeps_the_bench_contract():
    """`bench.py --impl reference` (rank 0): ONE JSON line with the contract's keys, timed on the CPU port of the reference."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "cfg2_static_lss_b8"
