"""CPU, world_size 2 over gloo: the batch x time sharding used for N > 1 GPUs.

The lift has no parameters and no cross-frame dependency (fiery/models/fiery.py:231), so the multi-GPU path is:
rank r lifts frames shard_frames(B', N, r); no collective on the data path; a training step all-reduces (averages) the
gradients of whatever produced the head tensor, exactly once.  On this CPU box the per-rank compute is the oracle; the
host-side logic under test (sharding, gather order, gradient averaging) is the same code path bench.py --gpus N uses.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fiery_b200.synthetic import CONFIGS, LiftConfig, make_calibration, make_grad_bev, make_head, shard_frames
from oracle import lift_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(frames):
    cfg = LiftConfig(**{**CONFIGS["cfg1_tiny"].__dict__, "frames": frames})
    K, E = make_calibration(cfg, seed=11)
    return cfg, torch.from_numpy(K), torch.from_numpy(E), torch.from_numpy(make_head(cfg, seed=11)), \
        torch.from_numpy(make_grad_bev(cfg, seed=11))


def _worker(rank, world, port, frames, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg, K, E, head, gout = _inputs(frames)
    mine = shard_frames(frames, world, rank)
    n = cfg.n_cameras
    sub = LiftConfig(**{**cfg.__dict__, "frames": len(mine)})
    oracle = O.LiftOracle.from_config(sub)
    sl = slice(mine.start, mine.stop)
    # a stand-in for Encoder.depth_layer (1x1 conv, encoder.py:36): the only parameters upstream of the lift
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(cfg.head_channels, cfg.head_channels) * 0.1)
    feats = head[sl.start * n:sl.stop * n]
    h = torch.einsum("oc,bchw->bohw", w, feats)
    bev = oracle.lift(h, K[sl], E[sl])
    loss = (bev * gout[sl]).sum() / frames                       # mean over the GLOBAL batch
    loss.backward()
    grad = w.grad.clone()
    dist.all_reduce(grad, op=dist.ReduceOp.SUM)                  # the single gradient all-reduce of the step
    gathered = [torch.zeros_like(bev) for _ in range(world)] if len(mine) * world == frames else None
    if gathered is not None:
        dist.all_gather(gathered, bev.detach())
    dist.barrier()
    if rank == 0:
        # numpy arrays travel by value; torch tensors would be shared through the producer's fd server, which dies with it
        out_q.put((grad.numpy(), torch.cat(gathered).numpy() if gathered is not None else None))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_matches_single_process():
    frames, world = 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    grad, bev = q.get(timeout=240)
    grad, bev = torch.from_numpy(grad), torch.from_numpy(bev)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process reference on the whole batch
    cfg, K, E, head, gout = _inputs(frames)
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(cfg.head_channels, cfg.head_channels) * 0.1)
    full = O.LiftOracle.from_config(cfg).lift(torch.einsum("oc,bchw->bohw", w, head), K, E)
    ((full * gout).sum() / frames).backward()
    assert torch.allclose(bev, full.detach(), rtol=1e-5, atol=1e-6)      # rank-major gather == batch order
    assert torch.allclose(grad, w.grad, rtol=1e-4, atol=1e-6)            # summed shard grads == full-batch grad


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
