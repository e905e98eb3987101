"""GPU: the data-parallel training step around the fused CUDA lift (fiery_b200/train.py; reference: train.py:33-46,
fiery/trainer.py:69-120,200-208,254-260).  Single process: the step trains (loss falls, all parameters receive gradients through
the lift's backward, AMP on).  Two ranks over NCCL (needs 2 GPUs; skipped otherwise): ranks keep identical weights, and two ranks
on half the batch each reproduce one process on the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch

from fiery_b200.synthetic import CONFIGS, LiftConfig
from fiery_b200.train import LiftTrainer, rank_shard, synthetic_batch

pytestmark = pytest.mark.gpu
CFG = LiftConfig(**{**CONFIGS["cfg1_tiny"].__dict__})


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("precision,feature_input", [(16, False), (32, True)])
def test_single_process_step_trains(precision, feature_input):
    dev = torch.device("cuda:0")
    tr = LiftTrainer(CFG, dev, precision=precision, feature_input=feature_input, seed=3)
    batch = synthetic_batch(CFG, 2, 2, dev, seed=9, feature_input=feature_input)
    before = {k: v.detach().clone() for k, v in tr.model.named_parameters() if v.requires_grad}
    losses = [float(tr.step(batch)) for _ in range(12)]
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses                 # finite, and it learns this batch
    moved = [k for k, v in tr.model.named_parameters() if v.requires_grad and not torch.equal(v.detach(), before[k])]
    assert any(k.startswith("encoder.depth_layer") for k in moved)                        # gradients crossed the lift's backward
    if not feature_input:
        assert any(k.startswith("encoder.features") for k in moved)
    assert tr.bucket.nbytes == 4 * sum(p.numel() for p in tr.model.parameters() if p.requires_grad)
    out = tr.model(batch["image"], batch["intrinsics"], batch["extrinsics"])
    assert tuple(out["segmentation"].shape) == (2, 2, 2, *CFG.bev_hw)


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    tr = LiftTrainer(CFG, dev, precision=32, feature_input=True, seed=3)
    first, count = rank_shard(4, world, rank)
    batch = synthetic_batch(CFG, count, 1, dev, seed=21, feature_input=True, first_sample=first)
    tr.forward_backward(batch)                       # gradient of this rank's shard, averaged over the ranks by ONE all-reduce
    grad = tr.bucket.flat.clone()
    for _ in range(3):
        tr.step(batch)
    flat = torch.cat([p.detach().flatten() for p in tr.model.parameters() if p.requires_grad])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        q.put((grad.cpu().numpy(), [g.cpu().numpy() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_over_nccl_match_one_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    grad, (w0, w1) = q.get(timeout=500)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert np.array_equal(w0, w1)                                                         # ranks stay in lockstep after 3 steps
    # the averaged gradient of two half batches == the gradient of one process on the whole batch (Adam would amplify rounding
    # noise of near-zero gradients, so the comparison is on the gradient, not on the weights)
    dev = torch.device("cuda:0")
    tr = LiftTrainer(CFG, dev, precision=32, feature_input=True, seed=3)
    tr.forward_backward(synthetic_batch(CFG, 4, 1, dev, seed=21, feature_input=True))
    whole = tr.bucket.flat.cpu()
    assert float(whole.abs().max()) > 0
    assert torch.allclose(torch.from_numpy(grad), whole, rtol=1e-3, atol=1e-5 * float(whole.abs().max()))
