"""GPU, 2 ranks over NCCL: the reference's DDP wrap (models.model_utils.make_data_parallel,
reference src/models/model_utils.py:41-59) around the engine.  Checks on hardware that
  * the rank-averaged gradients of a 2-rank step equal the mean of the two single-GPU gradients of the same shards
    (eval-mode BatchNorm, so the per-rank statistics cannot differ), for the engine's overlapped exchange AND for stock DDP
    bucketing of the engine's one-node autograd output;
  * both ranks hold identical gradients afterwards, NCCL reports 2 ranks.
Skipped with fewer than 2 GPUs (run it with `gpurun --gpus 2`; the log of this round's run is profiles/r2_ddp_2gpu_test.log)."""
import json
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _shard(rank, size=224, B=2):
    sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
    from cy4 import synth
    x = synth.make_bev(B, img_size=size, seed=50 + rank)
    tg = torch.tensor(synth.make_targets(B, per_image=3, seed=60 + rank, img_size=size, strides=(16, 32)))
    return x, tg


def _make_model(dev):
    from cy4 import netdefs
    from cy4.darknet import Darknet
    torch.manual_seed(7)
    model = Darknet(netdefs.cfg_path("complex_yolov4_tiny"), True)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
    return model.to(dev).eval()              # eval-mode BN: running statistics on every rank; gradients still flow


def _worker(rank, world, port, out, stock):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from types import SimpleNamespace
    from models.model_utils import make_data_parallel
    model = _make_model(dev)
    if stock:
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank])
    else:
        ddp = make_data_parallel(model, SimpleNamespace(distributed=True, gpu_idx=rank, batch_size=4, ngpus_per_node=world, num_workers=0))
        assert model.engine_allreduce
    x, tg = _shard(rank)
    for _ in range(2):                        # second step: DDP has rebuilt its buckets, grads are bucket views
        ddp.zero_grad(set_to_none=True)
        loss, _ = ddp(x.to(dev), tg.to(dev))
        loss.backward()
    torch.cuda.synchronize()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save({"grad": flat.cpu(), "same_on_all_ranks": bool(all(torch.equal(g, gathered[0]) for g in gathered)),
                    "world": dist.get_world_size(), "backend": dist.get_backend()}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("stock", [False, True])
def test_two_rank_gradients_equal_mean_of_single_gpu_gradients(tmp_path, stock):
    out = str(tmp_path / "ddp.pt")
    mp.spawn(_worker, args=(2, 29631 + int(stock), out, stock), nprocs=2, join=True)
    r = torch.load(out)
    assert r["world"] == 2 and r["backend"] == "nccl" and r["same_on_all_ranks"]
    # the same two shards, one after the other, on one GPU
    dev = torch.device("cuda", 0)
    model = _make_model(dev)
    singles = []
    for rank in range(2):
        x, tg = _shard(rank)
        model.zero_grad(set_to_none=True)
        loss, _ = model(x.to(dev), tg.to(dev))
        loss.backward()
        singles.append(torch.cat([p.grad.reshape(-1) for p in model.parameters()]).cpu())
    want = (singles[0] + singles[1]) / 2
    got = r["grad"]
    err = (got - want).abs().max().item()
    print("stock" if stock else "engine-overlapped", "max |ddp - mean of singles|", err, "max |want|", want.abs().max().item(),
          "shards differ by", (singles[0] - singles[1]).abs().max().item())
    assert (singles[0] - singles[1]).abs().max().item() > 1e-3 * want.abs().max().item()      # the shards really differ
    assert err <= 2e-3 * want.abs().max().item() + 1e-7                                       # fp32 atomics order only
