"""CPU, gloo, world_size 2: the multi-GPU plumbing of the hot path without GPUs.

The B200 engine exposes the whole network as ONE autograd node that returns every parameter
gradient at once (cy4/engine.py `_NetFn`).  These tests check, with the same construction on CPU
tensors, that PyTorch DDP averages such gradients correctly over ranks (data parallel,
independent images, one gradient all-reduce per step), that the per-rank synthetic shards differ,
and that bench.py's reference arm prints exactly one JSON line from rank 0."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class _OneNode(torch.autograd.Function):
    """y = sum((x @ W + b)^2) with all parameter gradients produced by a single backward call."""

    @staticmethod
    def forward(ctx, x, W, b):
        h = x @ W + b
        ctx.save_for_backward(x, h)
        return (h * h).sum().reshape(1)

    @staticmethod
    def backward(ctx, g):
        x, h = ctx.saved_tensors
        gh = 2 * h * g
        return None, x.t() @ gh, gh.sum(0)


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.W = torch.nn.Parameter(torch.randn(6, 4))
        self.b = torch.nn.Parameter(torch.zeros(4))

    def forward(self, x):
        return _OneNode.apply(x, self.W, self.b)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
    from cy4 import synth
    torch.manual_seed(0)
    net = _Net()
    ddp = torch.nn.parallel.DistributedDataParallel(net)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(5, 6, generator=g)
    loss = ddp(x)
    loss.backward()
    # expected: mean over ranks of the per-rank gradients
    xs = [torch.randn(5, 6, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    W0, b0 = net.W.detach(), net.b.detach()
    gW = sum(x_.t() @ (2 * (x_ @ W0 + b0)) for x_ in xs) / world
    ok_grad = torch.allclose(net.W.grad, gW, atol=1e-5)
    # per-rank synthetic shards (bench.py uses seed + rank) differ, device-timed max-over-ranks reduce works
    t_a = synth.make_targets(2, per_image=3, seed=4321 + rank)
    sig = torch.tensor([float(t_a[:, 2].sum())])
    gathered = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(gathered, sig)
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        json.dump({"ok_grad": bool(ok_grad), "distinct": bool(abs(gathered[0].item() - gathered[1].item()) > 1e-6),
                   "max_ms": float(ms.item())}, open(out, "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_single_node_function_gloo(tmp_path):
    out = str(tmp_path / "r.json")
    mp.spawn(_worker, args=(2, 29611, out), nprocs=2, join=True)
    r = json.load(open(out))
    assert r["ok_grad"] and r["distinct"] and r["max_ms"] == 11.0


def test_bench_reference_arm_under_torchrun(tmp_path):
    """`bench.py --impl reference` launched like the driver does for N=2: rank 0 prints ONE JSON line
    (the reference's own CPU step on a bounded sample -- the unmodified reference when oracle/_ref or /root/reference is
    present, else the oracle port), the other rank exits 0 without output; `steps` is what really ran."""
    env = dict(os.environ, CY4_BENCH_TEST_TINY="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 0 and d["unit"] == "img/s" and d["steps"] == 1 and d["warmup"] == 0
    assert abs(d["ms_per_step"] * d["value"] / 1e3 - 2.0) < 0.05          # ms_per_step is the bs=2 sample step that was timed
