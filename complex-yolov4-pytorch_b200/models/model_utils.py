"""Replaces the reference's src/models/model_utils.py: create_model / get_num_parameters /
make_data_parallel with identical behaviour (reference :20-67); DDP stays PyTorch DDP over NCCL."""
import torch

from cy4.darknet import Darknet


def create_model(configs):
    """Create model based on architecture name (reference :20-28)."""
    if (configs.arch == 'darknet') and (configs.cfgfile is not None):
        print('using darknet')
        model = Darknet(cfgfile=configs.cfgfile, use_giou_loss=configs.use_giou_loss)
    else:
        assert False, 'Undefined model backbone'
    return model


def get_num_parameters(model):
    """Count number of trained parameters of the model (reference :31-38)."""
    m = model.module if hasattr(model, 'module') else model
    return sum(p.numel() for p in m.parameters() if p.requires_grad)


def overlap_gradient_exchange(ddp_model):
    """The B200 engine exposes the whole network as one autograd node, so stock DDP would start its bucketed all-reduce only
    after the last backward kernel -- and it pays ~650 tiny copy / scale kernels per step to move the 327 freshly returned
    gradient tensors into its buckets (+2 ms per step, measured with a single rank: profiles/r2_ddp_host_overhead.md).
    Instead the engine averages the gradients itself while backward is still running (cy4/engine.py: grouped asynchronous NCCL
    all-reduce, heavy layers first) and DDP keeps what else it does -- parameter broadcast at construction, buffer broadcast
    every forward -- with its reducer switched off for good (the state `with ddp.no_sync():` sets for one block)."""
    ddp_model.module.engine_allreduce = True
    ddp_model.require_backward_grad_sync = False
    return ddp_model


def make_data_parallel(model, configs):
    """Reference :41-67: DistributedDataParallel (one process per GPU), single GPU, or DataParallel."""
    if configs.distributed:
        if configs.gpu_idx is not None:
            torch.cuda.set_device(configs.gpu_idx)
            model.cuda(configs.gpu_idx)
            configs.batch_size = int(configs.batch_size / configs.ngpus_per_node)
            configs.num_workers = int((configs.num_workers + configs.ngpus_per_node - 1) / configs.ngpus_per_node)
            model = overlap_gradient_exchange(torch.nn.parallel.DistributedDataParallel(model, device_ids=[configs.gpu_idx]))
        else:
            model.cuda()
            model = overlap_gradient_exchange(torch.nn.parallel.DistributedDataParallel(model))
    elif configs.gpu_idx is not None:
        torch.cuda.set_device(configs.gpu_idx)
        model = model.cuda(configs.gpu_idx)
    else:
        model = torch.nn.DataParallel(model).cuda()
    return model
