"""Replaces the reference's src/utils/torch_utils.py (:16-29)."""
import torch

__all__ = ['convert2cpu', 'convert2cpu_long', 'to_cpu']


def convert2cpu(gpu_matrix):
    return gpu_matrix.detach().to("cpu", torch.float32)


def convert2cpu_long(gpu_matrix):
    return gpu_matrix.detach().to("cpu", torch.int64)


def to_cpu(tensor):
    return tensor.detach().cpu()
