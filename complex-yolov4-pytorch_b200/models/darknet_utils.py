"""Replaces the reference's src/models/darknet_utils.py: cfg parser / printer and darknet
`.weights` (de)serialisation helpers (reference :17-47, :50-196, :199-261)."""
import torch

from cy4.darknet import parse_cfg, print_cfg  # noqa: F401

__all__ = ['parse_cfg', 'print_cfg', 'load_conv', 'load_conv_bn', 'save_conv', 'save_conv_bn', 'load_fc', 'save_fc']


def _take(buf, start, t):
    n = t.numel()
    with torch.no_grad():         # in-place copy that bumps t._version: the engine re-packs its fp16 weights
        t.copy_(torch.from_numpy(buf[start:start + n]).reshape(t.shape))
    return start + n


def load_conv(buf, start, conv_model):
    start = _take(buf, start, conv_model.bias)
    return _take(buf, start, conv_model.weight)


def load_conv_bn(buf, start, conv_model, bn_model):
    for t in (bn_model.bias, bn_model.weight, bn_model.running_mean, bn_model.running_var, conv_model.weight):
        start = _take(buf, start, t)
    return start


def load_fc(buf, start, fc_model):
    start = _take(buf, start, fc_model.bias)
    return _take(buf, start, fc_model.weight)


def _dump(fp, t):
    t.detach().cpu().numpy().tofile(fp)


def save_conv(fp, conv_model):
    _dump(fp, conv_model.bias); _dump(fp, conv_model.weight)


def save_conv_bn(fp, conv_model, bn_model):
    for t in (bn_model.bias, bn_model.weight, bn_model.running_mean, bn_model.running_var, conv_model.weight):
        _dump(fp, t)


def save_fc(fp, fc_model):
    _dump(fp, fc_model.bias); _dump(fp, fc_model.weight)
