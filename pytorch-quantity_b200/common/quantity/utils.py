"""Host-side helpers (reference: quantity/common/quantity/utils.py:7-97)."""
import os

import torch
from torch import nn

from .fabu_layer import Identity


def _replace_module(root, dotted, new):
    parent = root
    parts = dotted.split(".")
    for p in parts[:-1]:
        parent = getattr(parent, p)
    parent.add_module(parts[-1], new)


def merge_bn(model, device="cpu"):
    """Fold every BatchNorm2d into the Conv2d registered just before it and leave an
    ``Identity`` in its place (utils.py:7-65).

        s = gamma / sqrt(var + 1e-5);  W' = s * W;  b' = s * (b - mean) + beta

    evaluated as separate fp32 torch ops in that order (no fused multiply-add), so the folded
    parameters are bit-identical to the reference's on the same device type.  ``device`` is
    kept for signature compatibility ('cuda' moves the scale like utils.py:38-40)."""
    pending = None
    for name, layer in list(model.named_modules()):
        kind = type(layer).__name__
        if kind == "Conv2d":
            pending = layer
        elif kind == "BatchNorm2d":
            assert pending is not None, "Please put bn right after the conv in your __init__()."
            conv = pending
            assert conv.weight is not None, "The conv weight can`t be None"
            w = conv.weight.data
            b = conv.bias.data if conv.bias is not None else torch.zeros(conv.out_channels, device=w.device)
            scale = layer.weight.data / torch.sqrt(layer.running_var + 1e-5)
            if device == "cuda":
                scale, b = scale.cuda(), b.cuda()
            new_w = scale.view(-1, 1, 1, 1) * w
            new_b = scale * (b - layer.running_mean) + layer.bias.data
            conv.weight = nn.Parameter(new_w)
            conv.bias = nn.Parameter(new_b)
            _replace_module(model, name, Identity())
            print("The layer change: {} ==>Identity".format(name))
            pending = None
    return model


def walk_dirs(dir_name, file_type=None):
    """All file paths under ``dir_name`` (optionally filtered by suffix), utils.py:67-77."""
    found = []
    for root, _dirs, files in os.walk(dir_name):
        found.extend(root + "/" + f for f in files if not file_type or f.endswith(file_type))
    return found


def tid(tensor):
    """Value-derived tensor id (utils.py:80-97).  Kept for API compatibility; this
    repository's tracer links producers to consumers by tensor identity instead, which cannot
    collide (tools/pytorch_quantizer.py)."""
    x = tensor.detach().float().cpu()
    first = x[..., 0]

    def h(v):
        return str(int(v.item() * 1e4 % 1e4))

    return h(first.max() + first.min()) + h(x.max() + x.min()) + h(first.mean()) + h(x.mean())
