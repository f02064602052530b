"""Seeded models and synthetic inputs shared by BOTH arms of the live-reference checks: the unmodified
reference running in its own process (baseline/ref_runner.py) and this repository's drop-in package running in
the test / bench process.  The model files import ``common.quantity`` by name, so the same definition binds to
whichever package is on sys.path of the process.  Test / bench infrastructure."""
import importlib.util
import os
import sys

import torch

TESTS = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(os.path.dirname(TESTS), "pytorch-quantity_b200")

INPUT_SHAPE = {"tiny": (1, 3, 16, 16), "lenet": (1, 1, 28, 28), "r18": (1, 3, 224, 224), "r50": (1, 3, 224, 224)}


def _load(modname, relpath):
    """Load a model file of this repo under a private module name (the reference tree has its own ``model``
    package; ``torch.save`` of a whole rebuilt model needs the module importable by that name)."""
    if modname in sys.modules:
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, os.path.join(PKG, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def build_model(name):
    """A fresh, seeded, BN-randomised model in eval mode, BEFORE merge_bn."""
    if name in ("r18", "r50"):
        mod = _load("pq_resnet_fabu", "model/resnet/resnet_fabu.py")
        torch.manual_seed(0)
        net = (mod.resnet18_fabu if name == "r18" else mod.resnet50_fabu)().eval()
        with torch.no_grad():
            mod.randomize_bn_(net, 0)
        return net
    if name == "tiny":
        if TESTS not in sys.path:
            sys.path.insert(0, TESTS)
        import tiny_fabu_net as tn
        return tn.build_tiny(0)
    if name == "lenet":
        mod = _load("pq_lenet", "model/lenet/lenet.py")
        torch.manual_seed(5)
        return mod.Cnn(1, 10).eval()
    raise KeyError(name)


def calib_batches(name, n_batches, batch):
    """``(images, None)`` loader items (PRE_PROCESS.IMG: 1), CPU tensors from CPU generators."""
    shape = INPUT_SHAPE[name][1:]
    out = []
    for i in range(n_batches):
        g = torch.Generator().manual_seed(1 + i)
        x = torch.rand(batch, *shape, generator=g) if name == "lenet" else torch.randn(batch, *shape, generator=g)
        out.append((x, None))
    return out


def eval_batch(name, batch, seed=99):
    shape = INPUT_SHAPE[name][1:]
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, *shape, generator=g) if name == "lenet" else torch.randn(batch, *shape, generator=g)


def set_deterministic():
    """Same library flags in both arms: true fp32 (no TF32), no autotuning, deterministic algorithms."""
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
