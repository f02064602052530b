"""The reference's second example model (quantity/model/lenet/lenet.py, quantity/test/lenet_quantity.py,
lenet_reconstruction.py) end to end on the GPU against the golden run of the unmodified reference
(tests/golden/lenet_e2e.npz): single input channel, 3x3 / 5x5 kernels, three stacked Linear layers whose
widths (120, 84, 10) are not multiples of 16, no BatchNorm."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import golden_json, load_golden
from test_gpu_e2e import _configs, _snapshot

pytestmark = pytest.mark.gpu


def _lenet(g):
    from model.lenet import Cnn
    net = Cnn(1, 10)
    net.load_state_dict({k[len("state/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")})
    return net.eval()


def _write_tables(cfg, j):
    os.makedirs(cfg["OUTPUT"]["WORK_DIR"], exist_ok=True)
    open(cfg["OUTPUT"]["FEAT_BIT_TABLE"], "w").write(j["after_second_rewrite"]["feat.table"])
    open(cfg["OUTPUT"]["WEIGHT_BIT_TABLE"], "w").write(j["after_second_rewrite"]["weight.table"])


def test_lenet_calibration_and_tables(tmp_path):
    import common.quantity as cq
    import tools
    g = load_golden("lenet_e2e.npz")
    j = golden_json(g)
    cfg, user = _configs(tmp_path, (1, 1, 28, 28), 3)
    with torch.no_grad():
        q = tools.Quantity(cq.merge_bn(_lenet(g), "cpu"), config=cfg, user_config=user, verbose=False)
        assert dict(q.net_info) == j["net_info"] and list(q.net_info) == list(j["net_info"])
        assert q.cared_op_layer_names == j["cared_op_layer_names"]
        assert q.get_merge_groups(q.net_info) == j["merge_groups"]
        q.activation_quantize([(torch.from_numpy(g["batch%d" % i]), None) for i in range(4)])
        assert open(cfg["OUTPUT"]["FEAT_BIT_TABLE"]).read() == j["after_weight_quantize"]["feat.table"]
        q.weight_quantize()
        snap1 = _snapshot(cfg)
        q.rewrite_weight()
        snap2 = _snapshot(cfg)
    for snap, key in ((snap1, "after_weight_quantize"), (snap2, "after_second_rewrite")):
        ref = j[key]
        assert snap["feat.table"] == ref["feat.table"] and snap["weight.table"] == ref["weight.table"]
        for name, val in ref.items():
            if isinstance(val, dict):
                assert hashlib.md5(snap[name]).hexdigest() == val["md5"], (key, name)


def test_lenet_statistics_boundary_bit_exact():
    """The reference run's hooked tensors in -> identical maxima, intervals, histograms, thresholds, bits out."""
    import common.quantity as cq
    g = load_golden("lenet_e2e.npz")
    j = golden_json(g)
    assert j["merge_groups"] == []
    top = ["image"] + list(j["net_info"])
    col = cq.DistributionCollector(top)
    feats = []
    while "feat%d/image" % len(feats) in g.files:
        feats.append({n: torch.from_numpy(g["feat%d/%s" % (len(feats), n)]).cuda() for n in top})
    assert len(feats) == 4
    for f in feats:
        col.refresh_max_val(f)
    intervals = col.distribution_intervals
    for n in top:
        assert float(intervals[n]) == j["intervals"][n], n
    for f in feats:
        col.add_to_distributions(f)
    dists = col.distributions
    for n in top:
        assert np.array_equal(dists[n], g["dist/" + n]), n
    qz = cq.Quantizer(top)
    qz.quantize(dists, intervals)
    assert qz.bits == j["raw_bits"]
    for n in top:
        assert float(qz.threshold_value[n]) == j["thresholds"][n], n


@pytest.mark.parametrize("pipeline", [False, True])
def test_lenet_reconmodel_bit_exact(tmp_path, pipeline):
    """Every NewConv2d / NewLinear output of the integer simulation equals the reference's, with the fp32 module
    boundaries and with the int8 pipeline (whose odd channel counts exercise the de-quantise fall-backs)."""
    import tools
    from common.quantity import QTensor, enable_int8_pipeline
    g = load_golden("lenet_e2e.npz")
    j = golden_json(g)
    cfg, user = _configs(tmp_path, (1, 1, 28, 28), 3)
    _write_tables(cfg, j)
    with torch.no_grad():
        r = tools.Reconstruction(_lenet(g), config=cfg)
        r.merge_bn()
        info = r.get_quantity_information()
        for name, d in j["quantity_information"].items():
            for key in ("weight_bit", "bias_bit", "output_bit", "input_bit"):
                assert info[name][key] == d[key], (name, key)
        model = r.ReconModel(info, str(tmp_path / "workdir" / "ReconModel.pth")).cuda()
        if pipeline:
            enable_int8_pipeline(model)
        outs = {}
        for name, mod in model.named_modules():
            if type(mod).__name__ in ("NewConv2d", "NewLinear"):
                mod.register_forward_hook(lambda m, i, o, name=name: outs.__setitem__(name, o))
        y = model(torch.from_numpy(g["eval_batch"]).cuda())
        assert not isinstance(y, QTensor)
        for name, o in outs.items():
            want = g["ReconModel/layer/" + name]
            got = (o.dequantize() if isinstance(o, QTensor) else o).cpu().numpy()
            if isinstance(o, QTensor) and o.nonneg:
                want = np.maximum(want, 0)            # the following ReLU was fused into this epilogue
            assert np.array_equal(got, want), name
    assert np.array_equal(y.cpu().numpy(), g["ReconModel/y"])


def test_lenet_recontest_outputs(tmp_path):
    """ReconTest rebuilt from the golden tables: every output sits exactly on its layer's int8 grid.  Output values
    are compared with zero tolerance against the reference running on this same GPU in test_gpu_vs_reference.py
    (the golden here is an MKLDNN forward, which no cuDNN forward reproduces bit for bit)."""
    import tools
    g = load_golden("lenet_e2e.npz")
    j = golden_json(g)
    cfg, user = _configs(tmp_path, (1, 1, 28, 28), 3)
    _write_tables(cfg, j)
    with torch.no_grad():
        r = tools.Reconstruction(_lenet(g), config=cfg)
        r.merge_bn()
        info = r.get_quantity_information()
        model = r.ReconTest(info, str(tmp_path / "workdir" / "ReconTest.pth")).cuda()
        outs = {}
        for name, mod in model.named_modules():
            if type(mod).__name__ in ("TestConv", "TestLinear"):
                mod.register_forward_hook(lambda m, i, o, name=name: outs.__setitem__(name, o.cpu().numpy()))
        y = model(torch.from_numpy(g["eval_batch"]).cuda()).cpu().numpy()
    assert y.shape == g["ReconTest/y"].shape and len(outs) > 0
    for name, o in outs.items():                       # every output sits on its layer's int8 grid
        s = o * 2.0 ** info[name]["output_bit"]
        assert np.array_equal(s, np.rint(s)) and s.max() <= 127 and s.min() >= -128, name


# ------------------------------------------------------------------ example drivers (quantity/test/*.py counterparts)
EXAMPLES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pytorch-quantity_b200", "test")


@pytest.mark.parametrize("model", ["lenet", "resnet18_cifar"])
def test_example_drivers_run(tmp_path, model):
    """quantity_example.py then reconstruction_example.py, as a user would run them (separate processes, tables
    and JSON passed through the work directory)."""
    import subprocess
    import sys
    wd = str(tmp_path / "workdir")
    r = subprocess.run([sys.executable, "quantity_example.py", "--model", model, "--batches", "3", "--batch", "4",
                        "--workdir", wd], cwd=EXAMPLES, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    table = open(os.path.join(wd, "feat.table")).read().strip().split("\n")
    assert table[0].startswith("image ") and len(table) > 5
    assert os.path.isdir(os.path.join(wd, "new_weight")) and os.listdir(os.path.join(wd, "new_bias"))
    r = subprocess.run([sys.executable, "reconstruction_example.py", "--model", model, "--batches", "2", "--batch", "8",
                        "--workdir", wd, "--int8-pipeline"], cwd=EXAMPLES, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    for line in ("origin model", "merge bn model", "reconstruction model", "q-dq reconstruction model"):
        assert line in r.stdout, r.stdout[-2000:]
    assert os.path.exists(os.path.join(wd, "quantity_model.pth"))


def test_cifar_resnet18_pipeline_equals_fp32_boundary(tmp_path):
    """The reference's own model (CIFAR ResNet-18, 32x32): 4x4 feature maps in the last stage, M smaller than one
    128-row tile per image; int8 pipeline == fp32-boundary ReconModel bit for bit."""
    import sys
    sys.path.insert(0, EXAMPLES)
    import _models
    import tools
    from common.quantity import enable_int8_pipeline, merge_bn
    net, shape = _models.build("resnet18_cifar")
    data = _models.batches(shape, 3, 8)
    cfg, user = _models.configs(str(tmp_path / "wd"), shape, len(data))
    with torch.no_grad():
        q = tools.Quantity(merge_bn(net, "cpu"), config=cfg, user_config=user, verbose=False)
        q.activation_quantize(data)
        q.weight_quantize()
        net2, _ = _models.build("resnet18_cifar")
        r = tools.Reconstruction(net2, config=cfg)
        r.merge_bn()
        model = r.ReconModel(r.get_quantity_information(), None).cuda().eval()
        x = _models.batches(shape, 1, 37, seed=500)[0][0].cuda()
        ref = model(x)
        enable_int8_pipeline(model)
        assert torch.equal(model(x), ref)
        assert torch.equal(model(x), ref)              # second forward: payload census in steady state
    assert torch.isfinite(ref).all() and ref.shape == (37, 10)
