"""End-to-end parity of the drop-in drivers on the GPU against the golden run of the
unmodified reference drivers (tests/golden/tiny_e2e.npz, r18_224_c1.npz)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import golden_json, load_golden

pytestmark = pytest.mark.gpu


def _configs(tmp_path, input_shape, max_cali):
    import tools._config as tc
    cfg = tc.load_tool_config(os.path.join(os.path.dirname(tc.__file__), "configs.yml"))
    wd = str(tmp_path / "workdir")
    cfg["OUTPUT"] = {"WORK_DIR": wd, "WEIGHT_BIT_TABLE": wd + "/weight.table",
                     "FEAT_BIT_TABLE": wd + "/feat.table", "WEIGHT_DIR": wd + "/weight",
                     "BIAS_DIR": wd + "/bias", "FINAL_WEIGHT_DIR": wd + "/new_weight",
                     "FINAL_BIAS_DIR": wd + "/new_bias"}
    cfg["SETTINGS"]["MAX_CALI_IMG_NUM"] = max_cali
    user = tc.load_user_config({"PATH": {}, "MODEL": {"INPUT_SHAPE": ",".join(map(str, input_shape))},
                                "PRE_PROCESS": {"IMG": 1}, "SETTINGS": {"DEVICE": "gpu", "GPU": 0}})
    return cfg, user


def _tiny_model(g):
    import tiny_fabu_net as tn
    net = tn.build_tiny(0)
    sd = {k[len("state/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}
    net.load_state_dict(sd)
    return net.eval()


def _snapshot(cfg):
    out = cfg["OUTPUT"]
    snap = {"feat.table": open(out["FEAT_BIT_TABLE"]).read(), "weight.table": open(out["WEIGHT_BIT_TABLE"]).read()}
    for key, sub in (("WEIGHT_DIR", "weight"), ("BIAS_DIR", "bias"), ("FINAL_WEIGHT_DIR", "new_weight"),
                     ("FINAL_BIAS_DIR", "new_bias")):
        for fn in sorted(os.listdir(out[key])):
            snap[sub + "/" + fn] = open(os.path.join(out[key], fn), "rb").read()
    return snap


def test_tiny_calibration_and_tables(tmp_path):
    """Tracer names, merge groups, per-tensor statistics on the golden activations, feat.table,
    weight.table, weight / bias JSON (md5) after weight_quantize and after the script's second
    rewrite_weight (quirk Q7)."""
    import hashlib
    import common.quantity as cq
    import tools
    g = load_golden("tiny_e2e.npz")
    j = golden_json(g)
    cfg, user = _configs(tmp_path, (1, 3, 16, 16), 2)
    with torch.no_grad():
        net = cq.merge_bn(_tiny_model(g), "cpu")
        q = tools.Quantity(net, config=cfg, user_config=user, verbose=False)
        assert {k: v for k, v in q.net_info.items()} == j["net_info"]
        assert list(q.net_info) == list(j["net_info"])
        assert q.cared_op_layer_names == j["cared_op_layer_names"]
        assert q.get_merge_groups(q.net_info) == j["merge_groups"]
        batches = [(torch.from_numpy(g["batch%d" % i]), None) for i in range(3)]
        q.activation_quantize(batches)
        # the GPU forward (cuDNN) is not bitwise the CPU forward (MKLDNN) of the golden run, so
        # bit-exact statistics parity is pinned at the statistics boundary in the next test; the
        # table (coarse fractional bits) must still agree:
        assert open(cfg["OUTPUT"]["FEAT_BIT_TABLE"]).read() == j["after_weight_quantize"]["feat.table"]
        q.weight_quantize()
        snap1 = _snapshot(cfg)
        q.rewrite_weight()
        snap2 = _snapshot(cfg)
    for snap, key in ((snap1, "after_weight_quantize"), (snap2, "after_second_rewrite")):
        ref = j[key]
        assert snap["weight.table"] == ref["weight.table"]
        for name, val in ref.items():
            if isinstance(val, dict):
                assert hashlib.md5(snap[name]).hexdigest() == val["md5"], (key, name)


def test_tiny_statistics_boundary_bit_exact():
    """Identical activation tensors in (the reference run's hooked tensors) -> identical maxima,
    intervals, histograms (incl. merged groups), thresholds and bits out."""
    import common.quantity as cq
    g = load_golden("tiny_e2e.npz")
    j = golden_json(g)
    net_info, groups = j["net_info"], j["merge_groups"]
    top = ["image"] + list(net_info)
    col = cq.DistributionCollector(top, worker_num=2)
    feats = []
    i = 0
    while "feat%d/image" % i in g.files:
        feats.append({n: torch.from_numpy(g["feat%d/%s" % (i, n)]).cuda() for n in top})
        i += 1
    for f in feats:
        col.refresh_max_val(f)
    intervals = col.distribution_intervals

    def has_elt(names):
        return any(net_info[n]["type"] == "Eltwise" for n in names)

    for names in groups:
        if not has_elt(names):
            w = max(intervals[n] for n in names)
            for n in names:
                intervals[n] = w
    for n in top:
        assert float(intervals[n]) == j["intervals"][n], n
    for f in feats:
        col.add_to_distributions(f)
    dists = col.distributions
    for names in groups:
        if not has_elt(names):
            tot = np.zeros(2048)
            for n in names:
                tot += dists[n]
            for n in names:
                dists[n] = tot
    for n in top:
        assert np.array_equal(np.asarray(dists[n], dtype=np.float64), g["dist/" + n].astype(np.float64)), n
    qz = cq.Quantizer(top)
    qz.quantize(dists, intervals)
    assert qz.bits == j["raw_bits"]
    for n in top:
        assert float(qz.threshold_value[n]) == j["thresholds"][n], n


def test_tiny_recontest_outputs(tmp_path):
    """ReconTest (fake-quant) model rebuilt from the golden tables: the per-layer bits equal the reference's and
    every output sits exactly on its layer's int8 grid.  Output VALUES are compared with zero tolerance against the
    reference running on this same GPU in test_gpu_vs_reference.py (the golden here is an MKLDNN forward, which
    no cuDNN forward reproduces bit for bit)."""
    import common.quantity as cq
    import tools
    g = load_golden("tiny_e2e.npz")
    j = golden_json(g)
    cfg, user = _configs(tmp_path, (1, 3, 16, 16), 2)
    os.makedirs(cfg["OUTPUT"]["WORK_DIR"], exist_ok=True)
    open(cfg["OUTPUT"]["FEAT_BIT_TABLE"], "w").write(j["after_second_rewrite"]["feat.table"])
    open(cfg["OUTPUT"]["WEIGHT_BIT_TABLE"], "w").write(j["after_second_rewrite"]["weight.table"])
    with torch.no_grad():
        net = _tiny_model(g)
        r = tools.Reconstruction(net, config=cfg)
        r.merge_bn()
        info = r.get_quantity_information()
        ref_info = j["quantity_information"]
        for name, d in ref_info.items():
            for key in ("weight_bit", "bias_bit", "output_bit", "input_bit"):
                assert info[name][key] == d[key], (name, key)
        model = r.ReconTest(info, str(tmp_path / "workdir" / "ReconTest.pth")).cuda()
        outs = {}
        for name, mod in model.named_modules():
            if type(mod).__name__ in ("TestConv", "TestLinear"):
                mod.register_forward_hook(lambda m, i, o, name=name: outs.__setitem__(name, o.cpu().numpy()))
        y = model(torch.from_numpy(g["eval_batch"]).cuda()).cpu().numpy()
    assert y.shape == g["ReconTest/y"].shape and len(outs) > 0
    for name, o in outs.items():
        s = o * 2.0 ** info[name]["output_bit"]
        assert np.array_equal(s, np.rint(s)) and s.max() <= 127 and s.min() >= -128, name


def test_tiny_dkl_weight_mode(tmp_path):
    """The reference's KL mode for weights (``_DKL_weight``, pytorch_quantizer.py:62, :644-648): weight histograms
    and the KL search on the GPU give the reference's weight.table and weight / bias JSON."""
    import hashlib
    import common.quantity as cq
    import tools
    g = load_golden("tiny_e2e.npz")
    j = golden_json(load_golden("tiny_dkl.npz"))
    cfg, user = _configs(tmp_path, (1, 3, 16, 16), 2)
    with torch.no_grad():
        q = tools.Quantity(cq.merge_bn(_tiny_model(g), "cpu"), config=cfg, user_config=user, verbose=False)
        q.activation_quantize([(torch.from_numpy(g["batch%d" % i]), None) for i in range(3)])
        q._DKL_weight = True
        q.weight_quantize()
        snap1 = _snapshot(cfg)
        q.rewrite_weight()
        snap2 = _snapshot(cfg)
    for snap, key in ((snap1, "after_weight_quantize"), (snap2, "after_second_rewrite")):
        ref = j[key]
        assert snap["weight.table"] == ref["weight.table"], key
        for name, val in ref.items():
            if isinstance(val, dict):
                assert hashlib.md5(snap[name]).hexdigest() == val["md5"], (key, name)


def test_c1_resnet18_kl_search_on_reference_histograms():
    """BASELINE config 1: the reference's merged histograms of its full ResNet-18 224x224 CPU run (30 tensors) ->
    the GPU KL search + host bit derivation give the reference's thresholds and raw bits exactly."""
    import common.quantity as cq
    g = load_golden("r18_224_c1.npz")
    j = golden_json(g)
    names = [k[len("dist/"):] for k in g.files if k.startswith("dist/")]
    qz = cq.Quantizer(names)
    qz.quantize({n: g["dist/" + n] for n in names}, {n: np.float32(j["intervals"][n]) for n in names})
    assert qz.bits == {n: j["raw_bits"][n] for n in names}
    for n in names:
        assert float(qz.threshold_value[n]) == j["thresholds"][n], n


def test_c1_resnet18_calibration_vs_reference_run(tmp_path):
    """BASELINE config 1 on the GPU against the reference's committed CPU run: same seeded ResNet-18, same 8 batches
    of 8 synthetic 224x224 images.  What does not depend on the forward library must be identical: tracer output,
    merge groups, table layout, weight.table and every int8 weight JSON.  The activation bits come from a cuDNN
    forward here and an MKLDNN forward there, so feat.table is compared -- with ZERO tolerance -- against the
    reference running on this same GPU instead: test_gpu_vs_reference.py::test_calibration_byte_identical_...[r18]."""
    import hashlib
    import sys
    import common.quantity as cq
    import tools
    from model.resnet.resnet_fabu import randomize_bn_, resnet18_fabu
    j = golden_json(load_golden("r18_224_c1.npz"))
    torch.manual_seed(0)
    net = resnet18_fabu().eval()
    with torch.no_grad():
        randomize_bn_(net, 0)
    batches = [(torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(1 + i)), None) for i in range(8)]
    cfg, user = _configs(tmp_path, (1, 3, 224, 224), 7)
    with torch.no_grad():
        q = tools.Quantity(cq.merge_bn(net, "cpu"), config=cfg, user_config=user, verbose=False)
        assert dict(q.net_info) == j["net_info"] and q.cared_op_layer_names == j["cared_op_layer_names"]
        assert q.get_merge_groups(q.net_info) == j["merge_groups"]
        q.activation_quantize(batches)
        q.weight_quantize()
    snap = _snapshot(cfg)
    ref = j["after_weight_quantize"]
    got_lines, ref_lines = snap["feat.table"].strip().split("\n"), ref["feat.table"].strip().split("\n")
    assert [l.split()[0] for l in got_lines] == [l.split()[0] for l in ref_lines]
    assert [len(l.split()) for l in got_lines] == [len(l.split()) for l in ref_lines]
    assert snap["weight.table"] == ref["weight.table"]
    for name, val in ref.items():
        if name.startswith("weight/"):                       # int8 weights: independent of the activation tables
            assert hashlib.md5(snap[name]).hexdigest() == val["md5"], name


class _InplaceNet(torch.nn.Module):
    """conv -> ReLU(inplace) -> conv -> ReLU(inplace) -> avgpool -> view -> fc: the in-place ReLUs overwrite the
    hooked conv outputs after their hooks fired (the reference snapshots to numpy at hook time and is immune)."""

    def __init__(self, inplace):
        super().__init__()
        import common.quantity as cq
        nn = torch.nn
        torch.manual_seed(3)
        self.c1 = nn.Conv2d(3, 8, 3, padding=1)
        self.r1 = nn.ReLU(inplace)
        self.c2 = nn.Conv2d(8, 8, 3, padding=1)
        self.r2 = nn.ReLU(inplace)
        self.pool = nn.AvgPool2d(8)
        self.view = cq.View()
        self.fc = nn.Linear(8, 4)

    def forward(self, x):
        return self.fc(self.view(self.pool(self.r2(self.c2(self.r1(self.c1(x)))))))


def test_inplace_relu_records_pre_modification_values(tmp_path):
    """calibration hooks keep device references; tensors a later in-place op overwrites are found by the
    tracing forward and cloned at hook time, so the statistics are those of the out-of-place model (what the reference
    records, pytorch_quantizer.py:509,513)."""
    import tools
    batches = [(torch.randn(4, 3, 8, 8, generator=torch.Generator().manual_seed(10 + i)), None) for i in range(2)]
    results = {}
    for inplace in (False, True):
        cfg, user = _configs(tmp_path / ("ip%d" % inplace), (1, 3, 8, 8), 1)
        with torch.no_grad():
            q = tools.Quantity(_InplaceNet(inplace).eval(), config=cfg, user_config=user, verbose=False)
            assert bool(q._inplace_modified) == inplace
            q.activation_quantize(batches)
        results[inplace] = q.last_calibration
    a, b = results[False], results[True]
    assert a["table_lines"] == b["table_lines"] and a["top_feat_names"] == b["top_feat_names"]
    for n in a["top_feat_names"]:
        assert float(a["max_vals"][n]) == float(b["max_vals"][n]), n
        assert np.array_equal(a["distributions"][n], b["distributions"][n]), n
    # the conv outputs are signed: had the post-ReLU values been recorded, half of the mass would be missing
    assert a["distributions"]["Conv2d_1"].sum() == 2 * 4 * 8 * 8 * 8


def test_calibration_hooks_reset_themselves_between_plain_forwards(tmp_path):
    """regist_hook_outfeature used the way the reference allows (:491-524): the caller runs model(x)
    directly, several times; names must not keep growing and 'image' must be the latest input."""
    import tools
    cfg, user = _configs(tmp_path, (1, 3, 8, 8), 1)
    with torch.no_grad():
        net = _InplaceNet(False).eval()
        q = tools.Quantity(net, config=cfg, user_config=user, verbose=False)
        feats, hooks = q.regist_hook_outfeature(q.model)
        keys = None
        for i in range(3):
            x = torch.randn(2, 3, 8, 8, generator=torch.Generator().manual_seed(50 + i)).cuda()
            q.model(x)
            if keys is None:
                keys = list(feats)
            assert list(feats) == keys and torch.equal(feats["image"], x)
        for h in hooks:
            h.remove()
    assert keys == ["image", "Conv2d_1", "Conv2d_3", "Linear_7"]
