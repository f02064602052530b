"""Generate the committed golden fixtures by running the UNMODIFIED reference.

    python tests/golden/gen_golden.py [stats kl fakequant intsim tiny r18_224]

Runs only in the build container (needs /root/reference; see ref_loader.py for the
import shims).  Inputs come from tests/det_inputs.py (regenerated bit-identically by
the tests), so the fixtures hold reference OUTPUTS plus the few small inputs that
depend on torch's RNG.  Test infrastructure -- never imported by the product.
"""
import hashlib
import importlib.util
import io
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
REPO = os.path.dirname(TESTS)
PKG = os.path.join(REPO, "pytorch-quantity_b200")
sys.path.insert(0, TESTS)
sys.path.insert(0, HERE)

import det_inputs  # noqa: E402
import ref_loader  # noqa: E402


def _save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")


def _scalar_record(v):
    """Keep value AND the python/numpy type the reference produced (NEP-50 audit)."""
    return np.array([float(v)], dtype=np.float64), type(v).__name__


# ------------------------------------------------------------------------- stats
def gen_stats():
    dc = ref_loader.load_l2("distribution_collector")
    out, meta = {}, {}
    cases = det_inputs.stats_cases()
    for name, batches in cases.items():
        col = dc.DistributionCollector([name], interval_num=2048, statistic=1, worker_num=1)
        for b in batches:
            col.refresh_max_val({name: b})
        mv = col.max_vals[name]
        iv = col.distribution_intervals[name]
        for b in batches:
            col.add_to_distributions({name: b})
        out[name + "/max"], tmax = _scalar_record(mv)
        out[name + "/interval"], tint = _scalar_record(iv)
        out[name + "/hist"] = col.distributions[name].copy()
        meta[name] = {"max_type": tmax, "interval_type": tint,
                      "hist_dtype": str(col.distributions[name].dtype),
                      "n_batches": len(batches), "sizes": [int(b.size) for b in batches]}
    # fan-out over a Pool of 3 workers on the first batch of every case: same counts expected
    names = list(cases)
    col = dc.DistributionCollector(names, worker_num=3)
    first = {n: cases[n][0] for n in names}
    col.refresh_max_val(first)
    col.add_to_distributions(first)
    for n in names:
        out["pool3/" + n + "/hist"] = col.distributions[n].copy()
        out["pool3/" + n + "/interval"], _ = _scalar_record(col.distribution_intervals[n])
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    _save("stats.npz", **out)


# ---------------------------------------------------------------------------- kl
def kl_histograms():
    """name -> (counts array as the reference would hold it, interval)."""
    def np_hist(x, nbins):
        # plain numpy counting, only to obtain realistic count vectors; the counts
        # themselves are stored in the fixture, so this need not match anything
        iv = np.float32(np.abs(x).max()) / np.float32(nbins) + np.float32(1e-12)
        nz = x[x != 0]
        idx = np.minimum((np.abs(nz) / iv).astype(np.int32), nbins - 1)
        return np.bincount(idx, minlength=nbins).astype(np.int32), np.float32(iv)

    h = {}
    bell = np_hist(det_inputs.bell(200000, 21), 2048)
    relu = np_hist(det_inputs.relu_bell(150000, 22), 2048)
    tail = np_hist(det_inputs.heavy_tail(100000, 23), 2048)
    h["bell"] = (bell[0], bell[1])
    h["relu"] = (relu[0], relu[1])
    h["heavy_tail"] = (tail[0], tail[1])
    h["big_counts_gt_2p24"] = ((bell[0].astype(np.int64) * 9973 + 12345).astype(np.int32), bell[1])
    h["group_sum_f64"] = (bell[0].astype(np.float64) + relu[0], np.float32(max(bell[1], relu[1])))
    sp = bell[0].copy()
    sp[np.arange(2048) % 7 != 0] = 0
    h["sparse_every_7th"] = (sp, np.float32(0.003))
    short = np.zeros(2048, dtype=np.int32)
    short[:100] = bell[0][:100] + 1
    h["support_below_100"] = (short, np.float32(1.5e-4))
    last = np.zeros(2048, dtype=np.int32)
    last[2047] = 5000
    h["all_in_last_bin"] = (last, np.float32(1.0 / 2048) + np.float32(1e-12))
    h["empty"] = (np.zeros(2048, dtype=np.int32), 1e-12)
    flat = np.full(2048, 37, dtype=np.int32)
    h["flat"] = (flat, np.float32(0.25))
    ramp = (np.arange(2048, dtype=np.int32)[::-1] // 3)
    h["ramp_down"] = (ramp, np.float32(7.0))
    h["pow2_interval"] = (bell[0], np.float32(2.0 ** -9))
    return h


def gen_kl():
    qz = ref_loader.load_l2("quantizer")
    hists = kl_histograms()
    names = list(hists)
    out, meta = {}, {}
    curves = {}

    class Rec(qz.Quantizer):
        def compute_kl_divergence(self, a, b):
            v = qz.Quantizer.compute_kl_divergence(self, a, b)
            curves.setdefault(self._cur, []).append(float(v))
            return v

        def normalize_distribution(self, d):
            p = qz.Quantizer.normalize_distribution(self, d)
            out[self._cur + "/P_dtype"] = np.frombuffer(str(p.dtype).encode(), dtype=np.uint8)
            return p

    t0 = time.time()
    for n in names:
        counts, iv = hists[n]
        r = Rec([n])
        r._cur = n
        _, bits, thr = r.quantize_worker([n], {n: counts}, {n: iv})
        out[n + "/counts"] = counts
        out[n + "/interval"], tint = _scalar_record(iv)
        out[n + "/kl"] = np.array(curves[n], dtype=np.float64)
        out[n + "/bit"] = np.array([bits[0]], dtype=np.int64)
        out[n + "/threshold_value"], tthr = _scalar_record(thr[0])
        meta[n] = {"interval_type": tint, "threshold_type": tthr,
                   "counts_dtype": str(counts.dtype)}
        print("kl", n, "bit", bits[0], "thr", thr[0], "%.1fs" % (time.time() - t0))
    # the public entry (multiprocessing.Pool fan-out) must agree
    q = qz.Quantizer(names, worker_num=4)
    q.quantize({n: hists[n][0] for n in names}, {n: hists[n][1] for n in names})
    for n in names:
        assert q.bits[n] == int(out[n + "/bit"][0]), n
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    _save("kl.npz", **out)


# --------------------------------------------------------------------- fakequant
FQ_BITS = (-2, 0, 4, 7, 12)


def fakequant_inputs():
    x = np.concatenate([
        det_inputs.int_grid(2048, 31, -300, 300, 16.0),          # exact .5 ties at bit=4.. and beyond clamp
        det_inputs.bell(2048, 32, 40.0),
        det_inputs.bell(1024, 33, 1e-3),
        np.array([0.0, -0.0, 0.5, -0.5, 1.5, 2.5, -1.5, -2.5, 127.5, -128.5, 1e9, -1e9,
                  1e-30, -1e-30, 0.03125, -0.03125, 127.49999, 126.5, -127.5, 3.0e38],
                 dtype=np.float32)]).astype(np.float32)
    return x


def gen_fakequant():
    import torch
    nq = ref_loader.load_l2("new_quantity_op")
    x = fakequant_inputs()
    out = {}
    for bit in FQ_BITS:
        y = nq.QuanDequan(8, bit)(torch.from_numpy(x.copy())).numpy()
        out["y_bit%d" % bit] = y
        q = nq.Quantity(bit)(torch.from_numpy(x.copy())).numpy()
        out["q_bit%d" % bit] = q
    _save("fakequant.npz", **out)


# ------------------------------------------------------------------------ intsim
INTSIM_CONV_CASES = [
    # name, B, Cin, H, W, Cout, k, stride, pad, bits(w, in, out), wscale, xscale
    ("c3x3_s1", 2, 16, 9, 9, 32, 3, 1, 1, (7, 4, 3), 0.9, 3.0),
    ("c1x1_s1", 3, 64, 7, 7, 128, 1, 1, 0, (8, 5, 4), 0.4, 2.0),
    ("c1x1_s2", 2, 64, 8, 8, 128, 1, 2, 0, (8, 5, 5), 0.4, 2.0),
    ("stem7x7_s2", 2, 3, 32, 32, 64, 7, 2, 3, (8, 5, 4), 0.45, 2.5),
    ("c3x3_s2", 2, 32, 10, 10, 64, 3, 2, 1, (9, 6, 5), 0.2, 1.0),
    ("rs_zero", 1, 8, 5, 5, 16, 3, 1, 1, (2, 1, 3), 20.0, 50.0),
    ("rs_negative", 1, 8, 5, 5, 16, 1, 1, 0, (1, 0, 3), 30.0, 60.0),
    ("saturating", 2, 32, 6, 6, 16, 3, 1, 1, (7, 7, 7), 0.9, 0.9),
    ("odd_shapes", 1, 24, 11, 13, 40, 3, 1, 1, (7, 5, 4), 0.7, 2.0),
    ("bias_none", 1, 16, 6, 6, 16, 3, 1, 1, (7, 4, 4), 0.8, 3.0),
]
INTSIM_LINEAR_CASES = [
    ("fc_small", 4, 64, 24, (8, 5, 3), 0.3, 2.0),
    ("fc_512_10", 3, 512, 10, (9, 6, 2), 0.1, 1.0),
    ("fc_rs_neg", 2, 16, 8, (0, 0, 2), 40.0, 60.0),
]


def intsim_conv_tensors(case):
    name, B, Cin, H, W, Cout, k, stride, pad, bits, wscale, xscale = case
    seed = 1000 + sum(ord(c) for c in name)
    x = det_inputs.bell(B * Cin * H * W, seed, xscale).reshape(B, Cin, H, W)
    w = det_inputs.bell(Cout * Cin * k * k, seed + 1, wscale).reshape(Cout, Cin, k, k)
    b = None if name == "bias_none" else det_inputs.bell(Cout, seed + 2, 4.0)
    info = {"weight_bit": bits[0], "input_bit": bits[1], "output_bit": bits[2], "bias_bit": bits[2]}
    return x, w, b, info


def intsim_linear_tensors(case):
    name, B, fin, fout, bits, wscale, xscale = case
    seed = 2000 + sum(ord(c) for c in name)
    x = det_inputs.bell(B * fin, seed, xscale).reshape(B, fin)
    w = det_inputs.bell(fout * fin, seed + 1, wscale).reshape(fout, fin)
    b = det_inputs.bell(fout, seed + 2, 3.0)
    info = {"weight_bit": bits[0], "input_bit": bits[1], "output_bit": bits[2], "bias_bit": bits[2]}
    return x, w, b, info


def gen_intsim():
    import torch
    import torch.nn as nn
    nq = ref_loader.load_l2("new_quantity_op")
    out = {}
    with torch.no_grad():
        for case in INTSIM_CONV_CASES:
            name, B, Cin, H, W, Cout, k, stride, pad = case[:9]
            x, w, b, info = intsim_conv_tensors(case)
            conv = nn.Conv2d(Cin, Cout, k, stride=stride, padding=pad, bias=b is not None)
            conv.weight.data.copy_(torch.from_numpy(w))
            if b is not None:
                conv.bias.data.copy_(torch.from_numpy(b))
            m = nq.NewConv2d(conv, dict(info))
            y = m(torch.from_numpy(x.copy()))
            out["conv/" + name + "/y"] = y.numpy()
            out["conv/" + name + "/wq"] = m.Conv.weight.data.numpy().astype(np.int8)
            out["conv/" + name + "/bq"] = m.quantized_bias.numpy().astype(np.int32)
            # intermediate after RightShift (before bias) to pin the first saturation
            acc = m.Conv(m.Quan(torch.from_numpy(x.copy())))
            out["conv/" + name + "/acc_absmax"] = np.array([float(acc.abs().max())])
            out["conv/" + name + "/shifted"] = m.RightShift(acc).numpy().astype(np.int32)
        for case in INTSIM_LINEAR_CASES:
            name, B, fin, fout = case[:4]
            x, w, b, info = intsim_linear_tensors(case)
            lin = nn.Linear(fin, fout)
            lin.weight.data.copy_(torch.from_numpy(w))
            lin.bias.data.copy_(torch.from_numpy(b))
            m = nq.NewLinear(lin, dict(info))
            out["linear/" + name + "/y"] = m(torch.from_numpy(x.copy())).numpy()
        # NewAdd and the stand-alone RightShift on crafted accumulators (all .5 ties, both signs)
        a = det_inputs.bell(4096, 41, 60.0)
        c = det_inputs.bell(4096, 42, 60.0)
        out["add/y"] = nq.NewAdd()(torch.from_numpy(a), torch.from_numpy(c)).numpy()
        accs = np.arange(-1100, 1100, dtype=np.float32)
        for rs in (-2, 0, 1, 3, 7):
            out["rshift/rs%d" % rs] = nq.RightShift(8, rs)(torch.from_numpy(accs.copy())).numpy()
    _save("intsim.npz", **out)


# ----------------------------------------------------- intsim: dilation and groups
# name, B, Cin, H, W, Cout, k, stride, pad, dilation, groups, bits(w, in, out), wscale, xscale
INTSIM_EXT_CASES = [
    ("dil2_3x3", 2, 32, 12, 12, 48, 3, 1, 2, 2, 1, (8, 5, 4), 0.4, 2.0),
    ("dil2_3x3_s2", 2, 64, 13, 11, 32, 3, 2, 2, 2, 1, (8, 5, 4), 0.3, 2.0),
    ("dil3_3x3_c16", 1, 16, 15, 15, 16, 3, 1, 3, 3, 1, (7, 4, 3), 0.8, 3.0),
    ("dil2_smallc", 2, 3, 20, 20, 16, 3, 2, 2, 2, 1, (8, 5, 4), 0.45, 2.5),
    ("groups2", 2, 64, 9, 9, 64, 3, 1, 1, 1, 2, (8, 5, 4), 0.4, 2.0),
    ("groups4_1x1", 2, 64, 7, 7, 128, 1, 1, 0, 1, 4, (8, 5, 4), 0.4, 2.0),
    ("groups8_cg4", 1, 32, 10, 10, 32, 3, 2, 1, 1, 8, (7, 5, 4), 0.6, 2.0),
    ("depthwise", 1, 24, 8, 8, 24, 3, 1, 1, 1, 24, (7, 4, 3), 0.9, 3.0),
    ("groups2_dil2", 1, 64, 11, 11, 32, 3, 1, 2, 2, 2, (8, 5, 4), 0.4, 2.0),
]


def intsim_ext_tensors(case):
    name, B, Cin, H, W, Cout, k, stride, pad, dil, groups, bits, wscale, xscale = case
    seed = 3000 + sum(ord(c) for c in name)
    x = det_inputs.bell(B * Cin * H * W, seed, xscale).reshape(B, Cin, H, W)
    w = det_inputs.bell(Cout * (Cin // groups) * k * k, seed + 1, wscale).reshape(Cout, Cin // groups, k, k)
    b = det_inputs.bell(Cout, seed + 2, 4.0)
    info = {"weight_bit": bits[0], "input_bit": bits[1], "output_bit": bits[2], "bias_bit": bits[2]}
    return x, w, b, info


def gen_intsim_ext():
    """NewConv2d around dilated / grouped nn.Conv2d modules (the reference wraps any Conv2d,
    new_quantity_op.py:104-133)."""
    import torch
    import torch.nn as nn
    nq = ref_loader.load_l2("new_quantity_op")
    out = {}
    with torch.no_grad():
        for case in INTSIM_EXT_CASES:
            name, B, Cin, H, W, Cout, k, stride, pad, dil, groups = case[:11]
            x, w, b, info = intsim_ext_tensors(case)
            conv = nn.Conv2d(Cin, Cout, k, stride=stride, padding=pad, dilation=dil, groups=groups)
            conv.weight.data.copy_(torch.from_numpy(w))
            conv.bias.data.copy_(torch.from_numpy(b))
            m = nq.NewConv2d(conv, dict(info))
            out["conv/" + name + "/y"] = m(torch.from_numpy(x.copy())).numpy()
            print(name, out["conv/" + name + "/y"].shape)
    _save("intsim_ext.npz", **out)


# ---------------------------------------------------------------- end-to-end runs
def _read_workdir(test_dir):
    wd = os.path.join(test_dir, "workdir")
    snap = {}
    for fn in ("feat.table", "weight.table"):
        p = os.path.join(wd, fn)
        if os.path.exists(p):
            snap[fn] = open(p).read()
    for sub in ("weight", "bias", "new_weight", "new_bias"):
        d = os.path.join(wd, sub)
        if not os.path.isdir(d):
            continue
        for fn in sorted(os.listdir(d)):
            raw = open(os.path.join(d, fn), "rb").read()
            snap[sub + "/" + fn] = {"md5": hashlib.md5(raw).hexdigest(),
                                    "values": np.array(json.loads(raw)).reshape(-1).tolist()
                                    if len(raw) < 200000 else None}
    return snap


def _run_reference_pipeline(build_model, batches, input_shape, eval_batch, tag, worker_num=4,
                            keep_feats=True, do_recon=True, dkl_weight=False):
    import torch
    with ref_loader.reference_tools(input_shape, max_cali=len(batches) - 1,
                                    worker_num=worker_num,
                                    extra_sys_path=(TESTS, PKG)) as (tools, test_dir):
        from common.quantity import merge_bn
        pq = sys.modules["tools.pytorch_quantizer"]
        rec = {"feats": [], "hist_calls": 0}

        # instrument by patching methods on the reference classes (module-level wrappers keep
        # the instances picklable for the reference's multiprocessing.Pool fan-out)
        orig_c, orig_q = pq.DistributionCollector, pq.Quantizer
        orig_refresh, orig_quantize = orig_c.refresh_max_val, orig_q.quantize

        def rec_refresh(self, tensors):
            if keep_feats:
                rec["feats"].append({k: np.array(v, copy=True) for k, v in tensors.items()})
            return orig_refresh(self, tensors)

        def rec_quantize(self, distributions, distribution_intervals):
            rec["dists"] = {k: np.array(v, copy=True) for k, v in distributions.items()}
            rec["intervals"] = dict(distribution_intervals)
            r = orig_quantize(self, distributions, distribution_intervals)
            rec["raw_bits"] = dict(self.bits)
            rec["thresholds"] = dict(self.threshold_value)
            return r

        orig_c.refresh_max_val, orig_q.quantize = rec_refresh, rec_quantize
        try:
            with torch.no_grad():
                t0 = time.time()
                net = merge_bn(build_model(), "cpu")
                q = tools.Quantity(net)
                q.activation_quantize(batches)
                t_act = time.time() - t0
                t0 = time.time()
                q._DKL_weight = bool(dkl_weight)        # pytorch_quantizer.py:62, :644-648
                q.weight_quantize()
                t_w = time.time() - t0
                snap1 = _read_workdir(test_dir)
                q.rewrite_weight()                      # the script's second call (quirk Q7)
                snap2 = _read_workdir(test_dir)
        finally:
            orig_c.refresh_max_val, orig_q.quantize = orig_refresh, orig_quantize
        res = {
            "net_info": {k: {"inputs": v["inputs"], "type": v["type"]} for k, v in q.net_info.items()},
            "cared_op_layer_names": q.cared_op_layer_names,
            "merge_groups": q.get_merge_groups(q.net_info),
            "raw_bits": rec["raw_bits"],
            "thresholds": {k: float(v) for k, v in rec["thresholds"].items()},
            "intervals": {k: float(v) for k, v in rec["intervals"].items()},
            "after_weight_quantize": {k: v for k, v in snap1.items()},
            "after_second_rewrite": {k: v for k, v in snap2.items()},
            "seconds_activation_quantize": t_act, "seconds_weight_quantize": t_w,
            "worker_num": worker_num, "cpu_count": os.cpu_count(),
        }
        arrays = {}
        for k, v in rec["dists"].items():
            arrays["dist/" + k] = v
        if keep_feats:
            for i, feats in enumerate(rec["feats"]):
                for k, v in feats.items():
                    arrays["feat%d/%s" % (i, k)] = v
        if do_recon:
            with torch.no_grad():
                for mode in ("ReconModel", "ReconTest"):
                    net = build_model()
                    r = tools.Reconstruction(net)
                    r.merge_bn()
                    net.eval()
                    info = r.get_quantity_information()
                    if mode == "ReconModel":
                        res["quantity_information"] = {
                            k: {kk: vv for kk, vv in v.items() if kk != "layer"}
                            for k, v in info.items()}
                    model = getattr(r, mode)(info, os.path.join(test_dir, "workdir", mode + ".pth"))
                    layer_out = {}
                    hooks = []
                    for name, mod in model.named_modules():
                        if type(mod).__name__ in ("NewConv2d", "NewLinear", "NewAdd", "TestConv", "TestLinear"):
                            hooks.append(mod.register_forward_hook(
                                lambda m, i, o, name=name: layer_out.__setitem__(name, o.detach().numpy().copy())))
                    y = model(eval_batch.clone())
                    for h in hooks:
                        h.remove()
                    arrays[mode + "/y"] = y.numpy()
                    if keep_feats:
                        for k, v in layer_out.items():
                            arrays[mode + "/layer/" + k] = v
        return res, arrays


def gen_tiny():
    import torch
    ref_loader.install_shims()
    sys.path.insert(0, ref_loader.REF_QUANTITY)   # tiny_fabu_net imports common.quantity (the reference's here)
    import tiny_fabu_net as tn
    batches = tn.tiny_batches(3, 2, seed=1)
    eval_batch = tn.tiny_batches(1, 4, seed=99)[0][0]
    res, arrays = _run_reference_pipeline(lambda: tn.build_tiny(0), batches, (1, 3, 16, 16),
                                          eval_batch, "tiny", worker_num=2)
    # the tiny model and inputs depend on torch's RNG: commit them (a few thousand floats)
    net = tn.build_tiny(0)
    for k, v in net.state_dict().items():
        arrays["state/" + k] = v.numpy()
    for i, (img, _) in enumerate(batches):
        arrays["batch%d" % i] = img.numpy()
    arrays["eval_batch"] = eval_batch.numpy()
    arrays["json"] = np.frombuffer(json.dumps(res).encode(), dtype=np.uint8)
    _save("tiny_e2e.npz", **arrays)


def gen_tiny_dkl():
    """The tiny net with the reference's KL mode for WEIGHTS switched on (``_DKL_weight``,
    pytorch_quantizer.py:62, :644-648): weight bits from histogram + KL search instead of max-abs."""
    ref_loader.install_shims()
    sys.path.insert(0, ref_loader.REF_QUANTITY)
    import tiny_fabu_net as tn
    batches = tn.tiny_batches(3, 2, seed=1)
    res, _arrays = _run_reference_pipeline(lambda: tn.build_tiny(0), batches, (1, 3, 16, 16), None, "tiny_dkl",
                                           worker_num=2, keep_feats=False, do_recon=False, dkl_weight=True)
    keep = {k: res[k] for k in ("after_weight_quantize", "after_second_rewrite")}
    _save("tiny_dkl.npz", json=np.frombuffer(json.dumps(keep).encode(), dtype=np.uint8))


def gen_lenet():
    """The reference's LeNet example (quantity/test/lenet_quantity.py + lenet_reconstruction.py) on synthetic
    MNIST-shaped inputs: single input channel, 5x5 kernel, three stacked Linear layers, no BatchNorm."""
    import torch
    ref_loader.install_shims()
    sys.path.insert(0, ref_loader.REF_QUANTITY)   # model.lenet imports common.quantity (the reference's here)
    spec = importlib.util.spec_from_file_location("pq_lenet", os.path.join(PKG, "model", "lenet", "lenet.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["pq_lenet"] = mod                 # the reference torch.save()s the whole reconstructed model
    spec.loader.exec_module(mod)

    def build():
        torch.manual_seed(5)
        return mod.Cnn(1, 10).eval()

    batches = mod.lenet_batches(4, 8, seed=1)
    eval_batch = mod.lenet_batches(1, 16, seed=77)[0][0]
    res, arrays = _run_reference_pipeline(build, batches, (1, 1, 28, 28), eval_batch, "lenet", worker_num=2)
    for k, v in build().state_dict().items():
        arrays["state/" + k] = v.numpy()
    for i, (img, _) in enumerate(batches):
        arrays["batch%d" % i] = img.numpy()
    arrays["eval_batch"] = eval_batch.numpy()
    arrays["json"] = np.frombuffer(json.dumps(res).encode(), dtype=np.uint8)
    _save("lenet_e2e.npz", **arrays)


def gen_r18_224():
    """BASELINE.json configs[0] (C1): ResNet-18 224x224, 64 images as 8 batches of 8,
    the reference's CPU path in full.  Only tables / bits / histograms are kept."""
    import torch
    spec = importlib.util.spec_from_file_location(
        "pq_resnet_fabu", os.path.join(PKG, "model", "resnet", "resnet_fabu.py"))

    def build():
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        torch.manual_seed(0)
        net = mod.resnet18_fabu().eval()
        with torch.no_grad():
            mod.randomize_bn_(net, 0)
        return net

    batches = []
    for i in range(8):
        g = torch.Generator().manual_seed(1 + i)
        batches.append((torch.randn(8, 3, 224, 224, generator=g), None))
    wn = min(os.cpu_count(), 30)
    res, arrays = _run_reference_pipeline(build, batches, (1, 3, 224, 224), None, "r18_224",
                                          worker_num=wn, keep_feats=False, do_recon=False)
    for k in ("after_weight_quantize", "after_second_rewrite"):
        for kk, vv in res[k].items():
            if isinstance(vv, dict):
                vv["values"] = None          # 11.7 M weights: keep md5 only
    arrays["json"] = np.frombuffer(json.dumps(res).encode(), dtype=np.uint8)
    _save("r18_224_c1.npz", **arrays)
    print("reference C1 wall: activation_quantize %.1fs, weight_quantize %.1fs on %d workers"
          % (res["seconds_activation_quantize"], res["seconds_weight_quantize"], wn))


# ------------------------------------------------- per-channel max-abs, pinned by composition
# name, shape, channel_dim, scale: activations [N][C][H][W], an FC activation [N][C], a conv weight per output channel
CHANNEL_CASES = [("act_nchw", (3, 16, 9, 7), 1, 2.0), ("act_7x7", (4, 40, 7, 7), 1, 1.0), ("fc", (5, 24), 1, 3.0),
                 ("weight_k", (12, 8, 3, 3), 0, 0.05)]


def channel_batches(case):
    name, shape, dim, scale = case
    seed = 5000 + sum(ord(c) for c in name)
    n = int(np.prod(shape))
    return [det_inputs.bell(n, seed + i, scale * (1 + i)).reshape(shape) for i in range(2)]


def gen_channel():
    """The reference has no per-channel reduction, but its per-tensor one (refresh_max_val,
    distribution_collector.py:70-78) applied to every channel slice AS A TENSOR OF ITS OWN is the per-channel
    max-abs by definition: the unmodified reference collector is run on the slices (two batches, running max)."""
    dc = ref_loader.load_l2("distribution_collector")
    out = {}
    for case in CHANNEL_CASES:
        name, shape, dim, _ = case
        C = shape[dim]
        names = ["c%d" % c for c in range(C)]
        col = dc.DistributionCollector(names)
        for x in channel_batches(case):
            col.refresh_max_val({"c%d" % c: np.ascontiguousarray(np.take(x, c, axis=dim)).reshape(-1) for c in range(C)})
        out[name + "/max"] = np.array([col.max_vals[n] for n in names], dtype=np.float32)
        print("channel", name, out[name + "/max"][:4])
    _save("channel_max.npz", **out)


# ------------------------------------------------------------ INTERVAL_NUM != 2048
BINS_CASES = (512, 1000, 4096)


def bins_batches():
    return [det_inputs.bell(60000, 71, 1.5), det_inputs.relu_bell(50000, 72, 3.0)]


def gen_bins():
    """INTERVAL_NUM is a configuration value (tools/configs.yml:23): the reference's collector and KL search
    at 512, 1000 (not a power of two) and 4096 bins."""
    dc = ref_loader.load_l2("distribution_collector")
    qz = ref_loader.load_l2("quantizer")
    out = {}
    curves = {}

    class Rec(qz.Quantizer):
        def compute_kl_divergence(self, a, b):
            v = qz.Quantizer.compute_kl_divergence(self, a, b)
            curves.setdefault(self._cur, []).append(float(v))
            return v

    for nbins in BINS_CASES:
        name = "t%d" % nbins
        col = dc.DistributionCollector([name], interval_num=nbins, statistic=1, worker_num=1)
        for b in bins_batches():
            col.refresh_max_val({name: b})
        iv = col.distribution_intervals[name]
        for b in bins_batches():
            col.add_to_distributions({name: b})
        hist = col.distributions[name].copy()
        r = Rec([name])
        r._cur = name
        _, bits, thr = r.quantize_worker([name], {name: hist}, {name: iv})
        out[name + "/max"], _ = _scalar_record(col.max_vals[name])
        out[name + "/interval"], _ = _scalar_record(iv)
        out[name + "/hist"] = hist
        out[name + "/kl"] = np.array(curves[name], dtype=np.float64)
        out[name + "/bit"] = np.array([bits[0]], dtype=np.int64)
        out[name + "/threshold_value"], _ = _scalar_record(thr[0])
        print("bins", nbins, "bit", bits[0], "thr", thr[0], len(curves[name]))
    _save("bins.npz", **out)


SECTIONS = {"channel": gen_channel, "bins": gen_bins, "intsim_ext": gen_intsim_ext, "stats": gen_stats, "kl": gen_kl, "fakequant": gen_fakequant, "intsim": gen_intsim,
            "tiny": gen_tiny, "tiny_dkl": gen_tiny_dkl, "lenet": gen_lenet, "r18_224": gen_r18_224}

if __name__ == "__main__":
    assert ref_loader.available(), "needs /root/reference"
    todo = sys.argv[1:] or ["stats", "kl", "fakequant", "intsim", "tiny"]
    for s in todo:
        print("==", s)
        SECTIONS[s]()
