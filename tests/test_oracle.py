"""The CPU oracle (oracle/) against the golden vectors produced by the unmodified
reference (tests/golden/gen_golden.py).  No GPU, no product code."""
import json

import numpy as np
import pytest

import det_inputs
from conftest import golden_json, load_golden
from golden import gen_golden as gg


def _meta(npz):
    return json.loads(bytes(npz["meta"]).decode())


# ------------------------------------------------------------------ a1 / a2 / a3
@pytest.mark.parametrize("case", list(det_inputs.stats_cases()))
def test_stats_case(oracle, case):
    g = load_golden("stats.npz")
    batches = det_inputs.stats_cases()[case]
    m = 0
    for b in batches:
        m = oracle.absmax_update(m, b)
    assert float(m) == float(g[case + "/max"][0])
    assert type(m).__name__ == _meta(g)[case]["max_type"]
    iv = oracle.interval(m)
    assert float(iv) == float(g[case + "/interval"][0])
    assert type(iv).__name__ == _meta(g)[case]["interval_type"]
    h = np.zeros(2048, dtype=np.int32)
    h_np = np.zeros(2048, dtype=np.int32)
    for b in batches:
        h += oracle.hist(b, iv)
        h_np += oracle.hist_np(b, iv)
    assert np.array_equal(h, g[case + "/hist"])
    assert np.array_equal(h_np, g[case + "/hist"])
    assert h.sum() == sum(int(np.count_nonzero(b)) for b in batches)
    # C absmax agrees with the numpy statement
    mc = np.float32(0)
    for b in batches:
        mc = oracle.absmax_c(b, mc)
    assert float(mc) == float(m)


def test_stats_pool_fanout_equivalent(oracle):
    g = load_golden("stats.npz")
    for case, batches in det_inputs.stats_cases().items():
        iv = g["pool3/" + case + "/interval"][0]
        assert np.array_equal(oracle.hist(batches[0], iv), g["pool3/" + case + "/hist"])


# ---------------------------------------------------------------------- a5 - a7
def _kl_names():
    g = load_golden("kl.npz")
    return sorted(k[:-3] for k in g.files if k.endswith("/kl"))


@pytest.mark.parametrize("name", _kl_names())
def test_kl_search(oracle, name):
    g = load_golden("kl.npz")
    counts = g[name + "/counts"]
    P = oracle.normalize(counts)
    assert np.array_equal(P, oracle.normalize_np(counts))      # C == numpy statement, bitwise
    assert P.dtype == np.float64
    t, kl = oracle.kl_search(P)
    ref_kl = g[name + "/kl"]
    assert kl.shape == ref_kl.shape == (1920,)
    # tolerance stated by north_star: KL within 1e-5 relative (measured: ~1e-15, libm vs numpy log)
    np.testing.assert_allclose(kl, ref_kl, rtol=1e-9, atol=1e-300)
    # the chosen threshold is the first strict minimum of the reference's own curve
    best, t_ref = 66666, 2047
    for i, v in enumerate(ref_kl):
        if v < best:
            best, t_ref = v, 128 + i
    assert t == t_ref
    iv = g[name + "/interval"][0]
    iv = np.float32(iv) if _meta(g)[name]["interval_type"] == "float32" else float(iv)
    bit, thr = oracle.threshold_to_bit(t, iv)
    assert bit == int(g[name + "/bit"][0])
    assert float(thr) == float(g[name + "/threshold_value"][0])
    assert type(thr).__name__ == _meta(g)[name]["threshold_type"]


def test_pairwise_sum_is_numpy_order(oracle):
    rng = np.random.default_rng(0)
    for n in list(range(0, 40)) + [127, 128, 129, 255, 256, 1000, 1919, 1920, 2047]:
        a = rng.random(n) * 10.0 ** rng.integers(-12, 3, size=n)
        assert oracle.pairwise_sum(a) == np.sum(a)


# --------------------------------------------------------------------- a10 / a12
@pytest.mark.parametrize("bit", gg.FQ_BITS)
def test_fakequant(oracle, bit):
    g = load_golden("fakequant.npz")
    x = gg.fakequant_inputs()
    y = oracle.fakequant(x, bit)
    ref = g["y_bit%d" % bit]
    assert np.array_equal(y, ref)
    assert np.array_equal(np.signbit(y), np.signbit(ref))      # -0.0 preserved like torch.round
    assert np.array_equal(oracle.fakequant_c(x, bit), ref)
    assert np.array_equal(oracle.quantize_input(x, bit), g["q_bit%d" % bit])


# ---------------------------------------------------------------- a13 / a14 / a15
@pytest.mark.parametrize("case", gg.INTSIM_CONV_CASES, ids=[c[0] for c in gg.INTSIM_CONV_CASES])
def test_int_conv(oracle, case):
    g = load_golden("intsim.npz")
    name, stride, pad = case[0], case[7], case[8]
    x, w, b, info = gg.intsim_conv_tensors(case)
    y, yq = oracle.int_conv_layer(x, w, b, info, stride=stride, padding=pad)
    assert np.array_equal(y, g["conv/" + name + "/y"])
    assert float(g["conv/" + name + "/acc_absmax"][0]) < 2 ** 24   # fp32 conv was exact in the reference


@pytest.mark.parametrize("case", gg.INTSIM_LINEAR_CASES, ids=[c[0] for c in gg.INTSIM_LINEAR_CASES])
def test_int_linear(oracle, case):
    g = load_golden("intsim.npz")
    x, w, b, info = gg.intsim_linear_tensors(case)
    y, _ = oracle.int_linear_layer(x, w, b, info)
    assert np.array_equal(y, g["linear/" + case[0] + "/y"])


def test_right_shift_and_add(oracle):
    g = load_golden("intsim.npz")
    accs = np.arange(-1100, 1100, dtype=np.int64)
    for rs in (-2, 0, 1, 3, 7):
        assert np.array_equal(oracle.right_shift(accs, rs), g["rshift/rs%d" % rs].astype(np.int64))
    a, c = det_inputs.bell(4096, 41, 60.0), det_inputs.bell(4096, 42, 60.0)
    assert np.array_equal(oracle.add_clamp(a, c), g["add/y"])


# ------------------------------------------------------------ a4 / a8 / a9 (tiny)
def test_tiny_calibration_tables(oracle):
    g = load_golden("tiny_e2e.npz")
    j = golden_json(g)
    net_info = j["net_info"]
    top = ["image"] + list(net_info)
    batches = []
    i = 0
    while "feat%d/image" % i in g.files:
        batches.append({n: g["feat%d/%s" % (i, n)] for n in top})
        i += 1
    assert i == 3
    r = oracle.calibrate(batches, top, net_info, j["merge_groups"], return_all=True)
    assert r["raw_bits"] == j["raw_bits"]
    for n in top:
        assert float(r["intervals"][n]) == j["intervals"][n]
        assert float(r["thresholds"][n]) == j["thresholds"][n]
        assert np.array_equal(np.asarray(r["dists"][n], dtype=np.float64), g["dist/" + n].astype(np.float64))
    lines = oracle.feat_table_lines(top, j["cared_op_layer_names"], net_info, r["bits"])
    assert "\n".join(lines) + "\n" == j["after_weight_quantize"]["feat.table"]


@pytest.mark.parametrize("dkl", [False, True], ids=["maxabs", "dkl"])
def test_tiny_weight_quantize_and_rewrite(oracle, dkl):
    """dkl: the reference's KL mode for weights (``_DKL_weight``, pytorch_quantizer.py:644-648; golden tiny_dkl.npz)."""
    import torch
    g = load_golden("tiny_e2e.npz")
    j = golden_json(load_golden("tiny_dkl.npz")) if dkl else golden_json(g)
    # merged-BN parameters, in named_parameters order of the merged model
    import tiny_fabu_net as tn
    net = tn.build_tiny(0)
    sd = {k[len("state/"):]: g[k] for k in g.files if k.startswith("state/")}
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    params = {}
    mods = dict(net.named_modules())
    prev_conv = None
    for name, m in net.named_modules():
        if type(m).__name__ == "Conv2d":
            prev_conv = (name, m)
        elif type(m).__name__ == "BatchNorm2d":
            cname, conv = prev_conv
            w, b = oracle.merge_bn_params(conv.weight.data, None if conv.bias is None else conv.bias.data,
                                          m.weight.data, m.bias.data, m.running_mean, m.running_var)
            params[cname + ".weight"], params[cname + ".bias"] = w, b
    params["fc.weight"], params["fc.bias"] = mods["fc"].weight.data.numpy(), mods["fc"].bias.data.numpy()
    bits, q = oracle.weight_quantize(params, dkl=dkl)
    snap = j["after_weight_quantize"]
    for name in params:
        sub = "weight/" if name.endswith("weight") else "bias/"
        assert q[name].reshape(-1).tolist() == snap[sub + name + ".json"]["values"], name
    # weight.table after the internal rewrite: bias bit := feat bit, weight bit capped by MAX_SHIFT
    feat = {l.split()[0]: [int(v) for v in l.split()[1:]] for l in snap["feat.table"].strip().split("\n")}
    wbits = {n[:-7]: b for n, b in bits.items() if n.endswith(".weight")}
    need, new_w = oracle.max_shift_limit({k: v[0] for k, v in feat.items()},
                                         {k: v[1:] for k, v in feat.items()}, wbits)
    lines = []
    for name in params:
        if name.endswith(".bias"):
            lines.append("%s %d" % (name, feat[name[:-5]][0]))
        else:
            lines.append("%s %d" % (name, new_w[name[:-7]]))
    assert "\n".join(lines) + "\n" == snap["weight.table"]
    # new_bias after the first rewrite = wrap-rescaled bias (tools/rewriter.py:53-55)
    for name in params:
        if name.endswith(".bias"):
            nb = oracle.rescale_wrap(q[name], bits[name], feat[name[:-5]][0])
            assert nb.reshape(-1).tolist() == snap["new_bias/" + name + ".json"]["values"], name


def test_tiny_recon_outputs(oracle):
    """ReconModel / ReconTest layer outputs of the reference from the oracle's layer restatements."""
    g = load_golden("tiny_e2e.npz")
    j = golden_json(g)
    info = j["quantity_information"]
    # NewConv2d of conv0.0 on the eval batch: needs the merged weights
    import torch
    import tiny_fabu_net as tn
    net = tn.build_tiny(0)
    x = g["eval_batch"]
    conv, bn = net.conv0[0], net.conv0[1]
    w, b = oracle.merge_bn_params(conv.weight.data, None, bn.weight.data, bn.bias.data,
                                  bn.running_mean, bn.running_var)
    y, _ = oracle.int_conv_layer(x, w, b, info["conv0.0"], stride=1, padding=1)
    assert np.array_equal(y, g["ReconModel/layer/conv0.0"])
    # TestConv: fake-quant weights/bias, fp32 conv, fake-quant output (a11)
    wq = oracle.fakequant(w, info["conv0.0"]["weight_bit"])
    bq = oracle.fakequant(b, info["conv0.0"]["bias_bit"])
    out = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(wq), torch.from_numpy(bq), 1, 1)
    assert np.array_equal(oracle.fakequant(out.numpy(), info["conv0.0"]["output_bit"]),
                          g["ReconTest/layer/conv0.0"])


def _requant_int(v, sh, relu):
    """The integer formula pq_concat_requant_s8 / pq_add_requant implement: round_half_even(v * 2^sh), saturated."""
    v = np.maximum(v.astype(np.int64), 0) if relu else v.astype(np.int64)
    if sh >= 0:
        r = np.clip(v, -128, 127) << sh
    else:
        d = -sh
        r = (v + (1 << (d - 1)) - 1 + ((v >> d) & 1)) >> d
    return np.clip(r, -128, 127)


def test_tiny_concat_composition(oracle):
    """Concat -> ReLU -> NewConv2d of the reference (golden layer outputs) from the oracle's restatements, and
    the integer requantisation formula of the concat kernel against that composition."""
    import tiny_fabu_net as tn
    g = load_golden("tiny_e2e.npz")
    info = golden_json(g)["quantity_information"]
    a, b = g["ReconModel/layer/branch_a.0"], g["ReconModel/layer/branch_b.0"]
    cat = np.maximum(np.concatenate([a, b], axis=1), 0)                    # Concat, relu1
    net = tn.build_tiny(0)
    sd = {k[len("state/"):]: g[k] for k in g.files if k.startswith("state/")}
    conv, bn = "conv_c.0", "conv_c.1"
    w, bias = oracle.merge_bn_params(sd[conv + ".weight"], None, sd[bn + ".weight"], sd[bn + ".bias"],
                                     sd[bn + ".running_mean"], sd[bn + ".running_var"])
    y, _ = oracle.int_conv_layer(cat, w, bias, info[conv], stride=1, padding=1)
    assert np.array_equal(y, g["ReconModel/layer/conv_c.0"])
    # integer payloads of the two branches at their own output bits -> the consumer's int8 operand
    ib = info[conv]["input_bit"]
    want = oracle.concat_quantize([np.maximum(a, 0), np.maximum(b, 0)], ib)
    parts = []
    for name, v in (("branch_a.0", a), ("branch_b.0", b)):
        ob = info[name]["output_bit"]
        payload = v * np.float32(2.0 ** ob)
        assert np.array_equal(payload, np.rint(payload)) and np.abs(payload).max() <= 128
        parts.append(_requant_int(payload.astype(np.int64), ib - ob, relu=True))
    assert np.array_equal(np.concatenate(parts, axis=1).astype(np.float32), want)


def test_requant_formula_matches_quantity_composition(oracle):
    """Every (payload bit, consumer bit) pair: the integer formula equals Quantity(q)(payload / 2^bit)."""
    v8 = np.arange(-128, 128, dtype=np.int64)
    for bit in range(0, 8):
        v16 = np.arange(-128 * 2 ** bit, 127 * 2 ** bit + 1, dtype=np.int64)
        for q in range(-4, 12):
            for v in (v8, v16):
                for relu in (False, True):
                    real = v.astype(np.float32) / np.float32(2.0 ** bit)
                    real = np.maximum(real, 0) if relu else real
                    assert np.array_equal(_requant_int(v, q - bit, relu), oracle.quantize_input(real, q)), (bit, q, relu)


def test_absmax_per_channel_oracle_consistency(oracle):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 5, 4, 6)).astype(np.float32)
    pc = oracle.absmax_per_channel(x, 1)
    assert pc.shape == (5,) and pc.dtype == np.float32
    for c in range(5):
        assert pc[c] == oracle.absmax_update(0, x[:, c])                    # the reference's per-tensor rule per plane
    assert pc.max() == oracle.absmax_update(0, x)
    w = rng.standard_normal((7, 3, 3, 3)).astype(np.float32)
    assert np.array_equal(oracle.absmax_per_channel(w, 0), np.abs(w).reshape(7, -1).max(axis=1))


def test_lenet_calibration_tables(oracle):
    """The reference's LeNet example: oracle calibration on the reference's hooked tensors reproduces its
    intervals, histograms, thresholds, bits and feat.table (no merge groups, no BatchNorm)."""
    g = load_golden("lenet_e2e.npz")
    j = golden_json(g)
    net_info = j["net_info"]
    top = ["image"] + list(net_info)
    batches = []
    while "feat%d/image" % len(batches) in g.files:
        batches.append({n: g["feat%d/%s" % (len(batches), n)] for n in top})
    assert len(batches) == 4 and j["merge_groups"] == []
    r = oracle.calibrate(batches, top, net_info, j["merge_groups"], return_all=True)
    assert r["raw_bits"] == j["raw_bits"]
    for n in top:
        assert float(r["intervals"][n]) == j["intervals"][n]
        assert float(r["thresholds"][n]) == j["thresholds"][n]
        assert np.array_equal(np.asarray(r["dists"][n], dtype=np.float64), g["dist/" + n].astype(np.float64))
    lines = oracle.feat_table_lines(top, j["cared_op_layer_names"], net_info, r["bits"])
    assert "\n".join(lines) + "\n" == j["after_weight_quantize"]["feat.table"]


def test_lenet_integer_layers(oracle):
    """ReconModel of LeNet layer by layer from the oracle's NewConv2d / NewLinear restatements: a 1-channel 3x3
    conv, a 5x5 conv on 6 channels (fed through ReLU + max-pool of the reference's own output) and the three
    stacked Linear layers (each fed with the reference's output of the previous one)."""
    import torch
    import torch.nn.functional as F
    g = load_golden("lenet_e2e.npz")
    info = golden_json(g)["quantity_information"]
    sd = {k[len("state/"):]: g[k] for k in g.files if k.startswith("state/")}
    x = g["eval_batch"]
    y0, _ = oracle.int_conv_layer(x, sd["conv.0.weight"], sd["conv.0.bias"], info["conv.0"], stride=1, padding=1)
    assert np.array_equal(y0, g["ReconModel/layer/conv.0"])
    p0 = F.max_pool2d(torch.relu(torch.from_numpy(g["ReconModel/layer/conv.0"])), 2, 2).numpy()
    y1, _ = oracle.int_conv_layer(p0, sd["conv.3.weight"], sd["conv.3.bias"], info["conv.3"], stride=1, padding=0)
    assert np.array_equal(y1, g["ReconModel/layer/conv.3"])
    p1 = F.max_pool2d(torch.relu(torch.from_numpy(g["ReconModel/layer/conv.3"])), 2, 2).numpy().reshape(x.shape[0], -1)
    prev = p1
    for name in ("fc.0", "fc.1", "fc.2"):
        out = oracle.int_linear_layer(prev, sd[name + ".weight"], sd[name + ".bias"], info[name])
        out = out[0] if isinstance(out, tuple) else out
        assert np.array_equal(out, g["ReconModel/layer/" + name]), name
        prev = g["ReconModel/layer/" + name]
    assert np.array_equal(prev, g["ReconModel/y"])


def test_folded_bias_identity(oracle):
    """The identities behind PQ_FLAG_BIAS_FOLDED (include/pq_sm100.h): BiasAdd folds into RightShift's rounding
    add, RightShift's saturation moves to the accumulator with channel-independent bounds, and under a fused
    ReLU the lower bound is redundant -- checked against the oracle's RightShift / BiasAdd / Sp composition."""
    rng = np.random.default_rng(7)
    for rs in (1, 2, 5, 9, 13, 20):
        half = 1 << (rs - 1)
        a_hi, a_lo = 127 * (1 << rs) + half - 1, -128 * (1 << rs) - half + 1
        edges = np.concatenate([np.arange(-6, 7) + e for e in (a_lo, a_hi, 0, -half, half, a_lo - (1 << rs), a_hi + (1 << rs))])
        acc = np.concatenate([edges, rng.integers(-2 ** 29, 2 ** 29, size=4000),
                              rng.integers(-300 << rs, 300 << rs, size=4000),
                              np.arange(-(1 << min(rs + 2, 12)), (1 << min(rs + 2, 12)) + 1)]).astype(np.int64)[:, None]
        b = np.arange(-128, 128, dtype=np.int64)[None, :]
        c = half + (b << rs)
        classic = np.clip(oracle.right_shift(acc, rs) + b, -128, 127)       # RightShift (saturating), BiasAdd, Sp

        def rha_plus_b(a):                                                  # (a + c + (a >> 31)) >> rs
            return (a + c - (a < 0)) >> rs

        folded = np.clip(rha_plus_b(np.clip(acc, a_lo, a_hi)), -128, 127)   # final clip == the saturating pack
        assert np.array_equal(folded, classic), rs
        relu = np.clip(np.maximum(rha_plus_b(np.minimum(acc, a_hi)), 0), -128, 127)
        assert np.array_equal(relu, np.maximum(classic, 0)), rs
        assert np.abs(np.clip(acc, a_lo, a_hi) + c).max() < 2 ** 31
        # round 2: both saturations BEHIND the shift as per-channel bounds on t = r + b (r NOT saturated), after a
        # saturating pack to int16 (monotone; |l|, |h| <= 255): Sp(Sp(r) + b) == clamp(sat16(t), l, h)
        t16 = np.clip(rha_plus_b(acc), -32768, 32767)
        h, l = 127 + np.minimum(b, 0), -128 + np.maximum(b, 0)
        assert np.array_equal(np.clip(t16, l, h), classic), rs
        assert np.array_equal(np.maximum(np.minimum(t16, h), 0), np.maximum(classic, 0)), rs      # VIMNMX.S16x2.RELU
        assert np.abs(acc + c).max() < 2 ** 31


def test_c1_resnet18_reference_histograms(oracle):
    """BASELINE config 1 (ResNet-18 224x224, 64 images, the reference's full CPU run): from the reference's own
    merged histograms and intervals, the oracle's KL search + bit derivation reproduce every threshold and raw bit."""
    g = load_golden("r18_224_c1.npz")
    j = golden_json(g)
    names = [k[len("dist/"):] for k in g.files if k.startswith("dist/")]
    assert len(names) == 30 and "image" in names
    assert int(g["dist/image"].sum()) == 64 * 3 * 224 * 224            # every (non-zero) input element counted once
    for n in names:
        interv = np.float32(j["intervals"][n])
        bit, thr, _t = oracle.quantize_distribution(g["dist/" + n], interv)
        assert float(thr) == j["thresholds"][n], n
        assert bit == j["raw_bits"][n], n


# ------------------------------------------------------------------ INTERVAL_NUM != 2048 (tools/configs.yml:23)
@pytest.mark.parametrize("nbins", [512, 1000, 4096])
def test_other_interval_num_vs_golden(oracle, nbins):
    """The oracle's collector / KL restatement at the reference's other bin counts (bins.npz: the unmodified
    reference with interval_num = 512, 1000, 4096)."""
    from golden import gen_golden as gg
    g = load_golden("bins.npz")
    name = "t%d" % nbins
    batches = gg.bins_batches()
    m = 0
    for b in batches:
        m = oracle.absmax_update(m, b)
    assert float(m) == float(g[name + "/max"][0])
    iv = oracle.interval(m, 1, nbins)
    assert float(iv) == float(g[name + "/interval"][0]) and type(iv).__name__ == "float32"
    h = sum(oracle.hist(b, iv, nbins) for b in batches)
    assert np.array_equal(h, g[name + "/hist"])
    t, kl = oracle.kl_search(oracle.normalize(h))
    assert kl.shape == g[name + "/kl"].shape == (nbins - 128,)
    np.testing.assert_allclose(kl, g[name + "/kl"], rtol=1e-12, atol=0)
    bit, thr = oracle.threshold_to_bit(t, iv)
    assert bit == int(g[name + "/bit"][0]) and float(thr) == float(g[name + "/threshold_value"][0])


# ------------------------------------------- per-channel max-abs, pinned by composition (extension of a1)
def test_per_channel_absmax_vs_reference_on_slices(oracle):
    """channel_max.npz: the UNMODIFIED reference collector run on every channel slice as a tensor of its own (two
    batches, running max) -- the per-channel max-abs by definition -- against the oracle's per-channel statement."""
    from golden import gen_golden as gg
    g = load_golden("channel_max.npz")
    for case in gg.CHANNEL_CASES:
        name, shape, dim, _ = case
        cur = None
        for x in gg.channel_batches(case):
            cur = oracle.absmax_per_channel(x, dim, cur=cur)
        assert cur.dtype == np.float32 and np.array_equal(cur, g[name + "/max"]), name
