"""GPU parity: the sm_100a kernels, called through the C ABI by the drop-in Python classes,
against (1) the golden vectors of the unmodified reference and (2) the CPU oracle on the same
seeded inputs.  Bit-exact for counts, maxima, thresholds, bits, fake-quant and integer
simulation; KL divergences within 1e-9 relative (north_star allows 1e-5)."""
import json

import numpy as np
import pytest
import torch

import det_inputs
from conftest import golden_json, load_golden
from golden import gen_golden as gg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cq():
    import common.quantity as cq
    from common.quantity import _native
    _native.lib()                       # fail loudly if the CUDA library is missing
    return cq


def _meta(npz):
    return json.loads(bytes(npz["meta"]).decode())


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


# ------------------------------------------------------------------ a1 / a2 / a3
@pytest.mark.parametrize("case", list(det_inputs.stats_cases()))
@pytest.mark.parametrize("as_numpy", [False, True])
def test_stats_case_vs_golden(cq, oracle, case, as_numpy):
    g = load_golden("stats.npz")
    batches = det_inputs.stats_cases()[case]
    col = cq.DistributionCollector([case], worker_num=4)
    for b in batches:
        col.refresh_max_val({case: b if as_numpy else dev(b)})
    m = col.max_vals[case]
    assert float(m) == float(g[case + "/max"][0])
    assert type(m).__name__ == _meta(g)[case]["max_type"]
    iv = col.distribution_intervals[case]
    assert float(iv) == float(g[case + "/interval"][0])
    assert type(iv).__name__ == _meta(g)[case]["interval_type"]
    for b in batches:
        col.add_to_distributions({case: b if as_numpy else dev(b)})
    h = col.distributions[case]
    assert h.dtype == np.int32
    assert np.array_equal(h, g[case + "/hist"])


def test_stats_multi_tensor_one_launch(cq, oracle):
    """All cases in ONE collector (one multi-tensor launch per pass) == per-case results."""
    g = load_golden("stats.npz")
    cases = det_inputs.stats_cases()
    names = list(cases)
    col = cq.DistributionCollector(names)
    first = {n: dev(cases[n][0]) for n in names}
    col.refresh_max_val(first)
    col.add_to_distributions(first)
    for n in names:
        assert float(col.distribution_intervals[n]) == float(g["pool3/" + n + "/interval"][0])
        assert np.array_equal(col.distributions[n], g["pool3/" + n + "/hist"]), n


@pytest.mark.parametrize("offset", [1, 2, 3, 5])
def test_stats_unaligned_views(cq, oracle, offset):
    """Views that start off a 16-byte boundary take the scalar head/tail path."""
    base = det_inputs.bell(100003, 77)
    x = base[offset:offset + 99991]
    t = dev(base)[offset:offset + 99991]
    col = cq.DistributionCollector(["x"])
    col.refresh_max_val({"x": t})
    m = oracle.absmax_update(0, x)
    assert float(col.max_vals["x"]) == float(m)
    col.add_to_distributions({"x": t})
    assert np.array_equal(col.distributions["x"], oracle.hist(x, oracle.interval(m)))


@pytest.mark.parametrize("kind", ["bell", "relu", "const", "tail"])
def test_stats_large_vs_oracle(cq, oracle, kind):
    n = 6_000_011
    x = {"bell": lambda: det_inputs.bell(n, 5),
         "relu": lambda: det_inputs.relu_bell(n, 6),
         "const": lambda: det_inputs.constant(n, -2.5),
         "tail": lambda: det_inputs.heavy_tail(n, 7)}[kind]()
    col = cq.DistributionCollector(["x"])
    col.refresh_max_val({"x": dev(x)})
    m = oracle.absmax_update(0, x)
    assert float(col.max_vals["x"]) == float(m)
    col.add_to_distributions({"x": dev(x)})
    assert np.array_equal(col.distributions["x"], oracle.hist(x, oracle.interval(m)))


def test_hist_division_exactness_at_bin_edges(cq, oracle):
    """The kernel replaces the IEEE division by reciprocal-multiply + FMA correction with an exact
    fallback; stress it where it could go wrong: values within +-3 ulps of EVERY bin edge k*d, for
    divisors that are powers of two, have all-ones / sparse mantissas, are tiny or huge."""
    rng = np.random.default_rng(11)
    ds = [np.float32(v) for v in (2.0 ** -9, 2.0 ** -20, 1.0, 3.0, 1e-12, 1.1754944e-38 * 4096, 1.6e35,
                                  0.0013031913, 7.0, 4.8828125, 1.9999999, 1.0000001, 0.33333334)]
    ds += [np.uint32(0x3a7fffff).view(np.float32), np.uint32(0x3affffff).view(np.float32),
           np.uint32(0x3a800001).view(np.float32)]
    ds += [np.float32(v) for v in np.exp(rng.uniform(np.log(1e-8), np.log(1e3), size=24))]
    names, tensors, expect = [], {}, {}
    for i, d in enumerate(ds):
        k = np.arange(1, 2050, dtype=np.float64)
        base = (k * np.float64(d)).astype(np.float32)
        vals = [base]
        up, dn = base.copy(), base.copy()
        for _ in range(3):
            up = np.nextafter(up, np.float32(np.inf)); dn = np.nextafter(dn, np.float32(0))
            vals += [up.copy(), dn.copy()]
        x = np.concatenate(vals + [-(base), (rng.random(5000) * 2048 * float(d)).astype(np.float32)])
        x = x[np.isfinite(x)]
        n = "d%d" % i
        names.append(n); tensors[n] = x; expect[n] = oracle.hist(x, d)
    col = cq.DistributionCollector(names)
    col.refresh_max_val({n: dev(tensors[n]) for n in names})
    col.distribution_intervals                         # materialise, then override the bin widths
    for n, d in zip(names, ds):
        col._distribution_intervals[n] = d
    col.add_to_distributions({n: dev(tensors[n]) for n in names})
    for n in names:
        assert np.array_equal(col.distributions[n], expect[n]), (n, float(col._distribution_intervals[n]))


def test_stats_full_size_properties(cq):
    """BASELINE-scale tensor (2^28 elements = 1 GiB): size-independent properties.
    sum(hist) == #non-zeros; hist(a) + hist(b) == hist(a ++ b); max == torch's."""
    n = 1 << 28
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, device="cuda", generator=g)
    x[::7] = 0.0
    col = cq.DistributionCollector(["x"])
    col.refresh_max_val({"x": x})
    assert float(col.max_vals["x"]) == float(x.abs().max())
    col.add_to_distributions({"x": x})
    whole = col._hist.clone()
    assert int(whole.sum()) == int((x != 0).sum())
    iv = col.distribution_intervals["x"]
    # reference arithmetic on the device with torch ops as an independent cross-check
    # (a 0-dim CUDA divisor forces a true IEEE division; a python scalar would become x * (1/iv))
    idx = torch.clamp((x[x != 0].abs() / torch.tensor(float(iv), device="cuda")).to(torch.int32), max=2047)
    assert torch.equal(torch.bincount(idx, minlength=2048), whole[0])
    col2 = cq.DistributionCollector(["x"])
    col2.refresh_max_val({"x": x})
    half = n // 2 + 12345
    col2.add_to_distributions({"x": x[:half]})
    col2.add_to_distributions({"x": x[half:]})
    assert torch.equal(col2._hist, whole)


# ---------------------------------------------------------------------- a5 - a7
def test_kl_search_vs_golden(cq):
    g = load_golden("kl.npz")
    names = sorted(k[:-3] for k in g.files if k.endswith("/kl"))
    meta = _meta(g)
    dists, intervals = {}, {}
    for n in names:
        dists[n] = g[n + "/counts"]
        iv = g[n + "/interval"][0]
        intervals[n] = np.float32(iv) if meta[n]["interval_type"] == "float32" else float(iv)
    q = cq.Quantizer(names, worker_num=4, keep_curves=True)
    q.quantize(dists, intervals)
    curves = q.kl_curves.cpu().numpy()
    for i, n in enumerate(names):
        ref = g[n + "/kl"]
        np.testing.assert_allclose(curves[i], ref, rtol=1e-9, atol=1e-300, err_msg=n)
        best, t_ref = 66666, 2047
        for j, v in enumerate(ref):
            if v < best:
                best, t_ref = v, 128 + j
        assert q.threshold_bin[n] == t_ref, n
        assert q.bits[n] == int(g[n + "/bit"][0]), n
        assert float(q.threshold_value[n]) == float(g[n + "/threshold_value"][0]), n


def test_kl_search_vs_oracle_random(cq, oracle):
    rng = np.random.default_rng(5)
    names, dists, intervals = [], {}, {}
    for i in range(24):
        sigma = rng.uniform(60, 900)
        x = np.abs(rng.normal(0, sigma, size=rng.integers(1000, 400000)))
        h = np.bincount(np.minimum(x.astype(np.int64), 2047), minlength=2048).astype(np.int32)
        if i % 5 == 0:
            h[rng.integers(0, 2048, size=600)] = 0
        n = "t%d" % i
        names.append(n)
        dists[n] = h
        intervals[n] = np.float32(rng.uniform(1e-4, 0.1))
    q = cq.Quantizer(names, keep_curves=True)
    q.quantize(dists, intervals)
    curves = q.kl_curves.cpu().numpy()
    for i, n in enumerate(names):
        t, kl = oracle.kl_search(oracle.normalize(dists[n]))
        np.testing.assert_allclose(curves[i], kl, rtol=1e-9, atol=1e-300)
        assert q.threshold_bin[n] == t
        assert q.bits[n] == oracle.threshold_to_bit(t, intervals[n])[0]


# --------------------------------------------------------------------- a10 / a12
@pytest.mark.parametrize("bit", gg.FQ_BITS)
def test_fakequant_vs_golden(cq, bit):
    g = load_golden("fakequant.npz")
    x = gg.fakequant_inputs()
    y = cq.QuanDequan(8, bit)(dev(x)).cpu().numpy()
    ref = g["y_bit%d" % bit]
    assert np.array_equal(y, ref)
    assert np.array_equal(np.signbit(y), np.signbit(ref))
    q = cq.Quantity(bit)(dev(x)).cpu().numpy()
    assert np.array_equal(q, g["q_bit%d" % bit])


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 1 << 20, (1 << 22) + 3])
def test_fakequant_sizes_vs_oracle(cq, oracle, n):
    x = det_inputs.bell(n, 123, 30.0)
    for bit in (-2, 3, 7):
        y = cq.QuanDequan(8, bit)(dev(x)).cpu().numpy()
        assert np.array_equal(y, oracle.fakequant(x, bit))
    t = dev(np.concatenate([np.zeros(1, np.float32), x]))[1:]       # unaligned view
    assert np.array_equal(cq.QuanDequan(8, 5)(t).cpu().numpy(), oracle.fakequant(x, 5))


def test_fakequant_full_size_properties(cq):
    """C2-scale tensor: idempotence, range and grid membership; equality with the 4 ATen ops."""
    n = 636_000_000 // 4
    x = torch.randn(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9)) * 3
    op = cq.QuanDequan(8, 4)
    y = op(x)
    assert torch.equal(op(y), y)
    assert float(y.max()) <= 127 / 16 and float(y.min()) >= -128 / 16
    assert torch.equal(y * 16, torch.round(y * 16))
    ref = torch.div(torch.round(torch.mul(x, 16)).clamp(-128, 127), 16)
    assert torch.equal(y, ref)


# ---------------------------------------------------------------------- a14 / a15
def test_rshift_add_standalone_vs_golden(cq, oracle):
    g = load_golden("intsim.npz")
    accs = np.arange(-1100, 1100, dtype=np.float32)
    for rs in (-2, 0, 1, 3, 7):
        y = cq.RightShift(8, rs)(dev(accs)).cpu().numpy()
        assert np.array_equal(y, g["rshift/rs%d" % rs])
    a, c = det_inputs.bell(4096, 41, 60.0), det_inputs.bell(4096, 42, 60.0)
    assert np.array_equal(cq.NewAdd()(dev(a), dev(c)).cpu().numpy(), g["add/y"])
    x = det_inputs.bell(5000, 43, 100.0)
    assert np.array_equal(cq.Sp(8)(dev(x)).cpu().numpy(), np.clip(x, -128, 127))
    assert np.array_equal(cq.DeQuantity(5)(dev(x)).cpu().numpy(), x / np.float32(32))


def test_quantize_nchw_to_nhwc(cq, oracle):
    from common.quantity import _native
    for (N, C, H, W, cpad) in [(2, 3, 9, 7, 16), (1, 64, 5, 5, 64), (3, 24, 11, 13, 32), (2, 130, 4, 4, 144)]:
        x = det_inputs.bell(N * C * H * W, 50 + C, 4.0).reshape(N, C, H, W)
        q = _native.quantize_nchw_to_nhwc_s8(dev(x), 4, cpad).cpu().numpy()
        ref = np.zeros((N, H, W, cpad), np.int8)
        ref[..., :C] = oracle.quantize_input(x, 4).transpose(0, 2, 3, 1).astype(np.int8)
        assert np.array_equal(q, ref)


# ------------------------------------------------------------------ C-ABI behaviour
def test_abi_error_codes(cq):
    import ctypes
    from common.quantity import _native
    lib = _native.lib()
    assert lib.pq_fakequant_f32(None, None, 10, 0, -128.0, 127.0, 1, None) == -1
    x = torch.zeros(8, device="cuda")
    assert lib.pq_fakequant_f32(x.data_ptr(), x.data_ptr(), 8, 0, -128.0, 127.0, 1, None) == -1   # aliasing
    assert lib.pq_fakequant_f32(x.data_ptr(), x.data_ptr(), 0, 0, -128.0, 127.0, 1, None) == 0    # empty
    k = _native.MAX_SEGMENTS + 1
    ptrs = (ctypes.c_void_p * k)(*([x.data_ptr()] * k))
    ns = (ctypes.c_uint64 * k)(*([8] * k))
    out = torch.zeros(k, dtype=torch.int32, device="cuda")
    assert lib.pq_absmax_multi_f32(ptrs, ns, k, out.data_ptr(), None) == -4
    iv = (ctypes.c_float * 1)(0.0)
    h = torch.zeros(2048, dtype=torch.int64, device="cuda")
    assert lib.pq_hist2048_multi_f32(ptrs, ns, iv, 1, h.data_ptr(), None) == -1                  # interval <= 0
    with pytest.raises(RuntimeError):
        cq.QuanDequan(8, 3)(torch.zeros(4))                                                       # CPU tensor: no fallback
    # per-channel max-abs (extension): empty is a no-op, more than 8192 channels / misaligned pointers are refused
    bits = torch.zeros(16, dtype=torch.int32, device="cuda")
    assert lib.pq_absmax_per_channel_f32(None, 0, 4, 10, bits.data_ptr(), None) == 0
    assert lib.pq_absmax_per_channel_f32(x.data_ptr(), 1, 8193, 1, bits.data_ptr(), None) == -2
    assert lib.pq_absmax_per_channel_f32(x.data_ptr() + 2, 1, 2, 2, bits.data_ptr(), None) == -3
    assert lib.pq_absmax_per_channel_f32(x.data_ptr(), 1, 2, 4, None, None) == -1
    assert not bits.any()
    # int8 max-pool: channel count must be a multiple of 16, padding smaller than the window
    q = torch.zeros((1, 4, 4, 16), dtype=torch.int8, device="cuda")
    assert lib.pq_maxpool_nhwc_s8(q.data_ptr(), q.data_ptr(), 1, 4, 4, 8, 2, 2, 0, 0, None) == -2
    assert lib.pq_maxpool_nhwc_s8(q.data_ptr(), q.data_ptr(), 1, 4, 4, 16, 2, 2, 2, 0, None) == -2
    assert lib.pq_maxpool_nhwc_s8(q.data_ptr(), q.data_ptr(), 0, 4, 4, 16, 2, 2, 0, 0, None) == -1


# ---------------------------------------------------------------- per-channel max-abs (extension of a1)
CHANNEL_CASES = [((4, 64, 56, 56), 1), ((2, 1000), 1), ((3, 7, 5, 3), 1), ((1, 2048, 7, 7), 1), ((5, 3, 224, 224), 1),
                 ((2, 16, 1, 1), 1), ((64, 3, 7, 7), 0), ((512, 2048), 0), ((6, 13, 3), 1), ((1, 1, 1), 1),
                 ((3, 8192, 2, 2), 1), ((32, 256, 56, 56), 1),
                 # the periodic variant (inner < 784, >= 16 images, channels * inner % 4 == 0) and its neighbours
                 ((64, 2048, 7, 7), 1), ((4096, 1000), 1), ((33, 12, 5, 5), 1), ((16, 8, 1, 1), 1), ((100, 6, 2), 1),
                 ((17, 4, 3, 3), 1), ((20, 7, 3), 1), ((15, 64, 7, 7), 1), ((300, 8192), 1), ((40, 3, 28, 27), 1)]


@pytest.mark.parametrize("shape,dim", CHANNEL_CASES, ids=["x".join(map(str, s)) + "_d%d" % d for s, d in CHANNEL_CASES])
def test_absmax_per_channel_vs_oracle(cq, oracle, shape, dim):
    from common.quantity import _native
    g = torch.Generator(device="cuda").manual_seed(len(shape) * 1000 + shape[-1])
    x = torch.randn(shape, device="cuda", generator=g)
    x.view(-1)[::max(1, x.numel() // 97)] *= 7.0                       # isolated peaks in various planes
    bits = torch.zeros(shape[dim], dtype=torch.int32, device="cuda")
    _native.absmax_per_channel(x, bits, dim)
    want = oracle.absmax_per_channel(x.cpu().numpy(), dim)
    assert np.array_equal(bits.view(torch.float32).cpu().numpy(), want)
    # a second batch accumulates into the running maxima; per-tensor max == max over channels
    x2 = torch.randn(shape, device="cuda", generator=g) * 1.5
    _native.absmax_per_channel(x2, bits, dim)
    want2 = oracle.absmax_per_channel(x2.cpu().numpy(), dim, cur=want)
    assert np.array_equal(bits.view(torch.float32).cpu().numpy(), want2)
    assert float(want2.max()) == float(max(x.abs().max(), x2.abs().max()))


@pytest.mark.parametrize("offset", [1, 2, 3])
def test_absmax_per_channel_unaligned_view(cq, oracle, offset):
    from common.quantity import _native
    base = torch.randn(4 * 24 * 11 * 13 + 8, device="cuda", generator=torch.Generator(device="cuda").manual_seed(offset))
    x = base[offset:offset + 4 * 24 * 11 * 13].view(4, 24, 11, 13)      # 4-byte aligned only
    bits = torch.zeros(24, dtype=torch.int32, device="cuda")
    _native.absmax_per_channel(x, bits, 1)
    assert np.array_equal(bits.view(torch.float32).cpu().numpy(), oracle.absmax_per_channel(x.cpu().numpy(), 1))


def test_collector_channel_max_extension(cq, oracle):
    names = ["a", "b"]
    g = torch.Generator(device="cuda").manual_seed(5)
    batches = [{"a": torch.randn(2, 16, 9, 9, device="cuda", generator=g),
                "b": torch.randn(2, 10, device="cuda", generator=g)} for _ in range(3)]
    dc = cq.DistributionCollector(names)
    for b in batches:
        dc.refresh_max_val({k: v.view(-1) for k, v in b.items()})
        dc.refresh_channel_max_val(b)
    got = dc.channel_max_vals
    for n in names:
        want = None
        for b in batches:
            want = oracle.absmax_per_channel(b[n].cpu().numpy(), 1, cur=want)
        assert np.array_equal(got[n], want)
        assert np.float32(got[n].max()) == dc.max_vals[n]             # consistent with the per-tensor reduction


# ------------------------------------------------------------------ INTERVAL_NUM != 2048 (tools/configs.yml:23)
@pytest.mark.parametrize("nbins", [512, 1000, 4096])
def test_other_interval_num_vs_golden(nbins):
    """DistributionCollector(interval_num=...) + Quantizer on the GPU at the reference's other bin counts: generic
    histogram kernel (pq_hist_multi_f32) and run-time-sized KL search (pq_kl_search_n_f64) against the unmodified
    reference's outputs (bins.npz)."""
    import common.quantity as cq
    from golden import gen_golden as gg
    g = load_golden("bins.npz")
    name = "t%d" % nbins
    batches = [torch.from_numpy(b).cuda() for b in gg.bins_batches()]
    col = cq.DistributionCollector([name], interval_num=nbins)
    for b in batches:
        col.refresh_max_val({name: b})
    assert float(col.max_vals[name]) == float(g[name + "/max"][0])
    iv = col.distribution_intervals[name]
    assert float(iv) == float(g[name + "/interval"][0])
    for b in batches:
        col.add_to_distributions({name: b})
    h = col.distributions[name]
    assert h.shape == (nbins,) and np.array_equal(h, g[name + "/hist"])
    qz = cq.Quantizer([name], keep_curves=True)
    qz.quantize(col.distributions, col.distribution_intervals)
    assert qz.bits[name] == int(g[name + "/bit"][0])
    assert float(qz.threshold_value[name]) == float(g[name + "/threshold_value"][0])
    np.testing.assert_allclose(qz.kl_curves.cpu().numpy()[0], g[name + "/kl"], rtol=1e-9, atol=1e-300)


def test_interval_num_limits():
    import common.quantity as cq
    from common.quantity import _native
    with pytest.raises(ValueError, match="INTERVAL_NUM"):
        cq.DistributionCollector(["t"], interval_num=8193)
    lib = _native.lib()
    c = torch.zeros((1, 128), dtype=torch.float64, device="cuda")
    ws = torch.zeros(4096, dtype=torch.float64, device="cuda")
    thr = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert lib.pq_kl_search_n_f64(c.data_ptr(), 1, 128, ws.data_ptr(), None, thr.data_ptr(), None) == -1    # PQ_EINVAL
    assert lib.pq_kl_search_n_f64(c.data_ptr(), 1, 8193, ws.data_ptr(), None, thr.data_ptr(), None) == -2   # PQ_EUNSUPPORTED
    assert lib.pq_hist_multi_f32(None, None, None, 0, 9000, None, None) == -2


def test_per_channel_absmax_vs_reference_on_slices():
    """pq_absmax_per_channel_f32 against channel_max.npz: the unmodified reference's per-tensor reduction
    (distribution_collector.py:70-78) applied to every channel slice as a tensor of its own, two batches."""
    from common.quantity import _native
    g = load_golden("channel_max.npz")
    for case in gg.CHANNEL_CASES:
        name, shape, dim, _ = case
        bits = torch.zeros(shape[dim], dtype=torch.int32, device="cuda")
        for x in gg.channel_batches(case):
            _native.absmax_per_channel(torch.from_numpy(x).cuda(), bits, dim)
        assert np.array_equal(bits.view(torch.float32).cpu().numpy(), g[name + "/max"]), name
