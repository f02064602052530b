"""CPU oracle for the calibration + simulation hot path of pytorch-quantity.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module;
the product (``pytorch-quantity_b200/``) never does and fails loudly without its
CUDA library instead.

Parity pinning: the reference ships no tests or golden vectors,
so this restatement is pinned against the UNMODIFIED reference executed in the build
container (numpy 2.3.5 / torch 2.11 semantics, NEP-50 promotion): the fixtures in
``tests/golden/`` written by ``tests/golden/gen_golden.py``.  ``tests/test_oracle.py``
checks every function here against them.

Citations are ``file:line`` relative to ``/root/reference/quantity/``;
``cq/`` = ``common/quantity/``.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

INTERVAL_NUM = 2048      # tools/configs.yml:23
TARGET_BIN = 128         # cq/quantizer.py:98


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libpq_oracle.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(
                os.path.join(_HERE, "pq_oracle.c")):
            build()
        lib = ctypes.CDLL(so)
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        i32p = ctypes.POINTER(ctypes.c_int32)
        lib.pqo_absmax_f32.argtypes = [f32p, ctypes.c_size_t, f32p]
        lib.pqo_absmax_f32.restype = None
        lib.pqo_hist_f32.argtypes = [f32p, ctypes.c_size_t, ctypes.c_float, i32p, ctypes.c_int]
        lib.pqo_hist_f32.restype = None
        lib.pqo_normalize.argtypes = [f64p, ctypes.c_int, f64p]
        lib.pqo_normalize.restype = None
        lib.pqo_kl_search.argtypes = [f64p, ctypes.c_int, ctypes.c_int, f64p]
        lib.pqo_kl_search.restype = ctypes.c_int
        lib.pqo_pairwise_sum.argtypes = [f64p, ctypes.c_long]
        lib.pqo_pairwise_sum.restype = ctypes.c_double
        lib.pqo_fakequant_f32.argtypes = [f32p, f32p, ctypes.c_size_t, ctypes.c_int]
        lib.pqo_fakequant_f32.restype = None
        _LIB = lib
    return _LIB


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1))


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


# --------------------------------------------------------------------------- a1
def absmax_update(cur_max, x):
    """cq/distribution_collector.py:77-78.  Returns python int 0 while nothing
    exceeded 0 (the reference's ``max(0, np.float32)`` keeps the int), else np.float32."""
    x = _f32(x)
    if x.size == 0:
        return cur_max
    m = np.float32(max(abs(np.max(x)), abs(np.min(x))))
    return max(cur_max, m)


def absmax_c(x, cur=0.0):
    x = _f32(x)
    out = np.array([cur], dtype=np.float32)
    _lib().pqo_absmax_f32(_p(x, ctypes.c_float), x.size, _p(out, ctypes.c_float))
    return out[0]


def absmax_per_channel(x, channel_dim=1, cur=None):
    """Extension of a1: the reference reduces per tensor only (cq/distribution_collector.py:77); this is the same
    ``max(|max x|, |min x|)`` taken per channel.  Pinned BY COMPOSITION: tests/golden/channel_max.npz holds the
    unmodified reference collector run on every channel slice as a tensor of its own (gen_golden.py::gen_channel)."""
    x = np.asarray(x, dtype=np.float32)
    axes = tuple(a for a in range(x.ndim) if a != channel_dim % x.ndim)
    m = np.maximum(np.abs(x.max(axis=axes)), np.abs(x.min(axis=axes))).astype(np.float32)
    return m if cur is None else np.maximum(np.asarray(cur, dtype=np.float32), m)


# --------------------------------------------------------------------------- a2
def interval(max_val, statistic=1, interval_num=INTERVAL_NUM):
    """cq/distribution_collector.py:60-61 under numpy 2: np.float32 when max_val is
    np.float32 (python scalars are weak), python float 1e-12 when max_val is the int 0."""
    return statistic * max_val / interval_num + 1e-12


# --------------------------------------------------------------------------- a3
def hist(x, interv, nbins=INTERVAL_NUM):
    """cq/distribution_collector.py:127-135 (C restatement)."""
    x = _f32(x)
    h = np.zeros(nbins, dtype=np.int32)
    _lib().pqo_hist_f32(_p(x, ctypes.c_float), x.size, np.float32(interv),
                        _p(h, ctypes.c_int32), nbins)
    return h


def hist_np(x, interv, nbins=INTERVAL_NUM):
    """Same statement in numpy (vectorised count instead of the python loop)."""
    x = _f32(x)
    nz = x[x != 0]
    idx = np.minimum((np.abs(nz) / np.float32(interv)).astype(np.int32), nbins - 1)
    return np.bincount(idx, minlength=nbins).astype(np.int32)


# ---------------------------------------------------------------------- a5 / a6
def normalize(counts):
    """cq/quantizer.py:95-96 -> float64[nbins]."""
    c = np.ascontiguousarray(np.asarray(counts, dtype=np.float64))
    out = np.empty_like(c)
    _lib().pqo_normalize(_p(c, ctypes.c_double), c.size, _p(out, ctypes.c_double))
    return out


def normalize_np(counts):
    counts = np.asarray(counts)
    return counts.astype(np.float32) / (counts.sum() + 1e-12)


def kl_search(P, target_bin=TARGET_BIN, want_curve=True):
    """cq/quantizer.py:98-174 -> (threshold_bin, kl[nbins-target_bin] or None)."""
    P = np.ascontiguousarray(np.asarray(P, dtype=np.float64))
    kl = np.empty(P.size - target_bin, dtype=np.float64) if want_curve else None
    t = _lib().pqo_kl_search(_p(P, ctypes.c_double), P.size, target_bin,
                             _p(kl, ctypes.c_double) if want_curve else None)
    return int(t), kl


def pairwise_sum(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return _lib().pqo_pairwise_sum(_p(a, ctypes.c_double), a.size)


# --------------------------------------------------------------------------- a7
def threshold_to_bit(threshold_bin, interv):
    """cq/quantizer.py:86-90.  ``(T + 0.5) * interval`` is an fp32 product when the
    interval is np.float32; math.log(x, 2) = log(x)/log(2) in C doubles."""
    threshold_bias = (threshold_bin + 0.5) * interv
    bit_int_d = math.ceil(math.log(threshold_bias, 2))
    return int(8 - 1 - bit_int_d), threshold_bias


def quantize_distribution(counts, interv):
    """normalize -> KL search -> bit for one tensor (cq/quantizer.py:83-90)."""
    t, _ = kl_search(normalize(counts), want_curve=False)
    bit, thr = threshold_to_bit(t, interv)
    return bit, thr, t


def maxabs_to_bit(max_val):
    """tools/pytorch_quantizer.py:651-652 (weights; raises on max_val == 0 like the reference)."""
    return int(8 - 1 - math.ceil(math.log(max_val, 2)))


# ------------------------------------------------------------------ a10 (and a12)
def fakequant(x, bit, lo=-128.0, hi=127.0):
    """cq/new_quantity_op.py:246-257."""
    x = np.asarray(x, dtype=np.float32)
    s = np.float32(2.0 ** bit)
    return (np.clip(np.rint(x * s), np.float32(lo), np.float32(hi)) / s).astype(np.float32)


def fakequant_c(x, bit):
    x = _f32(x)
    y = np.empty_like(x)
    _lib().pqo_fakequant_f32(_p(x, ctypes.c_float), _p(y, ctypes.c_float), x.size, int(bit))
    return y


def quantize_input(x, ib):
    """cq/new_quantity_op.py:48-58 -> integer-valued float32."""
    x = np.asarray(x, dtype=np.float32)
    return np.clip(np.rint(x * np.float32(2.0 ** ib)), -128.0, 127.0).astype(np.float32)


def concat_quantize(parts, ib, dim=1):
    """Concat (cq/fabu_layer.py:14-20; tools/reconstruction.py:219-238 leaves it in fp32) followed by
    the consuming layer's input quantiser (cq/new_quantity_op.py:48-58)."""
    return quantize_input(np.concatenate([np.asarray(p, dtype=np.float32) for p in parts], axis=dim), ib)


# --------------------------------------------------------------------------- a14
def right_shift(acc, rs):
    """cq/new_quantity_op.py:11-44 on exact integer accumulators (int64 array):
    v = acc / 2^rs ; round half away from zero ; clamp [-128, 127]."""
    acc = np.asarray(acc, dtype=np.int64)
    if rs >= 1:
        mag = (np.abs(acc) + (1 << (rs - 1))) >> rs
        r = np.sign(acc) * mag
    else:
        r = acc << (-rs)
    return np.clip(r, -128, 127)


# --------------------------------------------------------------------------- a13
def _quant_param(p, bit):
    """cq/new_quantity_op.py:147-152."""
    p = np.asarray(p, dtype=np.float32)
    return np.clip(np.rint(p * np.float32(2.0 ** bit)), -128.0, 127.0)


def int_conv_layer(x, weight, bias, info, stride=1, padding=0, dilation=1, groups=1):
    """cq/new_quantity_op.py:104-163 (NewConv2d) with exact integer accumulation
    (float64 conv on integers; every partial sum is far below 2^53).
    Returns (out_fp32_nchw, y_int) where y_int is the saturated int8-valued result."""
    import torch
    import torch.nn.functional as F
    wb, ib, ob = info["weight_bit"], info["input_bit"], info["output_bit"]
    bb = info["bias_bit"]
    q = quantize_input(x, ib)
    wq = _quant_param(weight, wb)
    bq = _quant_param(bias if bias is not None else np.zeros(weight.shape[0], np.float32), bb)
    acc = F.conv2d(torch.from_numpy(q.astype(np.float64)), torch.from_numpy(wq.astype(np.float64)),
                   None, stride, padding, dilation, groups).numpy()
    r = right_shift(np.rint(acc).astype(np.int64), wb + ib - ob)
    y = np.clip(r + bq.astype(np.int64)[None, :, None, None], -128, 127)
    return (y.astype(np.float32) / np.float32(2.0 ** ob)).astype(np.float32), y


def int_linear_layer(x, weight, bias, info):
    """cq/new_quantity_op.py:177-236 (NewLinear)."""
    wb, ib, ob = info["weight_bit"], info["input_bit"], info["output_bit"]
    bb = info["bias_bit"]
    q = quantize_input(x, ib).astype(np.int64)
    wq = _quant_param(weight, wb).astype(np.int64)
    bq = _quant_param(bias if bias is not None else np.zeros(weight.shape[0], np.float32), bb)
    acc = q @ wq.T
    r = right_shift(acc, wb + ib - ob)
    y = np.clip(r + bq.astype(np.int64)[None, :], -128, 127)
    return (y.astype(np.float32) / np.float32(2.0 ** ob)).astype(np.float32), y


def add_clamp(x, y):
    """cq/new_quantity_op.py:166-174 (NewAdd)."""
    return np.clip(np.asarray(x, np.float32) + np.asarray(y, np.float32),
                   np.float32(-128.0), np.float32(127.0))


# --------------------------------------------------------------------------- a16
def merge_bn_params(weight, bias, gamma, beta, mean, var):
    """cq/utils.py:37-42 in fp32 with separate mul/add (no FMA)."""
    import torch
    w = torch.as_tensor(weight, dtype=torch.float32)
    b = torch.zeros(w.shape[0]) if bias is None else torch.as_tensor(bias, dtype=torch.float32)
    g, bt = torch.as_tensor(gamma), torch.as_tensor(beta)
    mu, v = torch.as_tensor(mean), torch.as_tensor(var)
    tmp = g / torch.sqrt(v + 1e-5)
    return (tmp.view(-1, 1, 1, 1) * w).numpy(), (tmp * (b - mu) + bt).numpy()


# ---------------------------------------------------------------------- a8 / a9
def weight_quantize(params, dkl=False):
    """tools/pytorch_quantizer.py:635-669.  dkl=False: bit from max-abs (:650-653, the shipped setting
    ``_DKL_weight = False``, :62); dkl=True: histogram + KL search per parameter (:644-648).
    params: ordered dict name -> float32 ndarray.  Returns (bits, q_int32 arrays)."""
    bits, q = {}, {}
    for name, p in params.items():
        p32 = np.asarray(p, dtype=np.float32)
        m = absmax_update(0, p32)
        if dkl:
            interv = interval(m)
            bit = quantize_distribution(hist(p32, interv), interv)[0]
        else:
            bit = maxabs_to_bit(m)
        v = np.clip(np.around(p32.reshape(-1) * math.pow(2, bit)), -128, 127)
        bits[name] = bit
        q[name] = v.reshape(p32.shape).astype(np.int32)
    return bits, q


def rescale_wrap(q, old_bit, new_bit):
    """tools/rewriter.py:53-55,119-122: around(q / 2^old * 2^new).astype(int8) -- wraps."""
    lines = np.array(q, dtype=np.float32)
    lines = lines / 2 ** old_bit * 2 ** new_bit
    return np.around(lines).astype(np.int8)


def max_shift_limit(feat_bits, infeat_bits, weight_bits, max_shift=12):
    """tools/rewriter.py:75-103 -> (need_rewrite, new_weight_bits)."""
    need, new = False, {}
    for name, wb in weight_bits.items():
        ib, ob = int(infeat_bits[name][0]), feat_bits[name]
        nb = wb
        if wb + ib - ob > max_shift:
            nb = max_shift - ib + ob
            need = True
        new[name] = nb
    return need, new


# ---------------------------------------------------------------------------- a4
def calibrate(batches, top_names, net_info, merge_groups, statistic=1,
              interval_num=INTERVAL_NUM, return_all=False):
    """tools/pytorch_quantizer.py:379-465 given the hooked tensors.

    batches: list of dict name -> ndarray (one dict per calibration batch, the
    ``named_feats`` of :387-389).  top_names: ['image'] + cared tracer names.
    net_info: name -> {'inputs': [...], 'type': ...}.  merge_groups: :298-341.
    Returns bits dict (and the intermediates when return_all)."""
    max_vals = {n: 0 for n in top_names}
    for feats in batches:                                              # pass 1  :379-390
        for n in top_names:
            max_vals[n] = absmax_update(max_vals[n], feats[n])
    intervals = {n: interval(max_vals[n], statistic, interval_num) for n in top_names}

    def has_eltwise(group):
        return any(net_info[m]["type"] == "Eltwise" for m in group)

    for group in merge_groups:                                         # :396-411
        assert len(group) > 1
        if has_eltwise(group):
            continue
        top = 0
        for m in group:
            top = max(top, intervals[m])
        for m in group:
            intervals[m] = top
    hists = {n: np.zeros(interval_num, dtype=np.int32) for n in top_names}
    for feats in batches:                                              # pass 2  :415-426
        for n in top_names:
            hists[n] += hist(feats[n], intervals[n], interval_num)
    dists = dict(hists)
    for group in merge_groups:                                         # :432-445
        if has_eltwise(group):
            continue
        tmp = np.zeros(interval_num)
        for m in group:
            tmp += dists[m]
        for m in group:
            dists[m] = tmp
    bits, thresholds, tbins = {}, {}, {}
    for n in top_names:                                                # :448
        bits[n], thresholds[n], tbins[n] = quantize_distribution(dists[n], intervals[n])
    raw_bits = dict(bits)
    for group in merge_groups:                                         # :453-465
        elt_idx, found = 0, False
        for i, m in enumerate(group):
            if net_info[m]["type"] == "Eltwise":
                elt_idx, found = i, True
        if found:
            bits[group[1 - elt_idx]] = bits[group[elt_idx]]
    if return_all:
        return dict(bits=bits, raw_bits=raw_bits, max_vals=max_vals, intervals=intervals,
                    hists=hists, dists=dists, thresholds=thresholds, tbins=tbins)
    return bits


def feat_table_lines(top_names, cared_layer_names, net_info, bits):
    """tools/pytorch_quantizer.py:468-485."""
    lines, first = [], True
    for i, n in enumerate(top_names):
        if n == "image":
            s = "image " + str(bits["image"])
        elif first:
            s = cared_layer_names[i - 1] + " " + str(bits[n]) + " " + str(bits["image"])
            first = False
        else:
            assert len(net_info[n]["inputs"]) > 0
            s = cared_layer_names[i - 1] + " " + str(bits[n])
            for inp in net_info[n]["inputs"]:
                s += " " + str(bits[inp])
        lines.append(s)
    return lines
