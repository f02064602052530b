#!/usr/bin/env python
"""bench_stats.py -- statistics microbenchmark (BASELINE.json config 5, SURVEY.md 8d "C5").

    python bench_stats.py [--max-log2 32] [--min-log2 28] [--iters 5] [--skip-stats] [--skip-channel] [--skip-fakequant]

Part 1: max-abs + 2048-bin histogram + KL search over ONE fp32 tensor of 2^28 ... 2^32 elements
        (1 - 16 GB) for the four input families SURVEY 8d names:
          dense     N(0,1), signed                   (what a convolution output looks like)
          relu      relu(N(0,1)), 50 % exact zeros   (the post-ReLU / 'image' case)
          const     1.0 everywhere                   (every element in ONE bin: worst-case atomics)
          outlier   Laplace + one 1e4 outlier        (everything lands in bins 0-3)
        Both statistics kernels are HBM-bound at 4 algorithmic bytes per element; the KL search
        reads 16 KB per tensor and is reported in microseconds.
Part 1b: per-channel max-abs (extension kernel) over activation-, FC- and weight-shaped tensors.
Part 2: fake-quant bandwidth sweep, 2^20 ... 2^32 elements x bit in {-2, 0, 4, 7, 12},
        8 algorithmic bytes per element.  Sizes whose working set would fit in the 126 MB L2 rotate
        over enough buffers to exceed 2x L2, so every number is an HBM number.

Each case also checks a size-independent property on the full tensor (sum of counts == number of
non-zero elements; fake-quant is idempotent), so a fast-but-wrong kernel cannot post a number.
Prints one JSON line per case; timing is CUDA events on the launching stream after 2 warm-ups.
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "pytorch-quantity_b200")
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

L2_BYTES = 126 * 1024 * 1024
FAMILIES = ("dense", "relu", "const", "outlier")
SLAB = 1 << 26   # fill 256 MB at a time so the generators never hold a second full-size temporary


def make_input(kind, n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    for lo in range(0, n, SLAB):
        part = x[lo:lo + SLAB]
        if kind == "const":
            part.fill_(1.0)
        elif kind == "outlier":
            part.exponential_(1.0, generator=g)
            sign = torch.empty_like(part).uniform_(-1.0, 1.0, generator=g).sign_()
            part.mul_(sign)
            del sign
        else:
            part.normal_(0.0, 1.0, generator=g)
            if kind == "relu":
                part.clamp_(min=0.0)
    if kind == "outlier":
        x[n // 3] = 1.0e4
    return x


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def stats_case(kind, log2n, iters, peak):
    from common.quantity import _native
    n = 1 << log2n
    x = make_input(kind, n, 1234 + log2n)
    max_bits = torch.zeros(1, dtype=torch.int32, device="cuda")
    ms_max = timed(lambda: _native.absmax_multi([x], max_bits), iters)
    mx = max_bits.view(torch.float32).cpu().numpy()[0]
    interval = np.float32(1 * mx / 2048 + 1e-12)           # distribution_collector.py:60-61
    hist = torch.zeros((1, 2048), dtype=torch.int64, device="cuda")
    ms_hist = timed(lambda: _native.hist_multi([x], [interval], hist), iters)
    # property check on the whole tensor: every non-zero element was counted exactly once per launch
    hist.zero_()
    _native.hist_multi([x], [interval], hist)
    nonzero = 0
    for lo in range(0, n, SLAB):
        nonzero += int(torch.count_nonzero(x[lo:lo + SLAB]))
    total = int(hist.sum())
    assert total == nonzero, "histogram lost elements: %d counted, %d non-zero" % (total, nonzero)
    assert float(mx) == float(max(abs(float(x.max())), abs(float(x.min())))), "max-abs differs from torch"
    counts = hist.to(torch.float64)
    ms_kl = timed(lambda: _native.kl_search(counts), iters)
    thr, _ = _native.kl_search(counts)
    gbs_max = 4.0 * n / (ms_max * 1e-3) / 1e9
    gbs_hist = 4.0 * n / (ms_hist * 1e-3) / 1e9
    occupied = int((hist[0] != 0).sum())
    del x
    return {"bench": "stats", "family": kind, "log2_elements": log2n, "GB": round(4.0 * n / 2 ** 30, 2),
            "absmax_ms": round(ms_max, 4), "absmax_GBps": round(gbs_max, 1), "absmax_frac": round(gbs_max / peak, 4),
            "hist_ms": round(ms_hist, 4), "hist_GBps": round(gbs_hist, 1), "hist_frac": round(gbs_hist / peak, 4),
            "kl_us": round(ms_kl * 1e3, 1), "threshold_bin": int(thr[0]), "occupied_bins": occupied,
            "nonzero_fraction": round(nonzero / n, 4), "counts_check": "sum(counts) == count_nonzero(x)"}


def channel_case(shape, dim, iters, peak):
    """Per-channel max-abs (extension of a1): 4 algorithmic bytes per element; checked against torch.amax."""
    from common.quantity import _native
    n = int(np.prod(shape))
    x = make_input("dense", n, 4321).view(shape)
    bits = torch.zeros(shape[dim], dtype=torch.int32, device="cuda")
    ms = timed(lambda: _native.absmax_per_channel(x, bits, dim), iters)
    want = x.abs().amax(dim=[d for d in range(len(shape)) if d != dim])
    assert torch.equal(bits.view(torch.float32), want), "per-channel max-abs differs from torch.amax"
    gbs = 4.0 * n / (ms * 1e-3) / 1e9
    del x
    return {"bench": "absmax_per_channel", "shape": list(shape), "channel_dim": dim, "GB": round(4.0 * n / 2 ** 30, 3),
            "ms": round(ms, 4), "GBps": round(gbs, 1), "frac": round(gbs / peak, 4)}


def fakequant_case(log2n, bit, iters, peak):
    from common.quantity import _native
    n = 1 << log2n
    copies = max(1, -(-2 * L2_BYTES // (8 * n)))           # rotate so reads + writes exceed 2x L2
    copies = min(copies, 64)
    xs = [torch.randn(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(77 + i))
          if n <= SLAB else make_input("dense", n, 77 + i) for i in range(copies)]
    state = {"i": 0}

    def run():
        y = _native.fakequant(xs[state["i"] % copies], bit)
        state["i"] += 1
        return y

    ms = timed(run, max(iters, 2 * copies) if n <= SLAB else iters)
    y = _native.fakequant(xs[0], bit)
    probe = slice(0, min(n, SLAB))
    assert torch.equal(_native.fakequant(y[probe].clone(), bit), y[probe]), "fake-quant is not idempotent"
    scaled = y[probe] * float(2.0 ** bit)
    assert torch.equal(scaled, scaled.round()) and float(scaled.abs().max()) <= 128.0, "output off the int8 grid"
    gbs = 8.0 * n / (ms * 1e-3) / 1e9
    del xs, y
    return {"bench": "fakequant", "log2_elements": log2n, "bit": bit, "ms": round(ms, 5), "GBps": round(gbs, 1),
            "frac": round(gbs / peak, 4), "rotating_buffers": copies}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=28)
    ap.add_argument("--max-log2", type=int, default=32)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--families", default=",".join(FAMILIES))
    ap.add_argument("--skip-stats", action="store_true")
    ap.add_argument("--skip-fakequant", action="store_true")
    ap.add_argument("--skip-channel", action="store_true")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_stats.py needs a CUDA device (no CPU fallback)")
    pk = os.path.join(REPO, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    print(json.dumps({"bench": "header", "hbm_peak_GBps": peak,
                      "peak_source": "MEASURED_PEAKS.json" if os.path.exists(pk) else "B200_PROFILING.md fallback",
                      "gpu": torch.cuda.get_device_name(0)}), flush=True)
    if not args.skip_stats:
        for log2n in range(args.min_log2, args.max_log2 + 1):
            for kind in args.families.split(","):
                print(json.dumps(stats_case(kind, log2n, args.iters, peak)), flush=True)
                torch.cuda.empty_cache()
    if not args.skip_channel:
        for shape, dim in (((64, 256, 128, 128), 1), ((256, 64, 112, 112), 1), ((512, 512, 28, 28), 1),
                           ((2674, 2048, 7, 7), 1), ((65536, 1000), 1), ((2048, 512, 3, 3), 0)):
            print(json.dumps(channel_case(shape, dim, args.iters, peak)), flush=True)
            torch.cuda.empty_cache()
    if not args.skip_fakequant:
        for log2n in range(20, args.max_log2 + 1, 2):
            for bit in (-2, 0, 4, 7, 12):
                print(json.dumps(fakequant_case(log2n, bit, args.iters, peak)), flush=True)
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
