#!/usr/bin/env python
"""bench_sim.py -- secondary benchmark: the simulation side of the hot path (BASELINE.json configs 2 and 4).

    python bench_sim.py [--model r18|r50] [--mode test|model|both] [--batch B] [--iters K]

  ReconTest  (config 2): ResNet-18 224x224 fake-quant inference, batch 256.  Kernel of interest:
             pq_fakequant_f32 (HBM-bound, 8 algorithmic bytes per element).
  ReconModel (config 4): ResNet-50 224x224 integer-simulation inference, batch 512.  Kernels of
             interest: pq_conv2d_s8 / pq_gemm_s8 (tcgen05 kind::i8) and pq_quantize_nchw_to_nhwc_s8.
Prints one JSON line per mode.  The tables come from a short calibration of the same synthetic model
through this repo's own tools.Quantity (4 batches of 16 images).  bench.py stays the headline bench.
"""
import argparse
import json
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "pytorch-quantity_b200")
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def build(model_name):
    from model.resnet.resnet_fabu import randomize_bn_, resnet18_fabu, resnet50_fabu
    torch.manual_seed(0)
    net = (resnet18_fabu if model_name == "r18" else resnet50_fabu)().eval()
    with torch.no_grad():
        randomize_bn_(net, 0)
    return net


def configs(workdir, n_batches):
    import tools._config as tc
    cfg = tc.load_tool_config(os.path.join(os.path.dirname(tc.__file__), "configs.yml"))
    cfg["OUTPUT"] = {"WORK_DIR": workdir, "WEIGHT_BIT_TABLE": workdir + "/weight.table",
                     "FEAT_BIT_TABLE": workdir + "/feat.table", "WEIGHT_DIR": workdir + "/weight",
                     "BIAS_DIR": workdir + "/bias", "FINAL_WEIGHT_DIR": workdir + "/new_weight",
                     "FINAL_BIAS_DIR": workdir + "/new_bias"}
    cfg["SETTINGS"]["MAX_CALI_IMG_NUM"] = n_batches - 1
    user = tc.load_user_config({"PATH": {}, "MODEL": {"INPUT_SHAPE": "1,3,224,224"},
                                "PRE_PROCESS": {"IMG": 1}, "SETTINGS": {"DEVICE": "gpu", "GPU": 0}})
    return cfg, user


def int8_peak_tops():
    """Measured dense int8 tensor throughput on this GPU: torch._int_mm 8192^3 (library GEMM)."""
    a = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device="cuda")
    b = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device="cuda")
    for _ in range(3):
        torch._int_mm(a, b)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        s.record(); torch._int_mm(a, b); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return 2 * 8192 ** 3 / (best * 1e-3) / 1e12


def timed_forward(model, x, iters):
    """ms per forward (clean, CUDA events around `iters` forwards) + per-kernel statistics from a second,
    instrumented set of forwards (an event pair around every C-ABI launch)."""
    from common.quantity import _native
    with torch.no_grad():
        for _ in range(3):
            model(x)
        torch.cuda.synchronize()
        l0 = _native.LAUNCHES["total"]
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            y = model(x)
        e.record()
        torch.cuda.synchronize()
        launches = (_native.LAUNCHES["total"] - l0) // iters
        ms = s.elapsed_time(e) / iters
        prof = {}
        _native.set_profile(prof)
        for _ in range(iters):
            model(x)
        torch.cuda.synchronize()
        _native.set_profile(None)
    stats = {}
    for name, rows in prof.items():
        t = [a.elapsed_time(b) for a, b, _ in rows]
        nb = [c for _, _, c in rows]
        stats[name] = {"launches_per_fwd": len(rows) // iters, "ms_per_fwd": sum(t) / iters,
                       "alg_bytes_per_fwd": sum(nb) / iters}
    return ms, stats, launches, y


def timed_graph(model, x, iters):
    """The same forward captured once and replayed as a CUDA graph."""
    from common.quantity import GraphedForward
    fwd = GraphedForward(model, x)
    for _ in range(3):
        fwd(x)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        y = fwd(x)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters, fwd.launches, y.clone()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default=None)
    ap.add_argument("--mode", default="both", choices=["test", "model", "both"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--no-int8-pipeline", action="store_true",
                    help="ReconModel: keep the reference's fp32 NCHW module boundaries only")
    args = ap.parse_args()
    real_stdout, sys.stdout = sys.stdout, sys.stderr
    import tools
    from common.quantity import merge_bn
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    peak_hbm = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else 6650.0
    lines = []
    for mode in (["test", "model"] if args.mode == "both" else [args.mode]):
        model_name = args.model or ("r18" if mode == "test" else "r50")
        batch = args.batch or (256 if mode == "test" else 512)
        workdir = tempfile.mkdtemp(prefix="pq_sim_")
        cfg, user = configs(workdir, 4)
        with torch.no_grad():
            net = merge_bn(build(model_name), "cpu")
            q = tools.Quantity(net, config=cfg, user_config=user, verbose=False)
            cal = []
            for i in range(4):
                g = torch.Generator().manual_seed(1 + i)
                cal.append((torch.randn(16, 3, 224, 224, generator=g), None))
            q.activation_quantize(cal)
            q.weight_quantize()
            net2 = build(model_name)
            r = tools.Reconstruction(net2, config=cfg)
            r.merge_bn()
            info = r.get_quantity_information()
            model = (r.ReconTest if mode == "test" else r.ReconModel)(info, None).cuda().eval()
        x = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(7)).cuda()
        if mode == "test":   # the caller fake-quantises the image itself (resnet_reconstruction.py:135)
            from common.quantity import QuanDequan
            x = QuanDequan(8, info["image"]["output_bit"])(x)
        variants = [("fp32 module boundaries (reference semantics)", False)]
        if mode == "model" and not args.no_int8_pipeline:
            variants.append(("int8 inter-layer pipeline (bit-identical output)", True))
        y_ref = None
        for label, pipe in variants:
            if mode == "model":
                from common.quantity import enable_int8_pipeline
                enable_int8_pipeline(model, pipe)
            ms, stats, launches, y = timed_forward(model, x, args.iters)
            if y_ref is None:
                y_ref = y
            assert torch.equal(y, y_ref), "pipeline output differs"
            line = emit_line(mode, model_name, batch, label, ms, stats, launches, peak_hbm)
            assert torch.isfinite(y).all()
            lines.append(line)
            if pipe:
                gms, glaunches, gy = timed_graph(model, x, args.iters)
                assert torch.equal(gy, y_ref), "graph replay output differs"
                lines.append(emit_line(mode, model_name, batch, label + ", one CUDA graph per forward", gms, {},
                                       glaunches, peak_hbm))
    sys.stdout = real_stdout
    for line in lines:
        print(json.dumps(line), flush=True)


def emit_line(mode, model_name, batch, label, ms, stats, launches, peak_hbm):
    line = {"metric": "ReconTest images/sec" if mode == "test" else "ReconModel images/sec",
            "value": round(batch / (ms * 1e-3), 1), "unit": "images/s", "ms_per_forward": round(ms, 3),
            "config": {"workload": "%s 224x224 %s, batch %d" % (model_name, "ReconTest" if mode == "test" else "ReconModel", batch),
                       "variant": label},
            "gpu_launches_per_forward": launches, "kernels": {}}
    for name, st in stats.items():
        k = dict(st)
        k["ms_per_fwd"] = round(k["ms_per_fwd"], 4)
        if name in ("fakequant", "add_clamp", "quantize_s8", "relu_s8", "maxpool_s8", "add_requant"):
            k["GBps"] = round(st["alg_bytes_per_fwd"] / (st["ms_per_fwd"] * 1e-3) / 1e9, 1)
            k["frac_of_hbm_peak"] = round(k["GBps"] / peak_hbm, 4)
        line["kernels"][name] = k
    if mode == "model":
        peak = int8_peak_tops()
        line["int8_peak_TOPS_measured"] = round(peak, 1)
        for name in ("conv_s8", "gemm_s8", "conv_add_s8"):
            if name in stats and stats[name]["ms_per_fwd"] > 0:
                tops = stats[name]["alg_bytes_per_fwd"] / (stats[name]["ms_per_fwd"] * 1e-3) / 1e12
                line["kernels"][name]["int8_ops_per_fwd"] = line["kernels"][name].pop("alg_bytes_per_fwd")
                line["kernels"][name]["TOPS"] = round(tops, 1)
                line["kernels"][name]["frac_of_int8_peak"] = round(tops / peak, 4)
    return line


if __name__ == "__main__":
    main()
