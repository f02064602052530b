"""Run the UNMODIFIED reference (staged copy under baseline/_ref, see stage_ref.py) in its own process,
on the CPU or -- with ``--device gpu`` -- on the same B200 this repository's kernels run on
(quantity/test/user_configs.yml:24 ``DEVICE: gpu``, quantity/tools/pytorch_quantizer.py:43-45,291-292).

    python baseline/ref_runner.py --model r18 --device gpu --calib 8x8 --recon ReconTest,ReconModel \
        --eval 8 --dump-layers --out /tmp/ref_r18

It exists so that every end-to-end check can be GPU-vs-GPU with zero tolerance (same cuDNN forward in both
arms) and so that bench.py can time the reference's own modules "eager on this box" next to this
repository's kernels.  A separate process is required because the reference and the drop-in deliberately
share their top-level package names (``common``, ``tools``).

What the harness does to the reference: NOTHING to its arithmetic.  Import-time shims only
(tests/golden/ref_loader.py), the cwd-relative yml files the drivers read are staged in a scratch dir, and
for GPU runs the plain-attribute tensor ``quantized_bias`` of NewConv2d / NewLinear is moved to the device
(quirk Q5: new_quantity_op.py:163 -- it is not a Parameter/buffer, so ``model.cuda()`` leaves it behind).
Instrumentation records what the reference computed (maxima, intervals, merged histograms, thresholds,
bits) by wrapping methods, never by editing them.

Outputs under --out: result.json (tables, md5 of every JSON the reference wrote, net_info, bits, timings),
arrays.npz (maxima, histograms), layers/<mode>/<layer>.npy (per-layer outputs with --dump-layers).
Test / bench infrastructure; never imported by the product.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
TESTS = os.path.join(REPO, "tests")
GOLDEN = os.path.join(TESTS, "golden")

LAYER_TYPES = ("NewConv2d", "NewLinear", "NewAdd", "TestConv", "TestLinear")


def _reference_root():
    staged = os.path.join(HERE, "_ref")
    if os.path.isdir(os.path.join(staged, "quantity", "common", "quantity")):
        return staged
    return os.environ.get("PQ_REFERENCE_ROOT", "/root/reference")


def read_workdir(test_dir, with_values=False):
    """Tables as text, every JSON the drivers wrote as md5 (the files hold up to 2.4 M integers)."""
    wd = os.path.join(test_dir, "workdir")
    snap = {}
    for fn in ("feat.table", "weight.table"):
        p = os.path.join(wd, fn)
        if os.path.exists(p):
            snap[fn] = open(p).read()
    for sub in ("weight", "bias", "new_weight", "new_bias"):
        d = os.path.join(wd, sub)
        if os.path.isdir(d):
            for fn in sorted(os.listdir(d)):
                snap[sub + "/" + fn] = hashlib.md5(open(os.path.join(d, fn), "rb").read()).hexdigest()
    return snap


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", required=True, choices=["tiny", "lenet", "r18", "r50"])
    ap.add_argument("--device", default="cpu", choices=["cpu", "gpu"])
    ap.add_argument("--out", required=True)
    ap.add_argument("--calib", default="", help="NxB: N batches of B images through Quantity.activation_quantize "
                                                "+ weight_quantize + the script's second rewrite_weight")
    ap.add_argument("--calib-only-activations", action="store_true", help="skip weight_quantize (bench arm)")
    ap.add_argument("--warmup-batches", type=int, default=0, help="untimed forwards before the calibration job")
    ap.add_argument("--tables-from", default="", help="directory holding feat.table / weight.table to rebuild from")
    ap.add_argument("--recon", default="", help="comma list of ReconModel,ReconTest")
    ap.add_argument("--eval", type=int, default=8, help="batch size of the evaluation forward")
    ap.add_argument("--dump-layers", action="store_true")
    ap.add_argument("--time-forward", type=int, default=0, help="time this many forwards of each rebuilt model")
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--exact-batch", type=int, default=0,
                    help="':nocudnn' variants forward only the first N images of the evaluation batch (ATen's GEMM "
                         "convolution loops over the samples and is slow at batch 512)")
    ap.add_argument("--self-check", action="store_true",
                    help="per NewConv2d / NewLinear: compare the reference's own fp32 accumulator with an exact float64 "
                         "evaluation of the same integer operands (is the library's fp32 conv exact on this device?)")
    ap.add_argument("--no-cudnn", action="store_true",
                    help="run the rebuilt models with torch.backends.cudnn disabled (ATen GEMM convolution: exact on "
                         "integer operands, unlike cuDNN's fp32 Winograd / FFT algorithms)")
    ap.add_argument("--common", default="reference", choices=["reference", "ours"],
                    help="'ours': the reference's unmodified tools/ drivers on top of THIS repository's "
                         "common.quantity (SURVEY 8b boundary proof: numpy tensors in, CUDA kernels underneath)")
    args = ap.parse_args()

    os.environ["PQ_REFERENCE_ROOT"] = _reference_root()
    for p in (TESTS, GOLDEN):
        if p not in sys.path:
            sys.path.insert(0, p)
    import numpy as np
    import torch
    import ref_loader
    import ref_models
    assert ref_loader.available(), "no staged reference under %s (run baseline/stage_ref.py)" % ref_loader.REF_ROOT
    ours_common = args.common == "ours"
    if ours_common:
        pkg = os.path.join(REPO, "pytorch-quantity_b200")
        sys.path.insert(0, pkg)
        import common.quantity as _cq                     # noqa: F401  bound before the reference tree is on sys.path
        sys.path.remove(pkg)

    ref_models.set_deterministic()
    gpu = args.device == "gpu"
    if gpu:
        assert torch.cuda.is_available(), "--device gpu needs a GPU"
    os.makedirs(args.out, exist_ok=True)
    name = args.model
    n_batches, batch = (int(v) for v in args.calib.split("x")) if args.calib else (1, 1)
    result = {"model": name, "device": args.device, "reference_root": ref_loader.REF_ROOT,
              "cpu_count": os.cpu_count(), "torch": torch.__version__, "numpy": np.__version__}
    arrays = {}
    workers = args.workers or max(1, min(os.cpu_count() or 1, 30 if name == "r18" else 71 if name == "r50" else 2))
    result["worker_num"] = workers

    with ref_loader.reference_tools(ref_models.INPUT_SHAPE[name], max_cali=n_batches - 1, worker_num=workers,
                                    extra_sys_path=(TESTS,), device=args.device,
                                    foreign_common=ours_common) as (tools, test_dir):
        from common.quantity import merge_bn              # the reference's (sys.path[0] = staged quantity/)
        import common.quantity as ref_cq
        assert ours_common != os.path.abspath(ref_cq.__file__).startswith(os.path.abspath(ref_loader.REF_ROOT)), \
            ref_cq.__file__
        assert os.path.abspath(tools.__file__).startswith(os.path.abspath(ref_loader.REF_ROOT)), tools.__file__
        result["tools_file"] = tools.__file__
        result["common_quantity_file"] = ref_cq.__file__
        pq = sys.modules["tools.pytorch_quantizer"]
        rec = {}

        orig_c, orig_q = pq.DistributionCollector, pq.Quantizer
        orig_refresh, orig_quantize = orig_c.refresh_max_val, orig_q.quantize

        def rec_refresh(self, tensors):
            rec.setdefault("collector", self)      # the first one is the activation collector (weights get their own)
            return orig_refresh(self, tensors)

        def rec_quantize(self, distributions, distribution_intervals):
            rec["dists"] = {k: np.array(v, copy=True) for k, v in distributions.items()}
            rec["intervals"] = dict(distribution_intervals)
            r = orig_quantize(self, distributions, distribution_intervals)
            rec["raw_bits"] = dict(self.bits)
            rec["thresholds"] = dict(self.threshold_value)
            return r

        def prepared_model():
            net = merge_bn(ref_models.build_model(name), "cpu")
            net.eval()
            return net.cuda() if gpu else net

        if args.calib:
            orig_c.refresh_max_val, orig_q.quantize = rec_refresh, rec_quantize
            try:
                with torch.no_grad():
                    net = prepared_model()
                    batches = ref_models.calib_batches(name, n_batches, batch)
                    q = tools.Quantity(net)
                    for i in range(args.warmup_batches):
                        q.net_forward(net, batches[i % n_batches])
                    if gpu:
                        torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    q.activation_quantize(batches)
                    if gpu:
                        torch.cuda.synchronize()
                    result["seconds_activation_quantize"] = time.perf_counter() - t0
                    result["images"] = n_batches * batch
                    if not args.calib_only_activations:
                        t0 = time.perf_counter()
                        q.weight_quantize()
                        result["seconds_weight_quantize"] = time.perf_counter() - t0
                        result["after_weight_quantize"] = read_workdir(test_dir)
                        q.rewrite_weight()                 # the example script's second call (quirk Q7)
                        result["after_second_rewrite"] = read_workdir(test_dir)
                    else:
                        result["after_weight_quantize"] = read_workdir(test_dir)
            finally:
                orig_c.refresh_max_val, orig_q.quantize = orig_refresh, orig_quantize
            result["net_info"] = {k: {"inputs": v["inputs"], "type": v["type"]} for k, v in q.net_info.items()}
            result["cared_op_layer_names"] = q.cared_op_layer_names
            result["merge_groups"] = q.get_merge_groups(q.net_info)
            result["raw_bits"] = rec["raw_bits"]
            result["thresholds"] = {k: float(v) for k, v in rec["thresholds"].items()}
            result["intervals"] = {k: float(v) for k, v in rec["intervals"].items()}
            result["max_vals"] = {k: float(v) for k, v in rec["collector"].max_vals.items()}
            for k, v in rec["dists"].items():
                arrays["dist/" + k] = v
        elif args.tables_from:
            wd = os.path.join(test_dir, "workdir")
            os.makedirs(wd, exist_ok=True)
            for fn in ("feat.table", "weight.table"):
                shutil.copy(os.path.join(args.tables_from, fn), os.path.join(wd, fn))

        for spec in [m for m in args.recon.split(",") if m]:
            # "ReconModel" | "ReconTest" | "ReconModel:nocudnn" (that one forward with cuDNN disabled)
            mode, _, variant = spec.partition(":")
            no_cudnn = args.no_cudnn or variant == "nocudnn"
            with torch.no_grad():
                net = ref_models.build_model(name)
                r = tools.Reconstruction(net)
                r.merge_bn()
                net.eval()
                info = r.get_quantity_information()
                result.setdefault("quantity_information", {
                    k: {kk: vv for kk, vv in v.items() if kk != "layer"} for k, v in info.items()})
                t0 = time.perf_counter()
                model = getattr(r, mode)(info, os.path.join(test_dir, "workdir", mode + ".pth"))
                result["seconds_build_" + mode] = time.perf_counter() - t0
                recon_gpu = gpu or ours_common                       # the drop-in operators have no CPU path
                if recon_gpu:
                    model = model.cuda()
                    for mod in model.modules():                      # quirk Q5
                        if hasattr(mod, "quantized_bias") and isinstance(mod.quantized_bias, torch.Tensor):
                            mod.quantized_bias = mod.quantized_bias.cuda()
                x = ref_models.eval_batch(name, args.eval)
                x = x.cuda() if recon_gpu else x
                hooks, layer_md5 = [], {}
                layer_dir = os.path.join(args.out, "layers", spec.replace(":", "_"))
                if args.dump_layers:
                    os.makedirs(layer_dir, exist_ok=True)

                self_check = {}

                def exactness(lname, m, xin):
                    """fp32 accumulator the reference's layer computes vs the exact float64 one."""
                    import torch.nn.functional as F
                    xq = m.Quan(xin)
                    if type(m).__name__ == "NewConv2d":
                        c = m.Conv
                        got = c(xq)
                        with torch.backends.cudnn.flags(enabled=False):
                            want = F.conv2d(xq.double(), c.weight.double(), None, c.stride, c.padding, c.dilation,
                                            c.groups)
                    else:
                        got = m.Linear(xq)
                        want = F.linear(xq.double(), m.Linear.weight.double())
                    bad = got.double() != want
                    self_check[lname] = {"inexact_fraction": float(bad.double().mean()),
                                         "max_abs_err": float((got.double() - want).abs().max()),
                                         "max_abs_acc": float(want.abs().max())}

                def keep(lname):
                    def _h(m, i, o):
                        a = (o.detach() + 0.0).cpu().numpy()        # + 0.0: -0.0 and 0.0 compare equal, hash alike
                        layer_md5[lname] = hashlib.md5(a.tobytes()).hexdigest()
                        if args.dump_layers:
                            np.save(os.path.join(layer_dir, lname + ".npy"), a)
                        if args.self_check and type(m).__name__ in ("NewConv2d", "NewLinear"):
                            exactness(lname, m, i[0])
                    return _h

                for lname, mod in model.named_modules():
                    if type(mod).__name__ in LAYER_TYPES:
                        hooks.append(mod.register_forward_hook(keep(lname)))
                import contextlib

                def lib():
                    return torch.backends.cudnn.flags(enabled=False) if no_cudnn else contextlib.nullcontext()
                if variant == "nocudnn" and args.exact_batch:
                    x = x[:args.exact_batch]
                with lib():
                    y = model(x.clone())
                for h in hooks:
                    h.remove()
                if args.self_check:
                    result[spec + "/self_check"] = self_check
                arrays[spec + "/y"] = y.cpu().numpy()
                result[spec + "/layer_md5"] = layer_md5
                if args.time_forward and variant != "nocudnn":      # the stock (cuDNN) path is what gets timed
                    times = []
                    for it in range(args.time_forward + 2):
                        if gpu:
                            torch.cuda.synchronize()
                            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            s.record()
                            with lib():
                                model(x)
                            e.record()
                            torch.cuda.synchronize()
                            ms = s.elapsed_time(e)
                        else:
                            t0 = time.perf_counter()
                            model(x)
                            ms = (time.perf_counter() - t0) * 1e3
                        if it >= 2:
                            times.append(ms)
                    result[spec + "/forward_ms"] = times
                del model, net, r
                if gpu:
                    torch.cuda.empty_cache()

    np.savez(os.path.join(args.out, "arrays.npz"), **arrays)
    with open(os.path.join(args.out, "result.json"), "w") as f:
        json.dump(result, f)
    print("ref_runner ok:", args.out)


if __name__ == "__main__":
    main()
