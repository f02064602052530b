#!/usr/bin/env python
"""bench.py -- headline benchmark of the calibration hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric "calib images/sec (hist+KL)"
is quoted on): ResNet-50 (fabu-style, BN folded, random init) calibration on synthetic
3x224x224 images, micro-batch 64 per rank.  One STEP = one micro-batch taken through the whole
hot path: fp32 forward with device-resident hooks, pass-1 max-abs of the 71 observed tensors,
pass-2 2048-bin histograms; the job ends with the NCCL MAX / SUM merge, the KL threshold search
of all 71 tensors and the feat.table bits.  The timed region is the whole job for K steps per
rank; value = N*K*64 images / max-over-ranks device time.  Weak scaling: per-rank work fixed.

  value : inputs resident in HBM before the timed region (tools.Quantity.activation_quantize
          on device tensors).
  e2e   : the same public call on HOST batches in pinned memory; the H2D copy of every batch and
          the D2H read of maxima / histograms / thresholds are inside the timed region.
  roofline : the dominant kernel of this repo (pq_hist2048_multi_f32, HBM-bound, 4 algorithmic
          bytes per element per launch), timed live with CUDA events on its launching stream.
  cpu_baseline / --impl reference : the reference's CPU algorithm (the oracle port: torch CPU
          forward + oracle/ statistics and KL, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "pytorch-quantity_b200")
for p in (PKG, REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

MICRO_BATCH = 64
IMG_SHAPE = (3, 224, 224)
METRIC = "calib images/sec (hist+KL)"
UNIT = "images/s"
WORKLOAD = "ResNet-50 224x224 calibration (max-abs + 2048-bin hist + KL search), micro-batch 64 per GPU"


def build_model():
    from model.resnet.resnet_fabu import randomize_bn_, resnet50_fabu
    from common.quantity import merge_bn
    torch.manual_seed(0)
    net = resnet50_fabu().eval()
    with torch.no_grad():
        randomize_bn_(net, 0)
        merge_bn(net, "cpu")
    return net


def make_batch(index, batch=MICRO_BATCH, device="cpu", pin=False):
    g = torch.Generator().manual_seed(1 + index)
    x = torch.randn(batch, *IMG_SHAPE, generator=g)
    if device != "cpu":
        return x.to(device)
    return x.pin_memory() if pin else x


def tool_configs(workdir, n_batches):
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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc, self.first = gpu_index, [], None, 0

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self, wait_s=3.0):
        """Call right before the timed region: waits until nvidia-smi has delivered its first sample (its NVML
        start-up takes driver locks for ~50 ms and must not land inside the timed region) and discards what
        was sampled so far."""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < wait_s:
            time.sleep(0.02)
        self.first = len(self.rows)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows[self.first:] or self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        loaded = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full
    capture of this same bench command (profiles/r01_traffic.json, written by profiles/summarize.py)."""
    try:
        with open(os.path.join(REPO, "profiles", "r01_traffic.json")) as f:
            return float(json.load(f)[kernel]["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def fakequant_bandwidth(n_elems=1 << 28, iters=10):
    """Second half of BASELINE.json's metric: QuanDequan (pq_fakequant_f32) GB/s on a 1 GiB fp32 tensor
    (larger than L2), 8 algorithmic bytes per element, CUDA events on the launching stream."""
    from common.quantity import _native
    x = torch.randn(n_elems, device="cuda")
    for _ in range(3):
        y = _native.fakequant(x, 4)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        y = _native.fakequant(x, 4)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    del x, y
    return 8.0 * n_elems / (ms * 1e-3) / 1e9, ms


# ----------------------------------------------------------------------------- our arm
def run_job(net, batches, n_global, workdir, rank, world):
    """One whole calibration job through the public API; returns (seconds on device, Quantity)."""
    import tools
    cfg, user = tool_configs(workdir, n_global)
    q = tools.Quantity(net, config=cfg, user_config=user, verbose=False)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    q.activation_quantize(batches)
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    ms = start.elapsed_time(end)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return ms / 1e3, q


class ShardedBatches:
    """A global list of n_global batches of which only this rank's (index % world == rank) exist."""

    def __init__(self, mine, n_global, rank, world):
        self.mine, self.n, self.rank, self.world = mine, n_global, rank, world

    def __iter__(self):
        for i in range(self.n):
            yield (self.mine[i // self.world], None) if i % self.world == self.rank else None


def ours(args):
    from common.quantity import _native
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.backends.cudnn.allow_tf32 = False           # the forward stays true fp32, like the reference
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = os.environ.get("PQ_BENCH_NO_AUTOTUNE") != "1"   # off under ncu
    _native.lib()
    net = build_model()
    workdir = os.path.join("/tmp", "pq_bench_rank%d" % rank)
    K, W = args.steps, max(args.warmup, 3)

    # warm-up: W steps through the same job shape (cuDNN autotune, allocator, kernels)
    # (pinned host batches: the warm-up also exercises the copy stream / H2D path the e2e job uses)
    warm = [make_batch(10_000 + rank * W + i, pin=True) for i in range(W)]
    _, qw = run_job(net, ShardedBatches(warm, W * world, rank, world), W * world, workdir, rank, world)
    del warm
    # Pre-size the caching allocator's pool for the K-batch activation cache so that no cudaMalloc
    # lands inside the timed region: allocate one block of the expected size and release it to the pool.
    per_batch = qw.timings["cached_bytes"] // max(1, qw.timings["cached_batches"])
    free, _total = torch.cuda.mem_get_info()
    want = min(int((K + 2) * per_batch * 1.05), int(free * 0.8))
    if want > 0:
        torch.empty(want, dtype=torch.uint8, device="cuda")

    # ---- value: inputs resident in HBM ---------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                    # started early: see ClockSampler.mark
    dev_batches = [make_batch(rank + world * i, device="cuda") for i in range(K)]
    prof = {}
    _native.set_profile(prof)
    l0 = _native.LAUNCHES["total"]
    if rank == 0:
        sampler.mark()
    secs, q = run_job(net, ShardedBatches(dev_batches, K * world, rank, world), K * world, workdir, rank, world)
    clocks = sampler.stop() if rank == 0 else None
    launches = _native.LAUNCHES["total"] - l0
    _native.set_profile(None)
    torch.cuda.synchronize()
    value = world * K * MICRO_BATCH / secs
    timings = dict(q.timings)
    bits_image = q.last_calibration["bits"]["image"]
    del dev_batches

    def kernel_stats(name):
        rows = prof.get(name, [])
        if not rows:
            return None
        ms = [s.elapsed_time(e) for s, e, _ in rows]
        nbytes = [b for _, _, b in rows]
        return {"launches": len(rows), "avg_ms": sum(ms) / len(ms), "avg_bytes": sum(nbytes) / len(nbytes),
                "total_ms": sum(ms)}

    hist, amax, kl = kernel_stats("hist"), kernel_stats("absmax"), kernel_stats("kl")
    peak, peak_src = peaks()
    achieved = hist["avg_bytes"] / (hist["avg_ms"] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "pq::hist_multi_kernel (pq_hist2048_multi_f32)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": ncu_traffic("hist_multi"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": hist["avg_bytes"], "avg_launch_ms": round(hist["avg_ms"], 4),
                "absmax_GBps": round(amax["avg_bytes"] / (amax["avg_ms"] * 1e-3) / 1e9, 1),
                "kl_ms_per_job": round(kl["total_ms"], 3), "kl_us_per_tensor": round(kl["total_ms"] * 1e3 / 71, 1),
                "stats_share_of_step": round((hist["total_ms"] + amax["total_ms"]) / (secs * 1e3), 4)}

    # ---- e2e: host batches in pinned memory through the same public call ---------------------
    host_batches = [make_batch(rank + world * i, pin=True) for i in range(K)]
    e_secs, q2 = run_job(net, ShardedBatches(host_batches, K * world, rank, world), K * world, workdir, rank, world)
    n_fwd_pass2 = K - q2.timings.get("cached_batches", 0)
    h2d = (K + n_fwd_pass2) * MICRO_BATCH * int(np.prod(IMG_SHAPE)) * 4 / K
    d2h = (71 * 4 + 71 * 2048 * 8 + 71 * 4) / K
    e2e = {"value": round(world * K * MICRO_BATCH / e_secs, 2), "unit": UNIT,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "phases_s": {k: round(v, 4) for k, v in q2.timings.items() if k.endswith("_s")}}
    assert q2.last_calibration["bits"]["image"] == bits_image
    del host_batches

    fq_gbps, fq_ms = fakequant_bandwidth()
    line = None
    if rank == 0:
        cpu = cpu_baseline(sample_images=args.cpu_sample) if world == 1 and not args.no_cpu_baseline else None
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(secs * 1e3 / K, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": world * MICRO_BATCH, "images": world * K * MICRO_BATCH,
                           "observed_tensors": 71, "elements_per_image": 16784872,
                           "l2": "inputs larger than L2 (4.3 GB of activations per step)",
                           "pass2": "%d of %d steps served from the HBM activation cache"
                                    % (timings.get("cached_batches", 0), K),
                           "parallelism": "dp%d" % world},
                "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "fakequant": {"GBps": round(fq_gbps, 1), "frac_of_hbm_peak": round(fq_gbps / peak, 4),
                              "elements": 1 << 28, "ms": round(fq_ms, 4), "algorithmic_bytes_per_element": 8},
                "phases_s": {k: round(v, 4) for k, v in timings.items() if k.endswith("_s")}}
        if cpu:
            line["cpu_baseline"] = cpu
    if world > 1:
        torch.distributed.destroy_process_group()
    return line


# ------------------------------------------------------------------- CPU baseline / reference arm
def cpu_calibration_sample(n_images, batch, threads):
    """The reference's CPU algorithm on `n_images` synthetic images: torch CPU fp32 forward with
    hooks (both passes, like pytorch_quantizer.py:379-426), oracle max-abs / histograms per tensor
    fanned out over `threads` host threads (the reference fans out over a process pool), then the
    oracle KL search for all tensors.  Returns seconds."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pq_oracle as oracle
    oracle.build()
    torch.set_num_threads(threads)
    net = build_model()
    feats = {}
    order = []

    def hook(name):
        def _h(m, i, o):
            if not feats:
                feats["image"] = i[0].detach().numpy().reshape(-1)
            feats[name] = o.detach().numpy().reshape(-1)
        return _h

    for name, m in net.named_modules():
        if type(m).__name__ in ("Conv2d", "Linear", "Eltwise"):
            m.register_forward_hook(hook(name))
            order.append(name)
    names = ["image"] + order
    batches = [make_batch(i, batch) for i in range(max(1, n_images // batch))]
    pool = ThreadPoolExecutor(threads)
    t0 = time.perf_counter()
    maxes = {n: 0 for n in names}
    with torch.no_grad():
        for x in batches:                                           # pass 1
            feats.clear()
            net(x)
            for n, m in zip(names, pool.map(lambda n: oracle.absmax_update(maxes[n], feats[n]), names)):
                maxes[n] = m
        intervals = {n: oracle.interval(maxes[n]) for n in names}
        hists = {n: np.zeros(2048, np.int64) for n in names}
        for x in batches:                                           # pass 2
            feats.clear()
            net(x)
            for n, h in zip(names, pool.map(lambda n: oracle.hist(feats[n], intervals[n]), names)):
                hists[n] += h
    bits = list(pool.map(lambda n: oracle.quantize_distribution(hists[n], intervals[n])[0], names))
    secs = time.perf_counter() - t0
    assert len(bits) == 71
    return secs, len(batches) * batch


def cpu_baseline(sample_images=16):
    threads = os.cpu_count() or 1
    secs, n = cpu_calibration_sample(sample_images, min(8, sample_images), threads)
    return {"value": round(n / secs, 3), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d images (2 batches of 8) through torch-CPU fp32 forward x2 + oracle max-abs/hist "
                      "+ oracle KL search of 71 tensors; the reference itself is pure Python and is slower "
                      "(C1 golden run: 1.6 images/s on 8 cores)" % n}


def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return None
    threads = os.cpu_count() or 1
    K, W = args.steps, max(args.warmup, 1)
    per_step = 8                                       # bounded sample: 8 images per step
    for _ in range(min(W, 1)):
        cpu_calibration_sample(per_step, per_step, threads)
    k_eff = min(K, 6)                                   # keep the whole run within a few minutes
    secs, n = cpu_calibration_sample(per_step * k_eff, per_step, threads)
    value = round(n / secs, 3)
    return ({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": k_eff, "warmup": W,
        "ms_per_step": round(secs * 1e3 / k_eff, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_images_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d steps of %d images; torch-CPU forward x2 + oracle statistics + KL"
                                   % (k_eff, per_step)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # the drop-in classes print progress like the reference does; keep stdout to the ONE JSON line
    real_stdout = sys.stdout
    sys.stdout = sys.stderr
    try:
        line = reference_arm(args) if args.impl == "reference" else ours(args)
    finally:
        sys.stdout = real_stdout
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
