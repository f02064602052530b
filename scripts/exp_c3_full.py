#!/usr/bin/env python
"""Experiment: where does pass 1 of the full-size BASELINE config 3 job (8192 images on ONE GPU, 24 of 128 batches
cached) lose time against the 20-step headline job (22.3 vs 17.9 ms per batch)?  Per-batch device time from CUDA
events recorded after every max-abs launch, caching-allocator statistics before / after, and the same job after the
allocator pool has been pre-sized with one block (what bench.py does for the headline job)."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402  (puts the package on sys.path)
import torch  # noqa: E402


def traced_job(net, workdir, presize):
    if presize:
        free, _ = torch.cuda.mem_get_info()
        pooled = torch.cuda.memory_reserved() - torch.cuda.memory_allocated()
        t0 = time.perf_counter()
        blk = torch.empty(int((free + pooled) * 0.66) - pooled, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        print("pre-size: one block of %.1f GB in %.3f s" % (blk.numel() / 1e9, time.perf_counter() - t0))
        del blk
    os.environ["PQ_BENCH_C3_TRACE"] = "1"
    out = bench.c3_full_job(net, workdir, 0, 1)
    print("c3_full: %s" % {k: out[k] for k in ("seconds", "images_per_s", "pass2_from_hbm_cache", "phases_s")})
    return out


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    net = bench.build_model()
    workdir = "/tmp/pq_exp_c3"
    warm = [bench.make_batch(10_000 + i, pin=True) for i in range(3)]
    bench.run_job(net, bench.ShardedBatches(warm, 3, 0, 1), 3, workdir, 0, 1)
    del warm
    for presize in ([False, True] if "--both" in sys.argv else [False]):
        torch.cuda.empty_cache()
        print("==== presize =", presize, flush=True)
        traced_job(net, workdir, presize)


if __name__ == "__main__":
    main()
