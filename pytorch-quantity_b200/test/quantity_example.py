#!/usr/bin/env python
"""Calibrate a model and write feat.table / weight.table / weight + bias JSON -- the flow of the reference's
quantity/test/resnet18_quantity.py:16-52 and lenet_quantity.py:9-27 (merge_bn -> Quantity ->
activation_quantize -> weight_quantize -> rewrite_weight), on the GPU.

    python quantity_example.py --model lenet|resnet18_cifar|resnet18|resnet50 [--checkpoint x.pth]
                               [--images x.npy] [--batches 8] [--batch 16] [--workdir ./workdir]

Without --checkpoint / --images the weights are seeded random and the calibration images synthetic (this
environment has no datasets); the files written have the reference's formats either way."""
import argparse
import time

import _models
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet18_cifar", choices=sorted(_models.MODELS))
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--images", default=None, help=".npy of shape [N, C, H, W], already pre-processed")
    ap.add_argument("--batches", type=int, default=8)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--workdir", default="./workdir")
    args = ap.parse_args()
    from common.quantity import merge_bn
    from tools import Quantity
    net, shape = _models.build(args.model, args.checkpoint)
    data = _models.batches(shape, args.batches, args.batch, args.images)
    cfg, user = _models.configs(args.workdir, shape, len(data))
    with torch.no_grad():
        t0 = time.time()
        q = Quantity(merge_bn(net, "cpu"), config=cfg, user_config=user)
        q.activation_quantize(data)
        q.weight_quantize()
        q.rewrite_weight()          # the reference's scripts call it a second time (resnet18_quantity.py:52)
    print("calibrated %d images in %.2f s -> %s" % (sum(len(b[0]) for b in data), time.time() - t0, args.workdir))
    print(open(cfg["OUTPUT"]["FEAT_BIT_TABLE"]).read())


if __name__ == "__main__":
    main()
