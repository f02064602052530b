#!/usr/bin/env python
"""Rebuild the calibrated model as the integer simulation (ReconModel) and the fake-quant model (ReconTest) and
compare them with the fp32 model -- the flow of the reference's quantity/test/resnet_reconstruction.py:17-150
and lenet_reconstruction.py.  Run quantity_example.py with the same --model / --workdir first.

    python reconstruction_example.py --model resnet18_cifar [--checkpoint x.pth] [--images x.npy --labels y.npy]
                                     [--int8-pipeline] [--workdir ./workdir]

With labels it prints top-1 accuracy of every variant (as the reference's script does on CIFAR-10); without,
top-1 agreement with the fp32 model and the largest logit difference."""
import argparse

import _models
import torch


def evaluate(name, model, data, ref_pred=None):
    correct = agree = total = 0
    preds = []
    with torch.no_grad():
        for i, (x, y) in enumerate(data):
            out = model(x.cuda())
            p = out.argmax(1).cpu()
            preds.append(p)
            total += len(p)
            if y is not None:
                correct += int((p == y).sum())
            if ref_pred is not None:
                agree += int((p == ref_pred[i]).sum())
    msg = "%-28s" % name
    if data[0][1] is not None:
        msg += " acc %.3f" % (100.0 * correct / total)
    if ref_pred is not None:
        msg += " top-1 agreement with fp32 %.3f" % (100.0 * agree / total)
    print(msg)
    return preds


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet18_cifar", choices=sorted(_models.MODELS))
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--images", default=None)
    ap.add_argument("--labels", default=None)
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--workdir", default="./workdir")
    ap.add_argument("--int8-pipeline", action="store_true", help="keep activations int8 between layers (bit-identical)")
    args = ap.parse_args()
    from common.quantity import QuanDequan, enable_int8_pipeline
    from tools import Reconstruction
    net, shape = _models.build(args.model, args.checkpoint)
    data = _models.batches(shape, args.batches, args.batch, args.images, args.labels, seed=1000)
    cfg, user = _models.configs(args.workdir, shape, 1)
    with torch.no_grad():
        ref_pred = evaluate("origin model", net.cuda(), data)
        recon = Reconstruction(net, config=cfg)
        merged = recon.merge_bn().eval()
        evaluate("merge bn model", merged.cuda(), data, ref_pred)
        info = recon.get_quantity_information()
        model = recon.ReconModel(info, user["PATH"]["QUANTITY_MODEL_PATH"]).cuda().eval()
        if args.int8_pipeline:
            enable_int8_pipeline(model)
        evaluate("reconstruction model", model, data, ref_pred)
        # ReconTest needs a fresh copy: ReconModel replaced the layers of `net` in place (reconstruction.py:183-238)
        net2, _ = _models.build(args.model, args.checkpoint)
        recon2 = Reconstruction(net2, config=cfg)
        recon2.merge_bn()
        info2 = recon2.get_quantity_information()
        test_model = recon2.ReconTest(info2, None).cuda().eval()
        image_q = QuanDequan(8, info2["image"]["output_bit"])      # the caller fake-quantises the image itself
        evaluate("q-dq reconstruction model", test_model, [(image_q(x.cuda()).cpu(), y) for x, y in data], ref_pred)


if __name__ == "__main__":
    main()
