"""Diagnostic (gpurun): where does this repo's ReconModel first differ from the reference's on the GPU, and is the
reference's own fp32 convolution exact on integer operands there?  Usage: python scripts/diag_reconmodel.py r18 8"""
import json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "pytorch-quantity_b200"), REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import ref_models
from test_gpu_vs_reference import _configs, _snapshot

name, B = sys.argv[1], int(sys.argv[2])
import common.quantity as cq, tools
ref_models.set_deterministic()
wd = "/tmp/diag_%s/workdir" % name
os.makedirs(wd, exist_ok=True)
cfg, user = _configs(wd, ref_models.INPUT_SHAPE[name], 1)
with torch.no_grad():
    q = tools.Quantity(cq.merge_bn(ref_models.build_model(name), "cpu"), config=cfg, user_config=user, verbose=False)
    q.activation_quantize(ref_models.calib_batches(name, 2, 4)); q.weight_quantize(); q.rewrite_weight()
for tag, extra in (("cudnn", []), ("nocudnn", ["--no-cudnn"])):
    out = "/tmp/diag_%s/ref_%s" % (name, tag)
    p = subprocess.run([sys.executable, os.path.join(REPO, "baseline", "ref_runner.py"), "--model", name, "--device", "gpu",
                        "--out", out, "--tables-from", wd, "--recon", "ReconModel", "--eval", str(B), "--dump-layers",
                        "--self-check"] + extra, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout[-3000:]
    res = json.load(open(out + "/result.json"))
    r = tools.Reconstruction(ref_models.build_model(name), config=cfg); r.merge_bn()
    model = r.ReconModel(r.get_quantity_information(), None).cuda().eval()
    outs = {}
    for ln, m in model.named_modules():
        if type(m).__name__ in ("NewConv2d", "NewLinear", "NewAdd"):
            m.register_forward_hook(lambda m, i, o, ln=ln: outs.__setitem__(ln, o.detach().clone()))
    with torch.no_grad():
        model(ref_models.eval_batch(name, B).cuda())
    print("==", name, tag)
    for ln in outs:                                    # forward order
        want = torch.from_numpy(np.load(os.path.join(out, "layers", "ReconModel", ln + ".npy"))).cuda()
        sc = res["ReconModel/self_check"].get(ln, {})
        print("%-28s mismatch %.6f  ref_self_inexact %.6f max_err %.4g max|acc| %.4g" % (
            ln, float((outs[ln] != want).float().mean()), sc.get("inexact_fraction", -1), sc.get("max_abs_err", -1),
            sc.get("max_abs_acc", -1)))
