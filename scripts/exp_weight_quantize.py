import sys, os, time, tempfile
sys.path.insert(0, "/root/repo"); import bench, torch
import ref_models, tools
from common.quantity import merge_bn
workdir = tempfile.mkdtemp(prefix="pq_wq_")
cfg, user = bench.tool_configs(workdir, 8)
batches = ref_models.calib_batches("r18", 8, 8)
with torch.no_grad():
    net = merge_bn(ref_models.build_model("r18"), "cpu")
    q = tools.Quantity(net, config=cfg, user_config=user, verbose=False)
    q.activation_quantize(batches)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    q.weight_quantize()
    print("weight_quantize (tables + JSON + rewrite) %.2f s" % (time.perf_counter() - t0))
