"""Development experiment: fp32 forward of the calibration model under layout / graph / TF32 variants
(NCHW 17.2 ms, channels_last 21.1 ms, CUDA graph 16.7 ms, TF32 6.6 ms on B200, batch 64)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-quantity_b200"))
import torch
from bench import build_model
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
net = build_model().cuda().eval()
x = torch.randn(64, 3, 224, 224, device="cuda")
def t(fn, n=10):
    for _ in range(5): fn()
    torch.cuda.synchronize(); s = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - s) / n * 1e3
with torch.no_grad():
    print("NCHW fp32 fwd ms:", t(lambda: net(x)))
    net_cl = net.to(memory_format=torch.channels_last); xcl = x.contiguous(memory_format=torch.channels_last)
    print("channels_last fp32 fwd ms:", t(lambda: net_cl(xcl)))
    g = torch.cuda.CUDAGraph()
    net = net.to(memory_format=torch.contiguous_format)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): net(x)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        y = net(x)
    print("NCHW fp32 fwd CUDA graph ms:", t(lambda: g.replay()))
    torch.backends.cudnn.allow_tf32 = True
    print("NCHW tf32 fwd ms (for reference only):", t(lambda: net(x)))
