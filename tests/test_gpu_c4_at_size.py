"""BASELINE config 4 AT ITS STATED SIZE: every one of the 23 unique ResNet-50 convolution geometries at batch 512
(M up to 6.4 M GEMM rows, 50 176 output tiles through the persistent scheduler, all TMEM accumulator stages in
flight), compared bit for bit with an evaluation that shares nothing with the kernels under test:

    q   = clamp(round(x * 2^ib))                                  torch elementwise     (new_quantity_op.py:48-58)
    acc = conv2d(q, Wq) in float64, rounded to int64              exact: |acc| < 2^53   (:124-126)
    y   = clamp(clamp(round_half_away(acc / 2^rs)) + bq) / 2^ob   torch int64 ops       (:11-44, :127-133)

Both of the kernel's epilogue flavours are checked per shape: the fp32 NCHW module boundary (reference semantics)
and the int8 NHWC payload with the fused ReLU that the int8 pipeline uses.
"""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (Cin, H, W, Cout, k, stride): SURVEY.md App. D / bench_conv_layers.py
R50 = [(3, 224, 224, 64, 7, 2), (64, 56, 56, 64, 1, 1), (64, 56, 56, 64, 3, 1), (64, 56, 56, 256, 1, 1),
       (256, 56, 56, 64, 1, 1), (256, 56, 56, 128, 1, 1), (128, 56, 56, 128, 3, 2), (128, 28, 28, 512, 1, 1),
       (256, 56, 56, 512, 1, 2), (512, 28, 28, 128, 1, 1), (128, 28, 28, 128, 3, 1), (512, 28, 28, 256, 1, 1),
       (256, 28, 28, 256, 3, 2), (256, 14, 14, 1024, 1, 1), (512, 28, 28, 1024, 1, 2), (1024, 14, 14, 256, 1, 1),
       (256, 14, 14, 256, 3, 1), (1024, 14, 14, 512, 1, 1), (512, 14, 14, 512, 3, 2), (512, 7, 7, 2048, 1, 1),
       (1024, 14, 14, 2048, 1, 2), (2048, 7, 7, 512, 1, 1), (512, 7, 7, 512, 3, 1)]
BATCH = 512


def exact_layer(x, conv_w, bias, info, stride, pad, chunk=64):
    """int64 result of the layer before de-quantisation, [B][K][P][Q]."""
    rs = info["weight_bit"] + info["input_bit"] - info["output_bit"]
    wq = torch.clamp(torch.round(conv_w * 2.0 ** info["weight_bit"]), -128, 127).double()
    bq = torch.clamp(torch.round(bias * 2.0 ** info["bias_bit"]), -128, 127).to(torch.int64)
    outs = []
    for i in range(0, x.shape[0], chunk):
        q = torch.clamp(torch.round(x[i:i + chunk] * 2.0 ** info["input_bit"]), -128, 127).double()
        acc = torch.round(F.conv2d(q, wq, None, stride, pad)).to(torch.int64)
        if rs > 0:
            mag = (acc.abs() + (1 << (rs - 1))) >> rs          # round half away from zero
            sh = torch.where(acc < 0, -mag, mag)
        else:
            sh = acc << (-rs)
        outs.append(torch.clamp(torch.clamp(sh, -128, 127) + bq.view(1, -1, 1, 1), -128, 127).to(torch.int16))
    return torch.cat(outs)


@pytest.mark.parametrize("shape", R50, ids=["%dx%dx%d_o%d_k%d_s%d" % s for s in R50])
def test_resnet50_conv_at_batch_512_vs_exact_evaluation(shape):
    import common.quantity as cq
    cin, h, w, cout, k, stride = shape
    pad = k // 2
    g = torch.Generator(device="cuda").manual_seed(1000 + cin + cout + k)
    x = torch.randn(BATCH, cin, h, w, device="cuda", generator=g) * 2.0
    info = {"weight_bit": 9, "input_bit": 5, "output_bit": 4, "bias_bit": 4}
    conv = nn.Conv2d(cin, cout, k, stride=stride, padding=pad).cuda()
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, device="cuda", generator=g) * 0.05)
        conv.bias.copy_(torch.randn(cout, device="cuda", generator=g) * 2.0)
        w_f, b_f = conv.weight.detach().clone(), conv.bias.detach().clone()
        m = cq.NewConv2d(conv, dict(info))
        want = exact_layer(x, w_f, b_f, info, stride, pad)
        got = m(x)                                               # fp32 NCHW (reference module boundary)
        assert torch.equal(got, want.float() / 16.0)
        del got
        m.int8_pipeline = True                                   # int8 NHWC payload, ReLU fused into the epilogue
        q8 = torch.relu(m(x)).int8_payload()
        assert q8.dtype == torch.int8 and q8.shape[0] == BATCH
        assert torch.equal(q8.permute(0, 3, 1, 2)[:, :cout].to(torch.int16), want.clamp_min(0))
    frac_sat = float(((want == 127) | (want == -128)).float().mean())
    assert frac_sat < 0.9, "degenerate test: everything saturates"
