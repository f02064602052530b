#!/usr/bin/env python
"""Per-layer timing of the integer-simulation kernels on the 23 unique ResNet-50 conv geometries
 at batch B: quantise (fp32 NCHW -> int8 NHWC) and conv (int8 -> fp32 NCHW).
Prints a table; `--json` appends one JSON line.  Development tool, not the headline bench."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "pytorch-quantity_b200"))
import torch  # noqa: E402

from common.quantity import _native  # noqa: E402

# (Cin, H, W, Cout, k, stride, count)
R50 = [(3, 224, 224, 64, 7, 2, 1), (64, 56, 56, 64, 1, 1, 1), (64, 56, 56, 64, 3, 1, 3), (64, 56, 56, 256, 1, 1, 4),
       (256, 56, 56, 64, 1, 1, 2), (256, 56, 56, 128, 1, 1, 1), (128, 56, 56, 128, 3, 2, 1), (128, 28, 28, 512, 1, 1, 4),
       (256, 56, 56, 512, 1, 2, 1), (512, 28, 28, 128, 1, 1, 3), (128, 28, 28, 128, 3, 1, 3), (512, 28, 28, 256, 1, 1, 1),
       (256, 28, 28, 256, 3, 2, 1), (256, 14, 14, 1024, 1, 1, 6), (512, 28, 28, 1024, 1, 2, 1), (1024, 14, 14, 256, 1, 1, 5),
       (256, 14, 14, 256, 3, 1, 5), (1024, 14, 14, 512, 1, 1, 1), (512, 14, 14, 512, 3, 2, 1), (512, 7, 7, 2048, 1, 1, 3),
       (1024, 14, 14, 2048, 1, 2, 1), (2048, 7, 7, 512, 1, 1, 2), (512, 7, 7, 512, 3, 1, 2)]


def time_ms(fn, iters=5):
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--s8-out", action="store_true", help="store int8 NHWC instead of fp32 NCHW")
    ap.add_argument("--only", type=int, default=None, help="run only this row of the table (for ncu captures)")
    ap.add_argument("--fused-add", action="store_true", help="conv + NewAdd + ReLU in one kernel (int16 shortcut)")
    ap.add_argument("--classic-bias", action="store_true", help="plain int32 bias (no PQ_FLAG_BIAS_FOLDED constants)")
    ap.add_argument("--no-windows", action="store_true", help="3x3 layers through the im2col-TMA path (PQ_FLAG_NO_WINDOWS)")
    ap.add_argument("--no-s2d", action="store_true", help="the stem on 8-byte pixels instead of the space-to-depth form")
    args = ap.parse_args()
    B = args.batch
    tot_conv = tot_q = tot_ops = 0.0
    print("%-28s %9s %8s %8s %8s %8s %8s" % ("layer (Cin,H,W,Cout,k,s)xN", "GOP", "conv_ms", "TOPS", "outGB/s", "quant_ms", "qGB/s"))
    for (cin, h, w, cout, k, s, cnt) in (R50 if args.only is None else [R50[args.only]]):
        pad = k // 2
        plain = k == 1 and s == 1
        cpad = (cin + 15) // 16 * 16 if plain else (cin + 31) // 32 * 32
        x = torch.randn(B, cin, h, w, device="cuda")
        wk = torch.randint(-128, 127, (cout, k, k, cpad), dtype=torch.int8, device="cuda")
        wk[..., cin:] = 0
        bias = torch.randint(-128, 127, (cout,), dtype=torch.int32, device="cuda")
        if not args.classic_bias:
            bias = _native.bias_fold(bias, 9)          # what NewConv2d / NewLinear hold (rs = 9 below)
        P = (h + 2 * pad - k) // s + 1
        if cin <= 8 and not plain and s % 2 == 0:      # windowed small-channel convolution (the stem)
            w8 = torch.zeros((cout, k, 8, 8), dtype=torch.int8, device="cuda")
            w8[:, :, :k, :cin] = torch.randint(-128, 127, (cout, k, k, cin), dtype=torch.int8, device="cuda")
            w8 = w8.view(cout, k, 64)
            Hp = max((P - 1) * s + k, h + pad); Hp = (Hp + s - 1) // s * s
            Wp = max((P - 1) * s + 8, w + pad); Wp += Wp & 1
            if s == 2 and cin <= 4 and k <= 7 and not args.no_s2d:      # space-to-depth form (what NewConv2d runs)
                e = pad & 1
                ra = ((e + k - 1) >> 1) + 1
                w2 = torch.zeros((cout, ra, 4, 2, 2, 4), dtype=torch.int8, device="cuda")
                for r in range(k):
                    a, dy = divmod(e + r, 2)
                    for t in range(k):
                        b, dx = divmod(e + t, 2)
                        w2[:, a, b, dy, dx, :cin] = torch.randint(-128, 127, (cout, cin), dtype=torch.int8, device="cuda")
                w2 = w2.view(cout, ra, 64)
                hp2, wp2 = P + ra - 1, P + 3
                q = _native.quantize_s2d16_s8(x, 4, (pad + e, pad + e), hp2, wp2)
                t_q = time_ms(lambda: _native.quantize_s2d16_s8(x, 4, (pad + e, pad + e), hp2, wp2))
                q8 = q.view(B, hp2, 2 * wp2, 8)
                t_c = time_ms(lambda: _native.conv2d_smallc_s8(q8, w2, bias, (hp2, 2 * wp2), (ra, 8), (1, 2), (0, 0), 9, 4,
                                                               want_f32=not args.s8_out, want_s8=args.s8_out))
            else:
                q = _native.quantize_pad_nhwc8_s8(x, 4, (pad, pad), Hp, Wp)
                t_q = time_ms(lambda: _native.quantize_pad_nhwc8_s8(x, 4, (pad, pad), Hp, Wp))
                t_c = time_ms(lambda: _native.conv2d_smallc_s8(q, w8, bias, (h, w), (k, k), (s, s), (pad, pad), 9, 4,
                                                               want_f32=not args.s8_out, want_s8=args.s8_out))
        elif cin <= 8 and not plain:     # explicit im2col + GEMM
            kp = (k * k * cin + 63) // 64 * 64
            wn = torch.randint(-128, 127, (cout, kp), dtype=torch.int8, device="cuda")
            wn[:, k * k * cin:] = 0
            q, _ = _native.quantize_im2col_s8(x, 4, (k, k), (s, s), (pad, pad), kp)
            t_q = time_ms(lambda: _native.quantize_im2col_s8(x, 4, (k, k), (s, s), (pad, pad), kp))
            t_c = time_ms(lambda: _native.gemm_s8(q, wn, bias, 9, 4, hw=P * P, want_f32=not args.s8_out,
                                                  want_s8=args.s8_out))
        elif args.fused_add:
            q = _native.quantize_nchw_to_nhwc_s8(x, 4, cpad)
            t_q = 0.0
            sc = torch.randint(-2000, 2000, (B, P, P, cout), dtype=torch.int16, device="cuda")
            # the common case in feat.table: conv output, shortcut and Eltwise all at the same bit
            t_c = time_ms(lambda: _native.conv2d_s8_add(q, wk, bias, (s, s), (pad, pad), 9, 4, sc, 4, False, 4, True))
        else:
            q = _native.quantize_nchw_to_nhwc_s8(x, 4, cpad)
            t_q = time_ms(lambda: _native.quantize_nchw_to_nhwc_s8(x, 4, cpad))
            t_c = time_ms(lambda: _native.conv2d_s8(q, wk, bias, (s, s), (pad, pad), 9, 4, want_f32=not args.s8_out,
                                                    want_s8=args.s8_out, windows=not args.no_windows))
        ops = 2.0 * B * P * P * cout * k * k * cin
        out_bytes = B * P * P * cout * (1 if args.s8_out else 4)
        if args.fused_add and t_q == 0.0:             # all bytes of the fused kernel: A, shortcut int16, int16 + int8 out
            out_bytes = q.numel() + B * P * P * cout * 5
        qbytes = x.numel() * 4 + q.numel()
        print("%-28s %9.1f %8.3f %8.1f %8.0f %8.3f %8.0f" % ("(%d,%d,%d,%d,%d,%d)x%d" % (cin, h, w, cout, k, s, cnt), ops / 1e9,
              t_c, ops / t_c / 1e9, out_bytes / t_c / 1e6, t_q, qbytes / max(t_q, 1e-9) / 1e6))
        tot_conv += t_c * cnt; tot_q += t_q * cnt; tot_ops += ops * cnt
        del x, q, wk
    print("TOTAL conv %.2f ms (%.1f TOPS), quantise %.2f ms, %.2f TOP" % (tot_conv, tot_ops / tot_conv / 1e9, tot_q, tot_ops / 1e12))
    print(json.dumps({"batch": B, "conv_ms": round(tot_conv, 3), "quant_ms": round(tot_q, 3),
                      "TOPS": round(tot_ops / tot_conv / 1e9, 1), "s8_out": args.s8_out}))


if __name__ == "__main__":
    main()
