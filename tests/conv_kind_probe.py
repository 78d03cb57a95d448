"""GPU probe: accuracy and per-layer time of the two conv kinds (3xTF32 vs 3xFP16) on the layer shapes of the headline workload.
    timeout 300 python tests/conv_kind_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from detectorch_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def ref64(x, w, k, pad, stride):
    y = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(0, 3, 1, 2), stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()


def acc():
    for (N, H, W, Cin, Cout, k, pad, stride) in [(1, 1, 128, 64, 64, 1, 0, 1), (1, 20, 30, 64, 128, 3, 1, 1), (2, 25, 38, 256, 256, 3, 1, 1),
                                                  (1, 1, 512, 2048, 256, 1, 0, 1), (5, 14, 14, 256, 256, 3, 1, 1)]:
        g = torch.Generator().manual_seed(Cin + Cout)
        x = torch.randn((N, H, W, Cin), generator=g)
        w = torch.randn((Cout, k, k, Cin), generator=g) * (2.0 / (k * k * Cin)) ** 0.5
        y = ref64(x, w, k, pad, stride)
        one, zero = torch.ones(Cout, device=dev), torch.zeros(Cout, device=dev)
        out = {}
        for kind in ("tf32", "f16"):
            got = ops.conv2d_nhwc(x.to(dev), w.reshape(Cout, -1).contiguous().to(dev), one, zero, k, k, pad, stride, kind=kind)
            torch.cuda.synchronize()
            out[kind] = float((got.cpu().double() - y).abs().max() / y.abs().max())
        print("ACC N%d %dx%d Cin%d Cout%d k%d  tf32 %.2e  f16 %.2e" % (N, H, W, Cin, Cout, k, out["tf32"], out["f16"]), flush=True)


def timing(only=None, with_res=False):
    shapes = [("P2 3x3 256->256 200x304", 8, 200, 304, 256, 256, 3, 1, 1, 0), ("C4 3x3 256->256 50x76", 8, 50, 76, 256, 256, 3, 1, 1, 0),
              ("C4 1x1 1024->256", 8, 50, 76, 1024, 256, 1, 0, 1, 0), ("C4 1x1 256->1024", 8, 50, 76, 256, 1024, 1, 0, 1, 0),
              ("C2 3x3 64->64 200x304", 8, 200, 304, 64, 64, 3, 1, 1, 0), ("C2 1x1 64->256", 8, 200, 304, 64, 256, 1, 0, 1, 0),
              ("C3 3x3 128->128 100x152", 8, 100, 152, 128, 128, 3, 1, 1, 0), ("mask 3x3 256->256 precise", 800, 14, 14, 256, 256, 3, 1, 1, -1),
              ("mask 3x3 256->256 128 plain", 800, 14, 14, 256, 256, 3, 1, 1, 128), ("FC6 12544->1024", 1, 1, 8000, 12544, 1024, 1, 0, 1, 0)]
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
    if only is not None:
        shapes = [shapes[i] for i in only]
    for (name, N, H, W, Cin, Cout, k, pad, stride, fbn) in shapes:
        x = torch.randn((N, H, W, Cin), device=dev)
        w = torch.randn((Cout, k * k * Cin), device=dev) * 0.02
        one, zero = torch.ones(Cout, device=dev), torch.zeros(Cout, device=dev)
        wl = ops.tf32_residual(w)
        mult = ops.weight_multiplier(w)
        hi, lo = ops.fp16_split(w, mult)
        sc16 = (one / mult).contiguous()
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        out = torch.empty((N, Ho, Wo, Cout), device=dev)
        res = torch.randn((N, Ho, Wo, Cout), device=dev) if with_res else None
        rp, rm = (ops._p(res), 1) if with_res else (None, 0)
        L = ops._lib.lib()
        fl = 2.0 * N * Ho * Wo * Cout * Cin * k * k
        tms = {}
        for kind in ("tf32", "f16"):
            def run():
                if kind == "tf32":
                    L.dt_conv2d_nhwc(ops._p(x), N, H, W, Cin, Cin, ops._p(w), ops._p(wl), Cout, k, k, pad, stride, ops._p(one), ops._p(zero), rp, rm,
                                     None, 0, 0, 1, 0, 3, fbn, ops._p(out), Cout, ops._stream())
                else:
                    L.dt_conv2d_nhwc_f16x3(ops._p(x), N, H, W, Cin, Cin, ops._p(hi), ops._p(lo), Cout, k, k, pad, stride, ops._p(sc16), ops._p(zero),
                                           rp, rm, None, 0, 0, 1, 0, 3, fbn, None, ops._p(out), Cout, ops._stream())
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            tms[kind] = sorted(ts)[2]
        print("TIME %-30s tf32 %8.1f us (%6.1f TF/s alg)   f16 %8.1f us (%6.1f TF/s alg)   x%.2f" %
              (name, tms["tf32"] * 1e3, fl / tms["tf32"] / 1e9, tms["f16"] * 1e3, fl / tms["f16"] / 1e9, tms["tf32"] / tms["f16"]), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["acc", "timing"]
    if "acc" in which:
        acc()
    if "timing" in which:
        timing()
    if "one" in which:
        timing(only=[1, 2])
    if "mask" in which:
        timing(only=[7])
    if "res" in which:
        timing(only=[3, 5], with_res=True)
