"""Isolated timings of the fp16-engine layers of the ResNet-50 trunk on the bench workload's 8-image set
(7-scale pyramid + target at 480x640): CUDA events, L2 flushed between launches.

    python scripts/f16_layer_bench.py            # table: us, GB/s of algorithmic bytes, TFLOP/s
    python scripts/f16_layer_bench.py --once     # one launch per layer (for ncu -k regex:...)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ransac_flow_b200 as rf  # noqa: E402
from ransac_flow_b200.program import LayerProgram  # noqa: E402

ONCE = "--once" in sys.argv
SIZES = [(960, 1280), (800, 1056), (640, 848), (480, 640), (400, 528), (320, 416), (240, 320), (480, 640)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=12):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return 0.0
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


def hw(div):
    return [(h // div, w // div) for h, w in SIZES]


def conv_case(name, cin, cout, k, stride, div_in, res):
    g = torch.Generator(device="cuda").manual_seed(0)
    ihw = hw(div_in)
    pin = sum(h * w for h, w in ihw)
    x = rf.ops.Ragged(torch.randn(pin, cin, device="cuda", generator=g).half(), ihw)
    w = (torch.randn(cout, k * k * cin, device="cuda", generator=g) / np.sqrt(k * k * cin)).half()
    bias = torch.randn(cout, device="cuda", generator=g)
    ohw = hw(div_in * stride)
    pout = sum(h * w for h, w in ohw)
    r = rf.ops.Ragged(torch.randn(pout, cout, device="cuda", generator=g).half(), ohw) if res else None
    us = timed(lambda: rf.ops.conv2d(x, None, bias, cout, k, stride, k // 2, True, r, rf.ops.ENGINE_F16, w))
    by = (pin * cin + pout * cout * (2 if res else 1) + cout * cin * k * k) * 2
    fl = 2.0 * pout * cin * cout * k * k
    if not ONCE:
        print("%-28s %8.1f us  %7.1f MB  %6.0f GB/s  %6.1f TFLOP/s" % (name, us, by / 1e6, by / us / 1e3, fl / us / 1e6))


def stem_case():
    g = torch.Generator(device="cuda").manual_seed(0)
    pin = sum(h * w for h, w in SIZES)
    x = rf.ops.Ragged(torch.randn(pin, 3, device="cuda", generator=g), SIZES)

    class BN:
        weight, bias = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
        running_mean, running_var, eps = torch.zeros(64, device="cuda"), torch.ones(64, device="cuda"), 1e-5
    wt = torch.randn(64, 3, 7, 7, device="cuda", generator=g) / 12
    for fused in (True, False):
        P = LayerProgram(3)
        if fused:
            P.stem7_fused(0, wt, BN)
        else:
            P.stem(0, wt, BN, 2, 3, 64)
        us = timed(lambda: P.run(x, 2))
        pout = pin // 4
        by = pin * 3 * 4 + pout * 64 * 2
        if not ONCE:
            print("%-28s %8.1f us  %7.1f MB  %6.0f GB/s" % ("stem 7x7/2 " + ("fused" if fused else "im2col + 1x1"), us, by / 1e6, by / us / 1e3))


stem_case()
conv_case("l1.c1 256->64", 256, 64, 1, 1, 4, False)
conv_case("l1.c2 64->64 3x3", 64, 64, 3, 1, 4, False)
conv_case("l1.c3 64->256 +res", 64, 256, 1, 1, 4, True)
conv_case("l1.ds 64->256", 64, 256, 1, 1, 4, False)
conv_case("l2.c2 128->128 3x3 s2", 128, 128, 3, 2, 4, False)
conv_case("l2.c3 128->512 +res", 128, 512, 1, 1, 8, True)
conv_case("l2.c1 512->128", 512, 128, 1, 1, 8, False)
conv_case("l2.c2 128->128 3x3", 128, 128, 3, 1, 8, False)
conv_case("l3.ds 512->1024 s2", 512, 1024, 1, 2, 8, False)
conv_case("l3.c1 1024->256", 1024, 256, 1, 1, 16, False)
conv_case("l3.c2 256->256 3x3", 256, 256, 3, 1, 16, False)
conv_case("l3.c3 256->1024 +res", 256, 1024, 1, 1, 16, True)
