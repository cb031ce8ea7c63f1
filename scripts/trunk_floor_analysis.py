"""Roofline floors of every convolution of one pair (config 2, engine f16) - no GPU needed: for each layer the HBM floor
(fp16 activations in + out (+ residual) + weights once) and the tensor floor (2*P*Cout*K flops at the measured fp16 peak),
summed per network, next to the measured kernel time of the step profile.  Usage: python scripts/trunk_floor_analysis.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6575.1, "bf16_tflops": 1703.4}
HBM, TF = pk["hbm_gbs"] * 1e9, pk["bf16_tflops"] * 1e12
SCALES = [(960, 1280), (800, 1056), (640, 848), (480, 640), (400, 528), (320, 416), (240, 320), (480, 640)]   # 7 source scales + target (h, w)


def out_hw(hw, k, s, p):
    return [((h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1) for h, w in hw]


def px(hw):
    return sum(h * w for h, w in hw)


rows = []


def conv(name, hw, cin, cout, k, s, p, res=False, in_bytes=2, out_bytes=2):
    o = out_hw(hw, k, s, p)
    P = px(o)
    flops = 2.0 * P * cout * k * k * cin
    byt = px(hw) * cin * in_bytes + P * cout * out_bytes * (2 if res else 1) + cout * k * k * cin * 2
    rows.append((name, P, cin, cout, k, s, flops, byt))
    return o


def resnet(hw):
    x = conv("stem 7x7/2 3->64", hw, 3, 64, 7, 2, 3, in_bytes=4)
    x = out_hw(x, 3, 2, 1)                                   # max-pool (HBM only: counted separately below)
    cin = 64
    for layer, planes, blocks, stride in (("l1", 64, 3, 1), ("l2", 128, 4, 2), ("l3", 256, 6, 2)):
        for b in range(blocks):
            s = stride if b == 0 else 1
            y = conv("%s.%d c1" % (layer, b), x, cin, planes, 1, 1, 0)
            y = conv("%s.%d c2" % (layer, b), y, planes, planes, 3, s, 1)
            if b == 0:
                conv("%s.%d ds" % (layer, b), x, cin, planes * 4, 1, s, 0)
            x = conv("%s.%d c3+res" % (layer, b), y, planes, planes * 4, 1, 1, 0, res=True)
            cin = planes * 4
    return x


def feature_extractor(hw, tag):
    x = conv(tag + " conv1 3->64", hw, 3, 64, 3, 1, 1, in_bytes=4)
    x = out_hw(x, 4, 2, 1)                                   # maxpool(2,1) + blur/2
    cin = 64
    for layer, planes, stride in (("l1", 64, 1), ("l2", 128, 2), ("l3", 256, 2)):
        for b in range(2):
            s = stride if b == 0 else 1
            y = conv("%s %s.%d conv1" % (tag, layer, b), x, cin, planes, 3, s, 1)
            if b == 0 and stride != 1:
                conv("%s %s.%d shortcut 1x1" % (tag, layer, b), out_hw(x, 3, 2, 1), cin, planes, 1, 1, 0)
            x = conv("%s %s.%d conv2+res" % (tag, layer, b), y, planes, planes, 3, 1, 1, res=True)
            cin = planes
    return x


def head(hw, tag, n_img, cout_last):
    hw = hw * n_img
    x = conv(tag + " conv1 64(49)->512", hw, 64, 512, 3, 1, 1)
    x = conv(tag + " conv2 512->256", x, 512, 256, 3, 1, 1)
    x = conv(tag + " conv3 256->128", x, 256, 128, 3, 1, 1, out_bytes=4)
    conv(tag + " conv4 128->%d (tf32)" % cout_last, x, 128, cout_last, 3, 1, 1, in_bytes=4, out_bytes=4)


marks = {}
marks["resnet50 conv4, 8 images"] = (len(rows), None)
resnet(SCALES)
marks["resnet50 conv4, 8 images"] = (0, len(rows))
a = len(rows)
feature_extractor([(480, 640)], "FE(target)")
feature_extractor([(480, 640)], "FE(warped source)")
marks["FeatureExtractor x2"] = (a, len(rows))
a = len(rows)
head([(60, 80)], "flow head", 1, 49)
head([(60, 80)], "match head", 2, 1)
marks["heads (flow + 2x matchability)"] = (a, len(rows))

print("peaks: HBM %.0f GB/s, fp16 %.0f TF/s (MEASURED_PEAKS.json)" % (HBM / 1e9, TF / 1e12))
print("%-34s %9s %7s %9s %9s %9s  %s" % ("layer", "pixels", "GFLOP", "MB", "hbm us", "tensor us", "bound"))
tot = {}
for name, (i0, i1) in marks.items():
    sf = sb = sfloor = 0.0
    for r in rows[i0:i1]:
        n, P, cin, cout, k, s, fl, by = r
        th, tt = by / HBM * 1e6, fl / TF * 1e6
        print("%-34s %9d %7.2f %9.1f %9.1f %9.1f  %s" % (n, P, fl / 1e9, by / 1e6, th, tt, "hbm" if th > tt else "tensor"))
        sf += fl
        sb += by
        sfloor += max(th, tt)
    tot[name] = (sf, sb, sfloor)
print()
meas = {"resnet50 conv4, 8 images": 1803.2 - 74 - 46 - 42, "FeatureExtractor x2": 2 * 307.8, "heads (flow + 2x matchability)": None}
for name, (sf, sb, sfloor) in tot.items():
    print("%-34s %7.1f GFLOP %8.1f MB   sum of per-layer floors %7.1f us   (HBM alone %6.1f us, tensor alone %6.1f us)%s"
          % (name, sf / 1e9, sb / 1e6, sfloor, sb / HBM * 1e6, sf / TF * 1e6,
             "   measured ~%.0f us" % meas[name] if meas.get(name) else ""))
