"""Per-layer roofline floors of one config-2 pair on engine f16x3 (split operands) next to the measured duration of each
launch, from an ncu launch list of ONE pair (scripts/one_pair.py, launches serialised and cold-cache: compare shares).

  HBM floor    : split activations are 4 B / element (two fp16 planes): in + out (+ residual) + weights once, at the measured copy peak
  tensor floor : 3 kind::f16 MMAs per MAC at the measured fp16 peak

Usage: python scripts/split_floor_analysis.py gpurun_out/r2_launches_f16x3.csv > profiles/r2_split_floor_analysis.txt"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkf = os.path.join(ROOT, "MEASURED_PEAKS.json")
pk = json.load(open(pkf)) if os.path.exists(pkf) else {"hbm_gbs": 6575.1, "bf16_tflops": 1703.4}
HBM, TF = pk["hbm_gbs"] * 1e9, pk["bf16_tflops"] * 1e12
SCALES = [(960, 1280), (800, 1056), (640, 848), (480, 640), (400, 528), (320, 416), (240, 320), (480, 640)]   # 7 source scales + target


def out_hw(hw, k, s, p):
    return [((h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1) for h, w in hw]


def px(hw):
    return sum(h * w for h, w in hw)


layers = []          # (name, pixels_out, flops, bytes) in LAUNCH order of the tc_split kernels


def conv(name, hw, cin, cout, k, s, p, res=False, in_b=4, out_b=4):
    o = out_hw(hw, k, s, p)
    P = px(o)
    layers.append((name, P, 2.0 * P * cout * k * k * cin, px(hw) * cin * in_b + P * cout * out_b * (2 if res else 1) + cout * k * k * cin * 4))
    return o


def fe(tag, n_img=2):
    x = conv(tag + " stem (im2col 27->64) 64->64 1x1", [(480, 640)] * n_img, 64, 64, 1, 1, 0)
    x = out_hw(x, 4, 2, 1)
    cin = 64
    for layer, planes, stride in (("l1", 64, 1), ("l2", 128, 2), ("l3", 256, 2)):
        for b in range(2):
            s = stride if b == 0 else 1
            y = conv("%s %s.%d conv1 3x3/%d" % (tag, layer, b, s), x, cin, planes, 3, s, 1)
            r = x
            if b == 0 and stride != 1:
                r = conv("%s %s.%d shortcut 1x1" % (tag, layer, b), out_hw(x, 3, 2, 1), cin, planes, 1, 1, 0)
            x = conv("%s %s.%d conv2+res 3x3" % (tag, layer, b), y, planes, planes, 3, 1, 1, res=True)
            cin = planes
    return x


def trunk():
    x = out_hw(out_hw(SCALES, 7, 2, 3), 3, 2, 1)             # stem (own kernel) + max-pool
    cin = 64
    for layer, planes, blocks, stride in (("l1", 64, 3, 1), ("l2", 128, 4, 2), ("l3", 256, 6, 2)):
        for b in range(blocks):
            s = stride if b == 0 else 1
            y = conv("trunk %s.%d c1 1x1 %d->%d" % (layer, b, cin, planes), x, cin, planes, 1, 1, 0)
            y = conv("trunk %s.%d c2 3x3/%d" % (layer, b, s), y, planes, planes, 3, s, 1)
            if b == 0:
                # conv3 + the down-sampling 1x1 (stride s on the block's input) as ONE dual-input GEMM: reads y and the sampled x, writes once
                P = px(y)
                layers.append(("trunk %s.%d c3 + ds/%d fused %d|%d->%d" % (layer, b, s, planes, cin, planes * 4), P, 2.0 * P * planes * 4 * (planes + cin),
                               P * (planes + cin) * 4 + P * planes * 4 * 4 + planes * 4 * (planes + cin) * 4))
                x = y
            else:
                x = conv("trunk %s.%d c3+res 1x1 %d->%d" % (layer, b, planes, planes * 4), y, planes, planes * 4, 1, 1, 0, res=True)
            cin = planes * 4


def head(tag, n_img, cout_last):
    hw = [(60, 80)] * n_img
    x = conv(tag + " conv1 64(49)->512", hw, 64, 512, 3, 1, 1)
    x = conv(tag + " conv2 512->256", x, 512, 256, 3, 1, 1)
    x = conv(tag + " conv3 256->128", x, 256, 128, 3, 1, 1)
    conv(tag + " conv4 128->%d (fp32 out)" % cout_last, x, 128, cout_last, 3, 1, 1)


trunk()
fe("FE[source,target]")          # the sync-free path runs the FeatureExtractor once, on [warped source, target] as one batch
head("flow head", 1, 49)
head("match head", 2, 1)

times, other = [], {}
if len(sys.argv) > 1 and os.path.exists(sys.argv[1]):
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rows[1:]:
        us = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[mu], 1e-3)
        name = r[kn]
        if "tc_split_kernel" in name:
            times.append((name, us))
        else:
            key = name.split("(")[0].replace("void ", "").replace("rf::", "")[:48]
            a = other.setdefault(key, [0, 0.0])
            a[0] += 1
            a[1] += us

print("peaks: HBM %.0f GB/s, fp16 %.0f TF/s (MEASURED_PEAKS.json); engine f16x3: 4 B / activation element, 3 MMAs / MAC" % (HBM / 1e9, TF / 1e12))
print("%-46s %8s %7s %8s %8s %9s %9s %6s  %s" % ("layer (launch order of tc_split_kernel)", "pixels", "GFLOP", "MB", "hbm us", "tensor us", "meas us", "x", "kernel"))
groups = {}
for i, (name, P, fl, by) in enumerate(layers):
    th, tt = by / HBM * 1e6, 3 * fl / TF * 1e6
    kname, us = times[i] if i < len(times) else ("", float("nan"))
    targs = kname[kname.find("<") + 1:kname.find(">")].replace("(bool)", "").replace("(int)", "").replace(" ", "").split(",") if "<" in kname else []
    variant = "" if len(targs) < 2 else ("%s BN%s%s" % ("halo" if targs[0] in ("1", "true") else "tap", targs[1], " 2xstaging" if len(targs) > 2 and targs[2] in ("1", "true") else ""))
    fl_us = max(th, tt)
    print("%-46s %8d %7.2f %8.1f %8.1f %9.1f %9.1f %6.2f  %s" % (name, P, fl / 1e9, by / 1e6, th, tt, us, us / fl_us if us == us else float("nan"), variant))
    g = groups.setdefault(" ".join(name.split(" ")[:2]) if name.startswith(("flow", "match")) else name.split(" ")[0], [0.0, 0.0, 0.0, 0.0])
    g[0] += fl
    g[1] += by
    g[2] += fl_us
    g[3] += us if us == us else 0.0
print()
for gname, (fl, by, floor, meas) in groups.items():
    print("%-24s %7.1f GFLOP %8.1f MB   sum of per-layer floors %7.1f us   measured %7.1f us   (x %.2f)" % (gname, fl / 1e9, by / 1e6, floor, meas, meas / floor if floor else 0))
if len(times) != len(layers):
    print("\nWARNING: %d tc_split launches in the list, %d layers in the model of the pair" % (len(times), len(layers)))
if other:
    print("\nother kernels of the pair (count, total us):")
    for k, (c, t) in sorted(other.items(), key=lambda x: -x[1][1]):
        print("  %-50s x%-3d %8.1f" % (k, c, t))
