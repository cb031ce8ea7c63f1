"""Per-kernel time of a few steady-state bench steps with torch.profiler (CUPTI): real (pipelined) kernel durations,
GPU busy fraction and host-side time per step.  Usage: python scripts/profile_step.py [fp32|tf32]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ransac_flow_b200 as rf  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "tf32"
rf.model.set_engine(engine)
rf.outil.corr_precision = {"fp32": 0, "tf32": 1}.get(engine, 2)
rsd, fe_sd, nf_sd, nm_sd = bench.states()
net = {"netFeatCoarse": rf.model.FeatureExtractor(), "netCorr": rf.model.CorrNeigh(7),
       "netFlowCoarse": rf.model.NetFlowCoarse(7), "netMatch": rf.model.NetMatchability(7)}
net["netFeatCoarse"].load_state_dict(fe_sd)
net["netFlowCoarse"].load_state_dict(nf_sd)
net["netMatch"].load_state_dict(nm_sd)
for m in net.values():
    m.cuda()
    m.eval()
coarse = rf.CoarseAlignA(7, 1000, 0.05, "Homography", 480, 2, False, 2, True, False, resnet_state_dict=rsd, verbose=False)
coarse.device_preproc = True
pairs = [(torch.from_numpy(s).cuda(), torch.from_numpy(t).cuda()) for s, t in bench.make_pairs(2)]


def step(i):
    s, t = pairs[i % 2]
    torch.manual_seed(1000)
    return rf.pipeline.align_pair_single(coarse, net, s, t)


for i in range(5):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    step(i)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 10
from torch.profiler import ProfilerActivity, profile  # noqa: E402
N = 4
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(N):
        step(i)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in ev) / N if ev and hasattr(ev[0], "device_time") else sum(e.cuda_time for e in ev) / N
agg = {}
for e in ev:
    d = e.device_time if hasattr(e, "device_time") else e.cuda_time
    k = e.name.split("(")[0].replace("void ", "")[:90]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += d
print("engine %s: wall %.3f ms/step; GPU kernel time %.3f ms/step (busy %.0f%%), %d kernels/step" % (engine, wall * 1e3, tot / 1e3, 100 * tot / 1e3 / (wall * 1e3), len(ev) // N))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
    print("%9.1f us/step %5.1f%% x%-4d %s" % (t / N, 100 * t / N / tot, c // N, k))
if "--seq" in sys.argv:
    evs = sorted(ev, key=lambda e: e.time_range.start)
    per = len(evs) // N
    print("---- kernel sequence of the last step (us) ----")
    for e in evs[-per:]:
        d = e.device_time if hasattr(e, "device_time") else e.cuda_time
        print("%8.1f  %s" % (d, e.name.split("(")[0].replace("void ", "")[:60]))
