"""One steady-state config-2 pair, eager launches, for ncu: the profiler range (cudaProfilerStart / Stop) covers exactly the
third pair, so `ncu --profile-from-start off ...` sees the ~130 kernels of one pair with warm caches of everything that is
built once (TMA maps, layer programs, folded weights).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
        python scripts/one_pair.py f16x3
    ncu --set full --clock-control none --profile-from-start off -k regex:'tc_split|stem7_split|tc_corr_pipe' -o gpurun_out/pair \
        python scripts/one_pair.py f16x3
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ransac_flow_b200 as rf  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
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
for i in range(2):
    torch.manual_seed(1000)
    rf.pipeline.align_pair_single(coarse, net, *pairs[i % 2])
torch.cuda.synchronize()
torch.cuda.profiler.start()
torch.manual_seed(1000)
out = rf.pipeline.align_pair_single(coarse, net, *pairs[0])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("pair done: %d matches, %d inliers" % (out["nbMatch"], out["nbInlier"]))
