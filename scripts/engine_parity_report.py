"""Fine-flow stage (PredFlowMask) of every engine against the UNMODIFIED reference's golden outputs (tests/golden/
pred_flow_mask_*.npz: fixed coarse homography, 48 x 64 pair): prints the errors the parity tests assert on.
Usage: python scripts/engine_parity_report.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ransac_flow_b200 as rf  # noqa: E402
import synthdata as synth  # noqa: E402
from conftest import golden  # noqa: E402
from oracle import warp_oracle as WO  # noqa: E402  (checker only)


def networks():
    net = {"netFeatCoarse": rf.model.FeatureExtractor(), "netCorr": rf.model.CorrNeigh(7),
           "netFlowCoarse": rf.model.NetFlowCoarse(7), "netMatch": rf.model.NetMatchability(7)}
    net["netFeatCoarse"].load_state_dict(synth.feature_extractor_state(0))
    net["netFlowCoarse"].load_state_dict(synth.net_flow_coarse_state(1))
    net["netMatch"].load_state_dict(synth.net_matchability_state(2))
    for m in net.values():
        m.cuda()
        m.eval()
    return net


for engine in ("fp32", "tf32", "f16"):
    for tag, m21 in (("hpatch", False), ("corr", True)):
        g = golden("pred_flow_mask_" + tag)
        rf.model.set_engine(engine)
        net = networks()
        Is, It = torch.from_numpy(g["Is"]).cuda(), torch.from_numpy(g["It"]).cuda()
        featt = torch.nn.functional.normalize(net["netFeatCoarse"](It))
        Hh, Ww = 48, 64
        flowCoarse = rf.kornia_geometry.HomographyWarper(Hh, Ww).warp_grid(torch.from_numpy(g["H"]).cuda())
        grid = rf.pipeline.base_grid(Hh, Ww)
        flow12, match, f8, m8 = rf.pipeline.PredFlowMask(Is, featt, flowCoarse, grid, net, with_match21=m21)
        f12 = flow12.cpu().numpy()
        _, flowUp = WO.compose_fine(torch.from_numpy(g["flowDown8"][:1]), WO.warp_grid(g["H"][:1], Hh, Ww), WO.base_grid(Hh, Ww), clamp=True)
        fu = flowUp[0].numpy()
        interior = (np.abs(fu[..., 0]) < 1 - 4.0 / Ww) & (np.abs(fu[..., 1]) < 1 - 4.0 / Hh)
        d = np.abs(f12 - g["flow12"])[0]
        far = (np.abs(np.abs(g["flow12"]) - 1) > 1e-3).all(-1)[0]
        print("[%-4s %-6s] |flowDown8-ref| %.3g  |matchDown8-ref| %.3g  |flow12-ref| all %.3g interior(%.0f%%) %.3g  |match-ref| %.3g"
              % (engine, tag, np.abs(f8 - g["flowDown8"]).max(), np.abs(m8 - g["matchDown8"]).max(), d.max(), 100 * interior.mean(),
                 d[interior].max() if interior.any() else float("nan"), np.abs(match - g["match"])[far].max()))
rf.model.set_engine("fp32")
