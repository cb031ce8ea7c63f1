"""The on-disk formats of the evaluation drivers and the metrics of the getResults scripts, so
that results written here are read by the reference's ``getResults.py`` and vice versa.

  save_pair / load_pair            : evaluation/evalHpatch/evaluation.py:244-260 (identical in evalCorr :244-260)
  getFlow_all_from_files           : evaluation/evalHpatch/getResults.py:16-63 (file lookup + np.load + composition)
  getFlow_from_files               : evaluation/evalCorr/getResults.py:78-134 / evalYFCC/getResults.py:132-190 (flow AND matchability)
  save_pair_kitti / kitti_pairs    : evaluation/evalKITTI/evaluation.py:43-47,338-344; evalKITTI/getResults.py:190-193
  getFlow_all_kitti_from_files     : evaluation/evalKITTI/getResults.py:95-141
  epe_hpatches                     : evaluation/evalHpatch/getResults.py:147-157,224-250
  alignment_error                  : evaluation/evalCorr/getResults.py:15-38
  epe_kitti                        : evaluation/evalKITTI/getResults.py:221-230

File IO and the sparse-keypoint lookups are host code as in the reference; the dense compositions run on the library's
kernels (``pipeline.getFlow_all`` / ``getFlow_all_kitti``), the dense metrics are elementwise torch on the tensors' device.
"""
import os

import numpy as np
import torch

from . import pipeline


# --------------------------------------------------------------------------- evalHpatch / evalCorr / evalYFCC
def save_pair(outCoarse, outFine, idx, out, It_bg=None):
    """Write what evaluation.py:244-260 writes for pair ``idx``: ``outFine/maskBG_{idx}_{nH}H.npy`` (bool (H,W)),
    ``outFine/mask_{idx}_{nH}H.npy`` (nH,2,h8,w8) fp32, ``outCoarse/flow_{idx}_{nH}H.npy`` (nH,3,3) fp32 homographies and
    ``outFine/flow_{idx}_{nH}H.npy`` (nH,2,h8,w8) fp32.  ``out`` is the dict ``pipeline.align_pair*`` returns.  Nothing is
    written when no hypothesis was accepted (evaluation.py:244).  Returns nH."""
    nH = len(out["H"])
    if nH == 0:
        return 0
    f8 = np.asarray(out["flowDown8"], dtype=np.float32)
    if It_bg is None:
        It_bg = np.ones((f8.shape[2] * 8, f8.shape[3] * 8), dtype=np.float32)
    tag = "%s_%dH.npy" % (str(idx), nH)
    np.save(os.path.join(outFine, "maskBG_" + tag), np.asarray(It_bg).astype(bool))
    np.save(os.path.join(outFine, "mask_" + tag), np.asarray(out["matchDown8"], dtype=np.float32))
    np.save(os.path.join(outCoarse, "flow_" + tag), np.asarray(out["H"], dtype=np.float32))
    np.save(os.path.join(outFine, "flow_" + tag), f8)
    return nH


def find_nbH(pairID, flowList):
    """getResults.py:17-25: the hypothesis count encoded in the file name of pair ``pairID`` (None when absent)."""
    for flowName in flowList:
        parts = flowName.split("_")
        if len(parts) >= 3 and parts[1] == str(pairID):
            return parts[2].split("H")[0]
    return None


def load_pair(pairID, finePath, coarsePath, flowList=None):
    """np.load of the three tensors getFlow_all reads (getResults.py:27-36): (flow, param, match) or None."""
    nbH = find_nbH(pairID, os.listdir(finePath) if flowList is None else flowList)
    if nbH is None:
        return None
    tag = "{}_{}H.npy".format(pairID, nbH)
    return (np.load(os.path.join(finePath, "flow_" + tag)).astype(np.float32),
            np.load(os.path.join(coarsePath, "flow_" + tag)).astype(np.float32),
            np.load(os.path.join(finePath, "mask_" + tag)).astype(np.float32))


def getFlow_all_from_files(pairID, finePath, coarsePath, flowList, multiH, th, outW, outH, with_match21=False):
    """evaluation/evalHpatch/getResults.py:16-63 with its argument order minus ``warper`` / ``grid`` (regenerated on the
    device): returns flowGlobal (1,outH,outW,2) CUDA, or [] when the pair has no files (``with_match21``: the evalCorr /
    evalYFCC composition, evalCorr/getResults.py:78-136)."""
    t = load_pair(pairID, finePath, coarsePath, flowList)
    if t is None:
        return []
    flow, param, match = t
    return pipeline.getFlow_all(flow, param, match, outH, outW, th=th, multiH=multiH, with_match21=with_match21)


def getFlow_from_files(pairID, finePath, flowList, coarsePath, maskPath, multiH, th):
    """evaluation/evalCorr/getResults.py:78-134 ``getFlow`` with its own argument order (``maskPath`` only serves the
    reference's unused ``maskBG`` load): (flowGlobal, matchGlobal) CUDA at 8x the saved resolution, or ([], [])."""
    t = load_pair(pairID, finePath, coarsePath, flowList)
    if t is None:
        return [], []
    flow, param, match = t
    return pipeline.getFlow_corr(flow, param, match, th=th, multiH=multiH)


# --------------------------------------------------------------------------- evalKITTI
def save_pair_kitti(outDir, i, out, It_bg=None):
    """evaluation/evalKITTI/evaluation.py:338-344 (``save_output`` :43-47): ``Homograpy_{i}_{nH}.npy`` [sic] (nH,3,3),
    ``BG_{i}_{nH}H.npy`` bool, ``Finetune_D2_{i}_{nH}.npy``, ``Finetune_Mask_{i}_{nH}.npy``, ``Finetune_{i}_{nH}.npy`` (fp32).
    ``out`` is the dict ``pipeline.align_pair_kitti`` returns.  Returns nH."""
    nH = len(out["flow"])
    if nH == 0:
        return 0
    if It_bg is None:
        It_bg = np.ones(out["size"], dtype=np.float32)
    np.save(os.path.join(outDir, "Homograpy_{}_{}.npy".format(i, nH)), np.asarray(out["H"], dtype=np.float32))
    np.save(os.path.join(outDir, "BG_" + str(i) + "_{:d}H.npy".format(nH)), np.asarray(It_bg).astype(bool))
    np.save(os.path.join(outDir, "Finetune_D2_{}_{}.npy".format(i, nH)), np.asarray(out["flow_d2"], dtype=np.float32))
    np.save(os.path.join(outDir, "Finetune_Mask_{}_{}.npy".format(i, nH)), np.asarray(out["mask"], dtype=np.float32))
    np.save(os.path.join(outDir, "Finetune_{}_{}.npy".format(i, nH)), np.asarray(out["flow"], dtype=np.float32))
    return nH


def kitti_pairs(predDir):
    """evaluation/evalKITTI/getResults.py:190-193: {pair id: nbH} from the ``BG_*`` files of a prediction directory."""
    bg = [item for item in os.listdir(predDir) if "BG" in item]
    return dict((item.split("_")[1], item.split("_")[2].split("H")[0]) for item in bg)


def getFlow_all_kitti_from_files(pairID, predDir, nbH, res_name, Ith, Itw, multiH, th, cc_th, interpolate=False):
    """evaluation/evalKITTI/getResults.py:95-141 (``warper_org`` / ``grid_org`` regenerated on the device from the ground
    truth's size): flowGlobal (1,Ith,Itw,2) CUDA."""
    ld = lambda name: np.load(os.path.join(predDir, name)).astype(np.float32)
    param = ld("Homograpy_{}_{}.npy".format(pairID, nbH))
    flowd2 = ld("{}_D2_{}_{}.npy".format(res_name, pairID, nbH))
    flow = ld("{}_{}_{}.npy".format(res_name, pairID, nbH))
    match = ld("{}_Mask_{}_{}.npy".format(res_name, pairID, nbH))
    fg, _ = pipeline.getFlow_all_kitti(param, flowd2, flow, match, Ith, Itw, th=th, cc_th=cc_th, multiH=multiH, interpolate=interpolate)
    return fg


# --------------------------------------------------------------------------- metrics
def epe(input_flow, target_flow):
    """evaluation/evalHpatch/getResults.py:147-157."""
    return torch.norm(target_flow - input_flow, p=2, dim=1).mean()


def epe_hpatches(flow_est, flow_target, minSize):
    """evaluation/evalHpatch/getResults.py:224-250: average end-point error in pixels of a ``minSize`` x ``minSize`` image
    over the pixels whose ground-truth correspondence falls inside the image.  flow_est, flow_target: (1,H,W,2) in
    normalised [-1, 1] coordinates (any device)."""
    flow_target = flow_target.to(flow_est.device)
    mask = (flow_target[..., 0].ge(-1) & flow_target[..., 0].le(1)) & (flow_target[..., 1].ge(-1) & flow_target[..., 1].le(1))
    ft = (flow_target + 1) * (minSize - 1) / (1 + 1)
    fe = (flow_est + 1) * (minSize - 1) / (1 + 1)
    ft = torch.cat((ft[..., 0][mask].unsqueeze(1), ft[..., 1][mask].unsqueeze(1)), dim=1)
    fe = torch.cat((fe[..., 0][mask].unsqueeze(1), fe[..., 1][mask].unsqueeze(1)), dim=1)
    return epe(fe, ft).item()


def alignment_error(wB, hB, wA, hA, XA, YA, XB, YB, flow, match2, pixelGrid):
    """evaluation/evalCorr/getResults.py:15-38 (host code in the reference too: a lookup at the annotated keypoints):
    number of keypoints of the target aligned within each threshold of ``pixelGrid`` (1, T) and the number of keypoints
    covered by the matchability mask.  flow (1,hB,wB,2), match2 (1,hB,wB,1) or broadcastable; CUDA tensors are copied once."""
    flow = flow.detach().cpu()
    estimX = flow.narrow(3, 1, 1).reshape(hB, wB).numpy()
    estimY = flow.narrow(3, 0, 1).reshape(hB, wB).numpy()
    estimY = (estimY + 1) * 0.5 * (wA - 1)
    estimX = (estimX + 1) * 0.5 * (hA - 1)
    match = torch.as_tensor(match2).detach().cpu().squeeze().numpy()
    xa, ya, xb, yb = XA.astype(np.int64), YA.astype(np.int64), XB.astype(np.int64), YB.astype(np.int64)
    index = np.where(match[yb, xb] > 0.5)[0]
    nbAlign = len(index)
    if nbAlign > 0:
        xa, ya, xb, yb = xa[index], ya[index], xb[index], yb[index]
        pixelDiff = ((estimY[yb, xb] - xa) ** 2 + (estimX[yb, xb] - ya) ** 2) ** 0.5
        pixelDiffT = np.sum(pixelDiff.reshape((-1, 1)) <= pixelGrid, axis=0)
    else:
        pixelDiffT = np.zeros(pixelGrid.shape[1])
    return pixelDiffT, nbAlign


def epe_kitti(flow, u, v, valid):
    """evaluation/evalKITTI/getResults.py:221-230: flow (1,Ith,Itw,2) normalised target->source grid, ground truth (u, v)
    in pixels, ``valid`` mask -> average end-point error over the valid pixels."""
    Ith, Itw = u.shape
    dev = flow.device
    gy = torch.linspace(-1, 1, steps=Ith, device=dev).view(1, -1, 1, 1).expand(1, Ith, Itw, 1)
    gx = torch.linspace(-1, 1, steps=Itw, device=dev).view(1, 1, -1, 1).expand(1, Ith, Itw, 1)
    f = flow - torch.cat((gx, gy), dim=3)                          # fp32, as the reference's numpy arrays
    upred = (f[0, :, :, 0] * (Itw - 1) / 2).double()
    vpred = (f[0, :, :, 1] * (Ith - 1) / 2).double()
    u_t, v_t = torch.as_tensor(u, dtype=torch.float64, device=dev), torch.as_tensor(v, dtype=torch.float64, device=dev)
    val = torch.as_tensor(np.asarray(valid, dtype=np.float64), device=dev)
    error = ((upred - u_t) ** 2 + (vpred - v_t) ** 2) ** 0.5
    return (torch.sum(error * val) / torch.sum(val)).item()
