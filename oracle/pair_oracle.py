"""Oracle (test infrastructure): the whole per-pair path on the CPU in fp32.

Restates the ``CoarseAlign`` classes (variant A: evaluation/evalHpatch/
coarseAlignFeatMatch.py:35-179; variant C: quick_start/coarseAlignFeatMatch.py:
26-173), ``PredFlowMask`` (evaluation/evalHpatch/evaluation.py:23-55, evalCorr
variant evaluation/evalCorr/evaluation.py:29-59) and the multi-hypothesis driver
loop (evaluation/evalHpatch/evaluation.py:184-243) on top of the other oracle
modules.  This is what ``bench.py --impl reference`` / ``cpu_baseline`` time.
"""
import numpy as np
import PIL.Image as Image
import torch
import torch.nn.functional as F

from . import model_oracle as MO
from . import outil_oracle as OO
from . import warp_oracle as WO

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def scale_list(nbScale, scaleR):
    """coarseAlignFeatMatch.py:71-74 (all variants)."""
    if nbScale == 1:
        return [1]
    return (np.linspace(scaleR, 1, nbScale // 2 + 1).tolist()
            + np.linspace(1, 1 / scaleR, nbScale // 2 + 1).tolist()[1:])


def resized_size(w, h, minSize, strideNet=16, mode="min"):
    """ResizeMinSize (evalHpatch/coarseAlignFeatMatch.py:90-100) / ResizeMaxSize
    (quick_start/coarseAlignFeatMatch.py:80-90) target size."""
    if mode == "min":
        ratio = min(w / float(minSize), h / float(minSize))
    else:
        ratio = max(w / float(minSize), h / float(minSize))
    new_w, new_h = int(round(w / ratio)), int(round(h / ratio))
    return new_w // strideNet * strideNet, new_h // strideNet * strideNet


def resize(I, minSize, mode="min"):
    new_w, new_h = resized_size(I.size[0], I.size[1], minSize, 16, mode)
    return I.resize((new_w, new_h), resample=Image.LANCZOS)


def to_tensor(I):
    """torchvision ``ToTensor``: uint8 HWC -> fp32 CHW / 255."""
    a = torch.from_numpy(np.asarray(I, dtype=np.uint8).copy()).permute(2, 0, 1)
    return a.to(torch.float32).div(255)


def preproc(I):
    """ToTensor + Normalize(mean, std) (coarseAlignFeatMatch.py:63-66)."""
    t = to_tensor(I)
    mean = torch.tensor(MEAN).view(3, 1, 1)
    std = torch.tensor(STD).view(3, 1, 1)
    return (t - mean) / std


class CoarseAlignOracle:
    """Variant A when ``variant='A'`` (setPair / getCoarse(Mt) -> H | None),
    variant C when ``variant='C'`` (setSource / setTarget / getCoarse(Mt) -> (H, mask)),
    variant B (evaluation/evalYFCC/coarseAlignFeatMatch.py:35-196) = C's API with ResizeMinSize."""

    def __init__(self, resnet_sd, nbScale=7, nbIter=1000, tolerance=0.05, minSize=480, scaleR=2,
                 variant="A", seed=None, trunk=MO.resnet50_conv4):
        self.sd = resnet_sd
        self.nbIter, self.tolerance, self.minSize = nbIter, tolerance, minSize
        self.scaleList = scale_list(nbScale, scaleR)
        self.mode = "max" if variant == "C" else "min"      # B (evalYFCC) = C's API with ResizeMinSize
        self.variant = variant
        self.seed = seed
        self.trunk = trunk
        self.last_samples = None
        self.all_samples = []          # one (nbIter, 4) table per RANSAC call since the last setPair / setSource

    def _feat(self, I):
        return F.normalize(self.trunk(preproc(I).unsqueeze(0), self.sd))

    def _source(self, Is_org):
        IsList = [resize(Is_org, int(self.minSize * s), self.mode) for s in self.scaleList]
        self.Is = IsList[len(self.scaleList) // 2]
        self.IsTensor = to_tensor(self.Is).unsqueeze(0)
        feats, Ws, Hs = [], [], []
        for I in IsList:
            f = self._feat(I)
            W, H = OO.getWHTensor(f.shape[2], f.shape[3])
            feats.append(f.contiguous().view(f.shape[1], -1))
            Ws.append(W)
            Hs.append(H)
        self.featsMultiScale = torch.cat(feats, dim=1)
        self.WMultiScale = np.concatenate(Ws)
        self.HMultiScale = np.concatenate(Hs)

    def _target(self, It_org):
        self.It = resize(It_org, self.minSize, self.mode)
        self.ItTensor = to_tensor(self.It).unsqueeze(0)
        self.featt = self._feat(self.It)
        self.W2, self.H2 = self.featt.shape[2], self.featt.shape[3]
        self.Wt, self.Ht = OO.getWHTensor(self.W2, self.H2)
        self.WtInt, self.HtInt = OO.getWHTensor_Int(self.W2, self.H2)

    # variant C API
    def setSource(self, Is_org):
        self.all_samples = []
        self._source(Is_org)

    def setTarget(self, It_org):
        self._target(It_org)

    # variant A API
    def setPair(self, Is_org, It_org):
        self.all_samples = []
        self._source(Is_org)
        self._target(It_org)
        featt = self.featt.contiguous().view(self.featt.shape[1], -1)
        self.index1, self.index2 = OO.mutualMatching(self.featsMultiScale.numpy(), featt.numpy())

    def _mask16(self, Mt):
        MtExtend = torch.from_numpy((1 - Mt).astype(np.float32))[None, None]
        MtTensor = F.interpolate(MtExtend, size=(self.W2, self.H2), mode="bilinear")
        return (MtTensor > 0.5)[0, 0]

    def _ransac(self, match1, match2):
        if getattr(self, "raw_samples", None) is not None:
            # injected table of non-negative integers, reduced modulo the match count (what rf_ransac_homography's
            # RF_SAMPLES_MOD does with the same table): lets a CPU run and a CUDA run share one sample stream
            samples = (np.asarray(self.raw_samples, dtype=np.int64) % len(match1))
        else:
            if self.seed is not None:
                torch.manual_seed(self.seed)
            samples = torch.randint(len(match1), (self.nbIter, 4)).numpy()
        self.last_samples = samples
        self.all_samples.append(samples)
        return OO.RANSAC_from_samples(match1, match2, samples, self.tolerance)

    def getCoarse(self, Mt):
        m16 = self._mask16(Mt)
        one = lambda n: np.ones(n, dtype=np.float32)
        if self.variant == "A":
            valid = m16.numpy()[self.WtInt[self.index2], self.HtInt[self.index2]]
            i1, i2 = self.index1[valid], self.index2[valid]
        else:
            featt = (self.featt * m16.float()[None, None]).contiguous().view(self.featt.shape[1], -1)
            i1, i2 = OO.mutualMatching(self.featsMultiScale.numpy(), featt.numpy())
        match1 = np.stack([self.HMultiScale[i1], self.WMultiScale[i1], one(len(i1))], axis=1)
        match2 = np.stack([self.Ht[i2], self.Wt[i2], one(len(i2))], axis=1)
        self.match1, self.match2 = match1, match2
        none = None if self.variant == "A" else (None, [])
        if len(match1) < 4:
            return none
        best, _, isInlier, _ = self._ransac(match1, match2)
        if best is None:
            return none
        if self.variant == "A":
            return best.astype(np.float32)
        mask = OO.inlier_mask_grid(self.Wt, self.Ht, i2, isInlier, self.W2, self.H2)
        return best.astype(np.float32), mask


def pred_flow_mask(IsTensor, featt, flowCoarse, grid, net, with_match21=False):
    """PredFlowMask, evaluation/evalHpatch/evaluation.py:23-55 (``with_match21``:
    evaluation/evalCorr/evaluation.py:54).  ``net`` = dict of state_dicts."""
    IsSample = WO.grid_sample(IsTensor, flowCoarse)
    featsSample = F.normalize(MO.feature_extractor(IsSample, net["netFeatCoarse"]))
    corr12 = MO.corr_neigh(featt, featsSample)
    flowDown8 = MO.net_flow_coarse(corr12, net["netFlowCoarse"])
    match12Down8 = MO.net_matchability(corr12, net["netMatch"])
    corr21 = MO.corr_neigh(featsSample, featt)
    match21Down8 = MO.net_matchability(corr21, net["netMatch"])
    size = (grid.shape[1], grid.shape[2])
    match12 = WO.interpolate_bilinear(match12Down8, size)
    match21 = WO.interpolate_bilinear(match21Down8, size)
    flow12, flowUp = WO.compose_fine(flowDown8, flowCoarse, grid, clamp=True)
    match = match12
    if with_match21:
        match = match * WO.grid_sample(match21, flowUp)
    match = match * WO.inside_mask(flow12)
    return (flow12, match[0, 0].numpy(), flowDown8.numpy(),
            torch.cat((match12Down8, match21Down8), dim=1).numpy())


def align_pair(coarse, net, Is, It, maxCoarse=0, maskRegionTh=0.01, with_match21=False):
    """One pair through evaluation/evalHpatch/evaluation.py:172-243 (no segNet).

    Returns dict(H (nH,3,3), flowDown8 (nH,2,h8,w8), matchDown8 (nH,2,h8,w8),
    flow12 list, match list)."""
    coarse.setPair(Is, It)
    Itw, Ith = coarse.It.size
    It_bg = np.ones((Ith, Itw), dtype=np.float32)
    featt = F.normalize(MO.feature_extractor(coarse.ItTensor, net["netFeatCoarse"]))
    grid = WO.base_grid(Ith, Itw)
    Mask = np.zeros((Ith, Itw), dtype=np.float32)
    Hs, flows8, matches8, flows, matches = [], [], [], [], []
    nbCoarse = 0
    while nbCoarse <= maxCoarse:
        fgMask = ((Mask + (1 - It_bg)) > 0.5).astype(np.float32)
        bestPara = coarse.getCoarse(fgMask)
        if bestPara is None:
            break
        flowCoarse = WO.warp_grid(bestPara[None], Ith, Itw)
        flowFine, matchFine, f8, m8 = pred_flow_mask(coarse.IsTensor, featt, flowCoarse, grid, net, with_match21)
        if (matchFine * (1 - fgMask)).mean() > maskRegionTh or nbCoarse == 0:
            Hs.append(bestPara[None])
            flows8.append(f8)
            matches8.append(m8)
            flows.append(flowFine)
            matches.append(matchFine)
            nbCoarse += 1
            matchFine = matchFine if len(matches8) == 0 else matchFine * (1 - fgMask)
            Mask = ((Mask + matchFine) >= 1.0).astype(np.float32)
        else:
            break
    cat = lambda l: np.concatenate(l, axis=0) if l else np.zeros((0,))
    return dict(H=cat(Hs), flowDown8=cat(flows8), matchDown8=cat(matches8), flow12=flows, match=matches)


def align2images(coarse, net, img1, img2):
    """quick_start/align2images.py:53-97 (variant C coarse model, no clamp on the fine flow, netCorr(source, target))."""
    coarse.setSource(img1)
    coarse.setTarget(img2)
    w, h = coarse.It.size
    res = coarse.getCoarse(np.zeros((h, w)))
    if res[0] is None:
        return None
    bestPrm, inlierMask = res
    flowCoarse = WO.warp_grid(bestPrm[None], h, w)
    img1_coarse = WO.grid_sample(coarse.IsTensor, flowCoarse)
    feat1 = F.normalize(MO.feature_extractor(img1_coarse, net["netFeatCoarse"]))
    feat2 = F.normalize(MO.feature_extractor(coarse.ItTensor, net["netFeatCoarse"]))
    corr12 = MO.corr_neigh(feat1, feat2)
    flowDown = MO.net_flow_coarse(corr12, net["netFlowCoarse"])
    flow12, _ = WO.compose_fine(flowDown, flowCoarse, WO.base_grid(h, w), clamp=False)
    img1_fine = WO.grid_sample(coarse.IsTensor, flow12)
    return dict(bestPrm=bestPrm, inlierMask=inlierMask, flowCoarse=flowCoarse, img1_coarse=img1_coarse, flowDown=flowDown,
                flow12=flow12, img1_fine=img1_fine)


# --------------------------------------------------------------------------------------------------
# KITTI: two-level fine flow (evaluation/evalKITTI/evaluation.py)
# --------------------------------------------------------------------------------------------------
def resize_img(I, strideNet, minSize):
    """utils/outil.py:6-19 ``resizeImg`` (rounds to the nearest multiple of strideNet, unlike ResizeMinSize)."""
    w, h = I.size
    ratio = min(w / minSize, h / minSize)
    w, h = w / ratio, h / ratio
    return I.resize((round(w / strideNet) * strideNet, round(h / strideNet) * strideNet), resample=Image.LANCZOS)


def pred_flow_mask_kitti(IsSample, ItSample, flowCoarse, grid, net):
    """evaluation/evalKITTI/evaluation.py:49-81: both images' fine features computed inside, matchability always
    ``match12 * grid_sample(match21)`` * inside; ``flowCoarse`` may have another size than ``grid`` (second level)."""
    featsSample = F.normalize(MO.feature_extractor(IsSample, net["netFeatCoarse"]))
    featt = F.normalize(MO.feature_extractor(ItSample, net["netFeatCoarse"]))
    corr12 = MO.corr_neigh(featt, featsSample)
    flowDown8 = MO.net_flow_coarse(corr12, net["netFlowCoarse"])
    match12Down8 = MO.net_matchability(corr12, net["netMatch"])
    corr21 = MO.corr_neigh(featsSample, featt)
    match21Down8 = MO.net_matchability(corr21, net["netMatch"])
    size = (grid.shape[1], grid.shape[2])
    match12 = WO.interpolate_bilinear(match12Down8, size)
    match21 = WO.interpolate_bilinear(match21Down8, size)
    flow12, flowUp = WO.compose_fine(flowDown8, flowCoarse, grid, clamp=True)
    match = match12 * WO.grid_sample(match21, flowUp) * WO.inside_mask(flow12)
    return flow12, match[0, 0].numpy(), flowDown8, torch.cat((match12Down8, match21Down8), dim=1)


def align_pair_kitti(coarse, net, Is, It, fineSize=650, cc_th=0.01, maskRegionTh=0.005, maxH=None):
    """One pair through evaluation/evalKITTI/evaluation.py:216-336 (no segNet): coarse homography on the full-size pair,
    first fine level on the half-size target, second level on the resized target sampled on the ORIGINAL image's grid,
    small connected components of the matchability removed, multi-hypothesis mask update.  ``maxH`` caps the
    reference's ``while True`` (BASELINE config 5 caps it at 5).  Returns dict(H (nH,3,3), flow_d2 (nH,2,.,.),
    flow (nH,2,.,.), mask (nH,2,.,.)) - the four tensors the script saves (:338-344) - plus the per-hypothesis maps."""
    strideNet = 8
    It_resize = resize_img(It, strideNet, fineSize)
    It_d2 = resize_img(It, strideNet, fineSize // 2)
    w_org, h_org = It.size
    tensor_org, tensor_s = to_tensor(It).unsqueeze(0), to_tensor(Is).unsqueeze(0)
    grid_org = WO.base_grid(h_org, w_org)
    w_r, h_r = It_resize.size
    tensor_resize, grid_resize = to_tensor(It_resize).unsqueeze(0), WO.base_grid(h_r, w_r)
    w_d2, h_d2 = It_d2.size
    tensor_d2, grid_d2 = to_tensor(It_d2).unsqueeze(0), WO.base_grid(h_d2, w_d2)
    coarse.setPair(Is, It)
    It_bg = np.ones((h_org, w_org), dtype=np.float32)
    Mask = np.zeros((h_org, w_org), dtype=np.float32)
    Hs, D2, Msk, Fin, maps = [], [], [], [], []
    nbCoarse = 0
    while maxH is None or nbCoarse < maxH:
        fgMask = ((Mask + (1 - It_bg)) > 0.5).astype(np.float32)
        bestPara = coarse.getCoarse(fgMask)
        if bestPara is None:
            break
        bp = torch.from_numpy(bestPara).unsqueeze(0)
        homography_d2 = WO.warp_grid(bp, h_d2, w_d2)
        homography_resize = WO.warp_grid(bp, h_r, w_r)
        IsSample_d2 = WO.grid_sample(tensor_s, homography_d2)
        _, _, flowFine_d2, _ = pred_flow_mask_kitti(IsSample_d2, tensor_d2, homography_d2, grid_d2, net)
        flowCoarse, _ = WO.compose_fine(flowFine_d2, homography_resize, grid_resize, clamp=True)
        IsSample = WO.grid_sample(tensor_s, flowCoarse)
        flowFine_org, matchFine_org, f8, m8 = pred_flow_mask_kitti(IsSample, tensor_resize, flowCoarse, grid_org, net)
        matchFine = WO.remove_small_cc(matchFine_org, 0.99, cc_th)
        if ((matchFine > 0.9999) * (1 - fgMask)).mean() > maskRegionTh or nbCoarse == 0:
            Hs.append(bp.numpy())
            D2.append(flowFine_d2.numpy())
            Msk.append(m8.numpy())
            Fin.append(f8.numpy())
            maps.append((flowFine_org, matchFine.copy()))
            nbCoarse += 1
            mf = matchFine if len(Msk) == 0 else matchFine * (1 - fgMask)
            Mask = ((Mask + mf) > 0.9999).astype(np.float32)
        else:
            break
    cat = lambda l: np.concatenate(l, axis=0) if l else np.zeros((0,))
    return dict(H=cat(Hs), flow_d2=cat(D2), mask=cat(Msk), flow=cat(Fin), maps=maps, size=(h_org, w_org))
