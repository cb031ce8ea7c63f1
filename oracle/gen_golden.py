"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference)
on the CPU in the build container.  Test infrastructure; run once:

    python -m oracle.gen_golden            # needs /root/reference, writes tests/golden/

The reference hard-codes ``.cuda()`` (utils/outil.py:86, coarseAlignFeatMatch.py)
and pretrained downloads; the harness monkeypatches, on its own side only,
``torch.Tensor.cuda`` / ``nn.Module.cuda`` -> identity, ``torch.cuda.FloatTensor`` ->
``torch.FloatTensor``, ``torchvision.models.resnet50`` -> seeded random weights
(oracle/synth.py), ``torch.randint`` -> recorded/replayed samples, and stubs for
absent imports (``scipy.misc.imresize``, ``segEval``, ``kornia``).  Functions that
live in driver *scripts* (PredFlowMask, getFlow_all) are extracted from the
script's AST at run time and executed as they are.  No reference source is
copied into this repository; only inputs/outputs are stored.
"""
import ast
import contextlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

from . import synth
from . import warp_oracle as WO

REF = os.environ.get("RF_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def cpu_as_cuda():
    """Harness-side shims so the reference's hard-coded CUDA calls run on the CPU."""
    import torch.nn as nn
    saved = (torch.Tensor.cuda, nn.Module.cuda, getattr(torch.cuda, "FloatTensor", None), torch.cuda.empty_cache)
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.empty_cache = lambda: None
    try:
        yield
    finally:
        torch.Tensor.cuda, nn.Module.cuda = saved[0], saved[1]
        torch.cuda.FloatTensor = saved[2]
        torch.cuda.empty_cache = saved[3]


@contextlib.contextmanager
def replay_randint(samples_list, record):
    """Make ``torch.randint`` (utils/outil.py:120) return prepared sample arrays
    (or record what the real one draws when ``samples_list`` is None)."""
    real = torch.randint
    it = iter(samples_list) if samples_list is not None else None

    def fake(high, size, **kw):
        kw.pop("device", None)
        if it is not None:
            s = torch.from_numpy(np.asarray(next(it))).clone()
        else:
            s = real(high, size, **kw)
        record.append((int(high), s.numpy().copy()))
        return s
    torch.randint = fake
    try:
        yield
    finally:
        torch.randint = real


def ref_outil():
    return _load("ref_outil", os.path.join(REF, "utils", "outil.py"))


def ref_model():
    sys.path.insert(0, os.path.join(REF, "model"))
    try:
        return _load("ref_model", os.path.join(REF, "model", "model.py"))
    finally:
        sys.path.pop(0)


def extract_function(path, name, extra_ns):
    """Compile one top-level function of a reference *script* without running the script."""
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            ns = dict(extra_ns)
            exec(code, ns)
            return ns[name]
    raise KeyError(name)


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrs.items()})


# --------------------------------------------------------------------------
def gen_outil():
    outil = ref_outil()
    rs = np.random.RandomState(7)
    # getWHTensor / getWHTensor_Int (utils/outil.py:21-29)
    feat = torch.zeros(1, 4, 5, 7)
    W, H = outil.getWHTensor(feat)
    Wi, Hi = outil.getWHTensor_Int(feat)
    save("wh_tensor", h=np.int64(5), w=np.int64(7), W=W.numpy(), H=H.numpy(), Wi=Wi.numpy(), Hi=Hi.numpy())

    # mutualMatching (utils/outil.py:32-45): non-negative unit columns, one all-zero target column
    C, NA, NB = 48, 300, 70
    A = np.abs(rs.randn(C, NA)).astype(np.float32)
    B = np.abs(rs.randn(C, NB)).astype(np.float32)
    B[:, :30] = A[:, rs.permutation(NA)[:30]] + 0.05 * np.abs(rs.randn(C, 30)).astype(np.float32)
    A /= np.linalg.norm(A, axis=0, keepdims=True)
    B /= np.linalg.norm(B, axis=0, keepdims=True)
    B[:, 11] = 0
    i1, i2 = outil.mutualMatching(torch.from_numpy(A), torch.from_numpy(B))
    save("mutual_matching", featA=A, featB=B, index1=i1.numpy(), index2=i2.numpy())

    # Homography / Prediction / ScoreRANSAC / RANSAC (utils/outil.py:68-164)
    cases = {
        "ransac_m120": dict(seed=11, M=120, nbIter=1000, tol=0.05, frac=0.6),
        "ransac_m636": dict(seed=12, M=636, nbIter=1000, tol=0.05, frac=0.6),
        "ransac_grid": dict(seed=13, M=300, nbIter=1000, tol=0.05, frac=0.5, grid=(30, 40)),
        "ransac_remainder_only": dict(seed=14, M=80, nbIter=60, tol=0.05, frac=0.7),
        "ransac_none": dict(seed=15, M=60, nbIter=400, tol=0.0, frac=0.0),
        "ransac_lowinlier": dict(seed=16, M=200, nbIter=1000, tol=0.02, frac=0.15),
    }
    with cpu_as_cuda():
        for name, c in cases.items():
            m1, m2, Hgt = synth.make_matches(c["seed"], c["M"], c["frac"], grid=c.get("grid"))
            samples = synth.draw_samples(c["seed"], c["M"], c["nbIter"])
            rec = []
            with replay_randint([samples], rec):
                H, nb, inl, m2in = outil.RANSAC(c["nbIter"], torch.from_numpy(m1), torch.from_numpy(m2),
                                                c["tol"], 4, outil.Homography)
            # first chunk scores, for a per-hypothesis check
            us = samples[[len(set(r)) == 4 for r in samples.tolist()]]
            n0 = min(100, len(us))
            H0, cnt0 = outil.ScoreRANSAC(torch.from_numpy(m1), torch.from_numpy(m2), c["tol"],
                                         torch.from_numpy(us[:n0]), outil.Homography)
            dets0 = torch.det(H0)
            err0 = outil.Prediction(torch.from_numpy(m1)[None], torch.from_numpy(m2)[None], H0[:8])
            save(name, match1=m1, match2=m2, samples=samples, tol=np.float64(c["tol"]),
                 is_none=np.bool_(H is None),
                 H=(np.zeros((3, 3), np.float32) if H is None else H),
                 nbInlier=np.int64(0 if H is None else nb),
                 isInlier=(np.zeros(c["M"], bool) if H is None else inl),
                 chunk0_H=H0.numpy(), chunk0_counts=cnt0.numpy(), chunk0_dets=dets0.numpy(),
                 chunk0_err8=err0.numpy())


def gen_models():
    model = ref_model()
    torch.manual_seed(0)
    x = torch.rand(1, 3, 48, 64)
    fe = model.FeatureExtractor()
    fe.load_state_dict(synth.feature_extractor_state(0))
    fe.eval()
    y = fe(x)
    save("feature_extractor", x=x.numpy(), y=y.numpy(), seed=np.int64(0))

    a = F.normalize(torch.randn(1, 256, 6, 8))
    b = F.normalize(torch.randn(1, 256, 6, 8))
    corr = model.CorrNeigh(7).eval()(a, b)
    nf = model.NetFlowCoarse(7)
    nf.load_state_dict(synth.net_flow_coarse_state(1))
    nf.eval()
    nm = model.NetMatchability(7)
    nm.load_state_dict(synth.net_matchability_state(2))
    nm.eval()
    save("fine_heads", a=a.numpy(), b=b.numpy(), corr=corr.numpy(), flow=nf(corr, False).numpy(),
         match=nm(corr, False).numpy())

    import torchvision
    net = torchvision.models.resnet50(weights=None)
    missing = net.load_state_dict(synth.resnet50_conv4_state(0), strict=False)
    assert all(k.startswith(("layer4", "fc")) for k in missing.missing_keys), missing
    trunk = torch.nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool, net.layer1, net.layer2, net.layer3).eval()
    xr = torch.randn(1, 3, 64, 96)
    with torch.no_grad():
        yr = trunk(xr)
    save("resnet50_conv4", x=xr.numpy(), y=yr.numpy(), seed=np.int64(0))


def _ref_networks(model):
    net = {"netFeatCoarse": model.FeatureExtractor(), "netCorr": model.CorrNeigh(7),
           "netFlowCoarse": model.NetFlowCoarse(7), "netMatch": model.NetMatchability(7)}
    net["netFeatCoarse"].load_state_dict(synth.feature_extractor_state(0))
    net["netFlowCoarse"].load_state_dict(synth.net_flow_coarse_state(1))
    net["netMatch"].load_state_dict(synth.net_matchability_state(2))
    for m in net.values():
        m.eval()
    return net


def gen_pred_flow_mask():
    """PredFlowMask as written in the two driver scripts, on a 48x64 pair."""
    model = ref_model()
    net = _ref_networks(model)
    src, tgt, Hgt = synth.make_pair(3, 48, 64)
    Is = torch.from_numpy(src).permute(2, 0, 1)[None].float() / 255
    It = torch.from_numpy(tgt).permute(2, 0, 1)[None].float() / 255
    Hm = torch.tensor(Hgt / np.linalg.norm(Hgt), dtype=torch.float32)[None]
    grid = WO.base_grid(48, 64)
    flowCoarse = WO.warp_grid(Hm, 48, 64)
    with cpu_as_cuda(), torch.no_grad():
        featt = F.normalize(net["netFeatCoarse"](It))
        for tag, script in (("hpatch", "evaluation/evalHpatch/evaluation.py"), ("corr", "evaluation/evalCorr/evaluation.py")):
            fn = extract_function(os.path.join(REF, script), "PredFlowMask", {"torch": torch, "F": F})
            flow12, match, f8, m8 = fn(Is, featt, flowCoarse, grid, net)
            save("pred_flow_mask_" + tag, Is=Is.numpy(), It=It.numpy(), H=Hm.numpy(), flow12=flow12.numpy(),
                 match=match, flowDown8=f8, matchDown8=m8)
    return f8, m8, Hm


def gen_get_flow(tmpdir="/tmp/rf_golden_getflow"):
    """getFlow_all (evaluation/evalHpatch/getResults.py:16-63) on two hypotheses."""
    os.makedirs(tmpdir + "/fine", exist_ok=True)
    os.makedirs(tmpdir + "/coarse", exist_ok=True)
    rs = np.random.RandomState(5)
    flow = (rs.randn(2, 2, 6, 8) * 0.02).astype(np.float32)
    mask = rs.rand(2, 2, 6, 8).astype(np.float32)
    Hs = np.stack([np.eye(3) + rs.uniform(-0.05, 0.05, (3, 3)) for _ in range(2)]).astype(np.float32)
    np.save(tmpdir + "/fine/flow_0_2H.npy", flow)
    np.save(tmpdir + "/fine/mask_0_2H.npy", mask)
    np.save(tmpdir + "/coarse/flow_0_2H.npy", Hs)

    class Warper:                      # kornia is absent: the oracle's restatement stands in
        def __init__(self, h, w):
            self.h, self.w = h, w

        def warp_grid(self, H):
            return WO.warp_grid(H, self.h, self.w)
    fn = extract_function(os.path.join(REF, "evaluation/evalHpatch/getResults.py"), "getFlow_all",
                          {"torch": torch, "F": F, "np": np, "os": os})
    outH, outW = 40, 56
    grid = WO.base_grid(outH, outW)
    fg = fn(0, tmpdir + "/fine", tmpdir + "/coarse", ["flow_0_2H.npy"], True, Warper(outH, outW), grid, 0.5, outW, outH)
    save("get_flow_all", flow=flow, mask=mask, H=Hs, flowGlobal=fg.numpy(), th=np.float64(0.5))


def _coarse_align_common(path, name, stubs):
    for k, v in stubs.items():
        sys.modules.setdefault(k, v)
    sys.path.insert(0, os.path.join(REF, "utils"))
    try:
        return _load(name, path)
    finally:
        sys.path.pop(0)


def gen_coarse_align():
    """CoarseAlign variant C (quick_start) and variant A (evalHpatch) on a 96x128 pair."""
    import PIL.Image as Image
    import torchvision
    src, tgt, Hgt = synth.make_pair(5, 96, 128)
    Is, It = Image.fromarray(src), Image.fromarray(tgt)
    real_resnet50 = torchvision.models.resnet50

    def seeded_resnet50(*a, **k):
        net = real_resnet50(weights=None)
        net.load_state_dict(synth.resnet50_conv4_state(0), strict=False)
        return net
    torchvision.models.resnet50 = seeded_resnet50
    seg = types.ModuleType("segEval")
    res = types.ModuleType("resnet50")
    res.resnet50 = seeded_resnet50
    misc = types.ModuleType("scipy.misc")
    misc.imresize = None
    import scipy
    scipy.misc = misc
    try:
        with cpu_as_cuda():
            # variant C
            modC = _coarse_align_common(os.path.join(REF, "quick_start/coarseAlignFeatMatch.py"), "ref_coarse_C", {})
            cC = modC.CoarseAlign(3, 500, 0.05, "Homography", 128, scaleR=1.5)
            cC.setSource(Is)
            cC.setTarget(It)
            rec = []
            torch.manual_seed(1000)
            with replay_randint(None, rec):
                H, mask = cC.getCoarse(np.zeros((It.size[1], It.size[0])))
            assert H is not None
            save("coarse_align_C", src=src, tgt=tgt, H=H, inlierMask=mask, samples=rec[0][1], nbMatch=np.int64(rec[0][0]),
                 featt=cC.featt.numpy(), feats_sum=cC.featsMultiScale.sum(0).numpy(),
                 WMulti=cC.WMultiScale.numpy(), HMulti=cC.HMultiScale.numpy(),
                 Is=np.asarray(cC.Is), It=np.asarray(cC.It))
            # variant A
            modA = _coarse_align_common(os.path.join(REF, "evaluation/evalHpatch/coarseAlignFeatMatch.py"), "ref_coarse_A",
                                        {"segEval": seg, "resnet50": res, "scipy.misc": misc})
            cA = modA.CoarseAlign(3, 500, 0.05, "Homography", 96, 2, False, 1.5, True, False)
            cA.setPair(Is, It)
            Mt = np.zeros((cA.It.size[1], cA.It.size[0]), dtype=np.float32)
            Mt[:, : cA.It.size[0] // 4] = 1          # mask out the left quarter of the target
            rec = []
            with replay_randint(None, rec):
                torch.manual_seed(1000)
                HA0 = cA.getCoarse(np.zeros_like(Mt))
                torch.manual_seed(1000)
                HA1 = cA.getCoarse(Mt)
            save("coarse_align_A", src=src, tgt=tgt, H0=HA0, H1=HA1, Mt=Mt,
                 samples0=rec[0][1], samples1=rec[1][1], nbMatch0=np.int64(rec[0][0]), nbMatch1=np.int64(rec[1][0]),
                 W1=cA.W1MutualMatch.numpy(), H1m=cA.H1MutualMatch.numpy(), W2=cA.W2MutualMatch.numpy(),
                 H2m=cA.H2MutualMatch.numpy(), W2i=cA.W2MutualMatchInt.numpy(), H2i=cA.H2MutualMatchInt.numpy(),
                 Is=np.asarray(cA.Is), It=np.asarray(cA.It))
    finally:
        torchvision.models.resnet50 = real_resnet50


def main():
    assert os.path.isdir(REF), "reference checkout not found at %s" % REF
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    gen_outil()
    gen_models()
    gen_pred_flow_mask()
    gen_get_flow()
    gen_coarse_align()


if __name__ == "__main__":
    main()
