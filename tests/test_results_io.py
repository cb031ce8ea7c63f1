"""Host side of SURVEY 8f-3: the .npy formats of the evaluation drivers and the metrics of the getResults scripts.
CPU tests (file names / dtypes / shapes, metric arithmetic against the reference's golden outputs); the compositions that
read these files run on the GPU and are tested in test_gpu_kitti.py / test_gpu_pair.py."""
import os

import numpy as np
import torch

from conftest import golden


def _fake_out(nH=2, h8=6, w8=8, seed=0):
    rs = np.random.RandomState(seed)
    return dict(H=rs.randn(nH, 3, 3).astype(np.float32), flowDown8=rs.randn(nH, 2, h8, w8).astype(np.float32),
                matchDown8=rs.rand(nH, 2, h8, w8).astype(np.float32), flow12=[None] * nH, match=[None] * nH)


def test_hpatch_file_format_round_trip(rf, tmp_path):
    """evaluation/evalHpatch/evaluation.py:244-260: four files per pair, hypothesis count in the name, fp32 / bool payloads."""
    fine, coarse = tmp_path / "fine", tmp_path / "coarse"
    fine.mkdir()
    coarse.mkdir()
    out = _fake_out(2)
    assert rf.results.save_pair(str(coarse), str(fine), 17, out) == 2
    assert sorted(os.listdir(fine)) == ["flow_17_2H.npy", "maskBG_17_2H.npy", "mask_17_2H.npy"] and os.listdir(coarse) == ["flow_17_2H.npy"]
    bg = np.load(fine / "maskBG_17_2H.npy")
    assert bg.dtype == bool and bg.shape == (48, 64) and bg.all()
    assert np.load(coarse / "flow_17_2H.npy").shape == (2, 3, 3) and np.load(fine / "flow_17_2H.npy").dtype == np.float32
    # the lookup the reference does on the directory listing (getResults.py:17-25)
    assert rf.results.find_nbH(17, os.listdir(fine)) == "2" and rf.results.find_nbH(1, os.listdir(fine)) is None
    flow, param, match = rf.results.load_pair(17, str(fine), str(coarse))
    assert np.array_equal(flow, out["flowDown8"]) and np.array_equal(param, out["H"]) and np.array_equal(match, out["matchDown8"])
    assert rf.results.load_pair(3, str(fine), str(coarse)) is None
    # nothing accepted -> nothing written (evaluation.py:244)
    assert rf.results.save_pair(str(coarse), str(fine), 18, dict(H=np.zeros((0,)), flowDown8=np.zeros((0,)), matchDown8=np.zeros((0,)))) == 0
    assert len(os.listdir(fine)) == 3


def test_kitti_file_format(rf, tmp_path):
    """evaluation/evalKITTI/evaluation.py:43-47,338-344 (note the reference's spelling 'Homograpy')."""
    rs = np.random.RandomState(1)
    out = dict(H=rs.randn(3, 3, 3).astype(np.float32), flow_d2=rs.randn(3, 2, 6, 20).astype(np.float32),
               mask=rs.rand(3, 2, 12, 40).astype(np.float32), flow=rs.randn(3, 2, 12, 40).astype(np.float32), size=(94, 310))
    assert rf.results.save_pair_kitti(str(tmp_path), 5, out) == 3
    assert sorted(os.listdir(tmp_path)) == ["BG_5_3H.npy", "Finetune_5_3.npy", "Finetune_D2_5_3.npy", "Finetune_Mask_5_3.npy", "Homograpy_5_3.npy"]
    assert np.load(tmp_path / "BG_5_3H.npy").shape == (94, 310) and np.load(tmp_path / "BG_5_3H.npy").dtype == bool
    assert rf.results.kitti_pairs(str(tmp_path)) == {"5": "3"}              # getResults.py:190-193
    assert np.array_equal(np.load(tmp_path / "Finetune_D2_5_3.npy"), out["flow_d2"])


def test_metrics_vs_reference(rf):
    g = golden("metrics")
    assert abs(rf.results.epe(torch.from_numpy(g["epe_in"]), torch.from_numpy(g["epe_tgt"])).item() - float(g["epe"])) < 1e-6
    wB, hB, wA, hA = [int(v) for v in g["dims"]]
    cnt, nb = rf.results.alignment_error(wB, hB, wA, hA, g["XA"], g["YA"], g["XB"], g["YB"], torch.from_numpy(g["flow"]),
                                         torch.from_numpy(g["match2"]), g["pixelGrid"])
    assert nb == int(g["nbAlign"]) and np.array_equal(np.asarray(cnt), g["counts"]) and g["counts"].max() > 0


def test_dense_epe_metrics(rf):
    """The inline metric code of the scripts (evalHpatch/getResults.py:224-250, evalKITTI/getResults.py:221-230), against a
    plain numpy transcription."""
    rs = np.random.RandomState(3)
    S = 24
    tgt = torch.from_numpy(rs.uniform(-1.3, 1.3, (1, S, S, 2)).astype(np.float32))
    est = tgt + torch.from_numpy(rs.randn(1, S, S, 2).astype(np.float32) * 0.01)
    t, e = tgt.numpy(), est.numpy()
    m = (t[..., 0] >= -1) & (t[..., 0] <= 1) & (t[..., 1] >= -1) & (t[..., 1] <= 1)
    tp, ep = (t + 1) * (S - 1) / 2, (e + 1) * (S - 1) / 2
    ref = np.sqrt(((tp - ep)[m] ** 2).sum(-1)).mean()
    assert abs(rf.results.epe_hpatches(est, tgt, S) - ref) < 1e-5
    # KITTI: flow in normalised coordinates, ground truth in pixels
    H, W = 10, 30
    gy, gx = np.meshgrid(np.linspace(-1, 1, H, dtype=np.float32), np.linspace(-1, 1, W, dtype=np.float32), indexing="ij")
    u, v = rs.randn(H, W) * 3, rs.randn(H, W)
    valid = rs.rand(H, W) > 0.3
    flow = np.stack([gx + (u + rs.randn(H, W) * 0.1).astype(np.float32) * 2 / (W - 1), gy + v.astype(np.float32) * 2 / (H - 1)], -1)[None].astype(np.float32)
    f = flow - np.stack([gx, gy], -1)[None]
    err = np.sqrt((f[0, :, :, 0] * (W - 1) / 2 - u) ** 2 + (f[0, :, :, 1] * (H - 1) / 2 - v) ** 2)
    ref = (err * valid).sum() / valid.sum()
    assert abs(rf.results.epe_kitti(torch.from_numpy(flow), u, v, valid) - ref) < 1e-6


def test_merge_first_wins_equals_the_reference_merge(rf):
    """pipeline.merge_first_wins is elementwise torch (device agnostic): on the CPU, fed with the oracle's per-hypothesis flows
    and matchabilities, it must reproduce the reference's golden flowGlobal / matchGlobal (evalCorr/getResults.py:121-134)."""
    import torch.nn.functional as F
    from oracle import warp_oracle as WO
    g = golden("get_flow_corr")
    flow, match = torch.from_numpy(g["flow"]), torch.from_numpy(g["mask"])
    h, w = flow.shape[2] * 8, flow.shape[3] * 8
    grid, coarse = WO.base_grid(h, w), WO.warp_grid(g["H"], h, w)
    flowUp = torch.clamp(F.interpolate(flow, scale_factor=8, mode="bilinear").permute(0, 2, 3, 1) + grid, min=-1, max=1)
    f = WO.grid_sample(coarse.permute(0, 3, 1, 2), flowUp).permute(0, 2, 3, 1).contiguous()
    m = F.interpolate(match, scale_factor=8, mode="bilinear")
    m = (m.narrow(1, 0, 1) * WO.grid_sample(m.narrow(1, 1, 1), flowUp) * WO.inside_mask(f)).permute(0, 2, 3, 1)
    fg, mg, mb = rf.pipeline.merge_first_wins(torch.clamp(f, min=-1, max=1), m, float(g["th"]), True)
    assert np.abs(fg.numpy() - g["flowGlobal"]).max() < 1e-6 and np.abs(mg.numpy() - g["matchGlobal"]).max() < 1e-6
    assert mb.dtype == torch.bool and 0.5 < mb.float().mean() < 1.0
    fg1, mg1, _ = rf.pipeline.merge_first_wins(torch.clamp(f, min=-1, max=1), m, float(g["th"]), False)
    assert torch.equal(fg1, torch.clamp(f, min=-1, max=1)[:1]) and torch.equal(mg1, m[:1])
