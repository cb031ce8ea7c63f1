"""Oracle (test infrastructure) for the warp / composition arithmetic.

``warp_grid`` restates kornia==0.1.4.post2 ``HomographyWarper.warp_grid``
(third-party, NOT in /root/reference, not installed: *parity unpinned*, see
SURVEY.md section 8c).  The library calls the reference makes
(``F.grid_sample``, ``F.interpolate``) are made here on the CPU in fp32.
"""
import numpy as np
import torch
import torch.nn.functional as F


def base_grid(h, w):
    """(1,h,w,2) grid with last dim (x, y), x = linspace(-1,1,w)[c], y = linspace(-1,1,h)[r].
    Same tensor the drivers build by hand (evaluation/evalHpatch/evaluation.py:187-189)
    and kornia's ``create_meshgrid(h, w, normalized_coordinates=True)``."""
    gy = torch.linspace(-1, 1, steps=h).view(1, -1, 1, 1).expand(1, h, w, 1)
    gx = torch.linspace(-1, 1, steps=w).view(1, 1, -1, 1).expand(1, h, w, 1)
    return torch.cat((gx, gy), dim=3).contiguous()


def warp_grid(Hm, h, w):
    """kornia 0.1.4 ``HomographyWarper(h, w).warp_grid(H)``: (N,3,3) -> (N,h,w,2).

    ``transform_points(H, grid)`` = homogeneous multiply then divide by z, no
    inversion of H, no epsilon.  Call sites: quick_start/align2images.py:61,65;
    evaluation/evalHpatch/evaluation.py:190,218."""
    Hm = torch.as_tensor(Hm, dtype=torch.float32).reshape(-1, 3, 3)
    g = base_grid(h, w)[0]                                   # (h,w,2)
    x, y = g[..., 0][None], g[..., 1][None]                  # (1,h,w)
    Hn = Hm[:, :, :, None, None]
    px = Hn[:, 0, 0] * x + Hn[:, 0, 1] * y + Hn[:, 0, 2]
    py = Hn[:, 1, 0] * x + Hn[:, 1, 1] * y + Hn[:, 1, 2]
    pz = Hn[:, 2, 0] * x + Hn[:, 2, 1] * y + Hn[:, 2, 2]
    return torch.stack((px / pz, py / pz), dim=-1)


def grid_sample(inp, grid, align_corners=False):
    """``F.grid_sample(inp, grid)``: bilinear, zeros padding; the torch default
    ``align_corners`` (False since torch 1.3)."""
    return F.grid_sample(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=align_corners)


def interpolate_bilinear(x, size):
    """``F.interpolate(x, size=size, mode='bilinear')`` (align_corners=False)."""
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


def inside_mask(flow12):
    """evaluation/evalHpatch/evaluation.py:51: 1 where both coords are in [-1, 1]."""
    fx, fy = flow12[..., 0:1], flow12[..., 1:2]
    m = ((fx >= -1) * (fx <= 1)).float() * ((fy >= -1) * (fy <= 1)).float()
    return m.permute(0, 3, 1, 2)


def compose_fine(flowDown8, flowCoarse, grid, clamp=True):
    """evaluation/evalHpatch/evaluation.py:40-45: upsample the /8 flow, add the
    base grid, clamp, and sample the coarse grid with it.
    (quick_start/align2images.py:91-95 is the same without the clamp.)"""
    h, w = grid.shape[1], grid.shape[2]
    flowUp = interpolate_bilinear(flowDown8, (h, w)).permute(0, 2, 3, 1)
    flowUp = flowUp + grid
    if clamp:
        flowUp = torch.clamp(flowUp, min=-1, max=1)
    flow12 = grid_sample(flowCoarse.permute(0, 3, 1, 2), flowUp).permute(0, 2, 3, 1).contiguous()
    return flow12, flowUp


def get_flow_all(flow, param, match, h, w, th=0.95, multiH=True, with_match21=False):
    """evaluation/evalHpatch/getResults.py:16-63 ``getFlow_all`` after the np.load calls:
    flow (nH,2,h8,w8), param (nH,3,3), match (nH,2,h8,w8) -> flowGlobal (1,h,w,2)
    (``with_match21``: the evalCorr/evalYFCC variant, evalCorr/getResults.py:78-136)."""
    flow = torch.as_tensor(flow, dtype=torch.float32)
    match = torch.as_tensor(match, dtype=torch.float32)
    grid = base_grid(h, w)
    coarse = warp_grid(param, h, w)
    f = interpolate_bilinear(flow, (h, w)).permute(0, 2, 3, 1)
    flowUp = torch.clamp(f + grid, min=-1, max=1)
    f = grid_sample(coarse.permute(0, 3, 1, 2), flowUp).permute(0, 2, 3, 1).contiguous()
    m = interpolate_bilinear(match, (h, w))
    m12 = m.narrow(1, 0, 1)
    if with_match21:
        m12 = m12 * grid_sample(m.narrow(1, 1, 1), flowUp)
    m12 = (m12 * inside_mask(f)).permute(0, 2, 3, 1)
    f = torch.clamp(f, min=-1, max=1)
    flowGlobal = f[:1].clone()
    if multiH:
        mb = m12[:1] >= th
        for i in range(1, len(m12)):
            tmp = (m12.narrow(0, i, 1) >= th) * (~mb)
            mb = mb + tmp
            tmp = tmp.expand_as(flowGlobal)
            flowGlobal[tmp] = f.narrow(0, i, 1)[tmp]
    return flowGlobal, m12


def get_flow_corr(flow, param, match, th=0.95, multiH=True):
    """evaluation/evalCorr/getResults.py:78-134 ``getFlow`` after the np.load calls (evalYFCC/getResults.py:150-190 is the
    same): flow (nH,2,h8,w8), param (nH,3,3), match (nH,2,h8,w8) -> (flowGlobal (1,8h8,8w8,2), matchGlobal (1,8h8,8w8,1)):
    x8 upsampling, ``match12 * grid_sample(match21) * inside``, first-hypothesis-wins merge of flow AND matchability."""
    flow = torch.as_tensor(flow, dtype=torch.float32)
    match = torch.as_tensor(match, dtype=torch.float32)
    h, w = flow.shape[2] * 8, flow.shape[3] * 8
    grid = base_grid(h, w)
    coarse = warp_grid(param, h, w)
    f = F.interpolate(flow, scale_factor=8, mode="bilinear").permute(0, 2, 3, 1)
    flowUp = torch.clamp(f + grid, min=-1, max=1)
    f = grid_sample(coarse.permute(0, 3, 1, 2), flowUp).permute(0, 2, 3, 1).contiguous()
    m = F.interpolate(match, scale_factor=8, mode="bilinear")
    m = (m.narrow(1, 0, 1) * grid_sample(m.narrow(1, 1, 1), flowUp) * inside_mask(f)).permute(0, 2, 3, 1)
    f = torch.clamp(f, min=-1, max=1)
    flowGlobal, matchGlobal = f[:1].clone(), m[:1].clone()
    mb = m[:1] >= th
    if multiH:
        for i in range(1, len(m)):
            tmp = (m.narrow(0, i, 1) >= th) * (~mb)
            matchGlobal[tmp] = m.narrow(0, i, 1)[tmp]
            mb = mb + tmp
            tmp = tmp.expand_as(flowGlobal)
            flowGlobal[tmp] = f.narrow(0, i, 1)[tmp]
    return flowGlobal, matchGlobal


# --------------------------------------------------------------------------------------------------
# KITTI extras (evaluation/evalKITTI)
# --------------------------------------------------------------------------------------------------
def remove_small_cc(match, match_th, cc_th):
    """evaluation/evalKITTI/evaluation.py:85-100 on one (H, W) float array (a copy is returned).
    ``skimage.measure.label(binary, background=0)`` labels 8-connected components of a 2-D array (default
    connectivity = ndim); skimage is not installed here, ``scipy.ndimage.label`` with a full 3x3 structure gives the
    same components (their numbering is irrelevant: only each component's area fraction is used)."""
    import scipy.ndimage as nd
    match = np.array(match, dtype=np.float32, copy=True)
    if cc_th == 0:
        return match
    labels, n = nd.label(match > match_th, structure=np.ones((3, 3), dtype=np.int32))
    if n == 0:                                                   # len(np.unique(all_labels)) == 1
        return match
    for i in range(1, n + 1):
        comp = labels == i
        if np.mean(comp) <= cc_th:
            match[comp] = 0
    return match


def interpolate_flow_match(flowGlobal, match_binary):
    """evaluation/evalKITTI/getResults.py:87-93: fill the unmatched pixels with the flow of the nearest matched pixel
    (scipy's exact EDT with indices, as the reference)."""
    import scipy.ndimage as nd
    mb = (~match_binary).squeeze().numpy()
    idx = nd.distance_transform_edt(mb, return_distances=False, return_indices=True)
    f = flowGlobal.squeeze().numpy()
    return torch.from_numpy(f[tuple(idx)]).unsqueeze(0)


def get_flow_all_kitti(param, flowd2, flow, match, h, w, th=1.0, cc_th=0.01, multiH=True, interpolate=False):
    """evaluation/evalKITTI/getResults.py:95-141 ``getFlow_all`` after the np.load calls: param (nH,3,3), flowd2
    (nH,2,hd,wd) the /8 flow of the half-size level, flow (nH,2,h8,w8) the /8 flow of the second level, match
    (nH,2,h8,w8) -> flowGlobal (1,h,w,2) (and the binary match map)."""
    param = torch.as_tensor(param, dtype=torch.float32)
    flowd2 = torch.as_tensor(flowd2, dtype=torch.float32)
    flow = torch.as_tensor(flow, dtype=torch.float32)
    match = torch.as_tensor(match, dtype=torch.float32)
    grid = base_grid(h, w)
    homography_org = warp_grid(param, h, w)
    fd2 = interpolate_bilinear(flowd2, (h, w)).permute(0, 2, 3, 1)
    fd2 = torch.clamp(fd2 + grid, min=-1, max=1)
    fd2 = grid_sample(homography_org.permute(0, 3, 1, 2), fd2).permute(0, 2, 3, 1).contiguous()
    f = interpolate_bilinear(flow, (h, w)).permute(0, 2, 3, 1)
    flowUp = torch.clamp(f + grid, min=-1, max=1)
    f = grid_sample(fd2.permute(0, 3, 1, 2), flowUp).permute(0, 2, 3, 1).contiguous()
    m = interpolate_bilinear(match, (h, w))
    m = m.narrow(1, 0, 1) * grid_sample(m.narrow(1, 1, 1), flowUp) * inside_mask(f)
    m = torch.from_numpy(np.stack([remove_small_cc(m[j, 0].numpy(), 0.99, cc_th) for j in range(m.shape[0])]))[:, None]
    m = m.permute(0, 2, 3, 1)
    f = torch.clamp(f, min=-1, max=1)
    flowGlobal = f[:1].clone()
    mb = m[:1] >= th
    if multiH:
        for i in range(1, len(m)):
            tmp = (m.narrow(0, i, 1) >= th) * (~mb)
            mb = mb + tmp
            tmp = tmp.expand_as(flowGlobal)
            flowGlobal[tmp] = f.narrow(0, i, 1)[tmp]
    if interpolate:
        flowGlobal = interpolate_flow_match(flowGlobal, mb)
    return flowGlobal, mb
