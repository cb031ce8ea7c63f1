"""Oracle (test infrastructure) for the networks on the hot path.

torch-CPU fp32 functional restatement of ``model/model.py`` (FeatureExtractor,
CorrNeigh, NetFlowCoarse, NetMatchability), ``model/downsample.py`` and the
torchvision ResNet-50 conv1..layer3 trunk that ``CoarseAlign`` builds
(quick_start/coarseAlignFeatMatch.py:34-52).  All functions take plain
``state_dict``s with the reference's key names.
"""
import torch
import torch.nn.functional as F

EPS = 1e-5


def _bn(x, sd, p, eps=EPS):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, eps)


def blur_downsample(x, stride=2):
    """model/downsample.py:12-46: reflect-pad 1, depthwise [1 2 1]x[1 2 1]/16, stride."""
    c = x.shape[1]
    a = torch.tensor([1.0, 2.0, 1.0])
    filt = a[:, None] * a[None, :]
    filt = (filt / filt.sum())[None, None].repeat(c, 1, 1, 1)
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), filt, stride=stride, groups=c)


def basic_block(x, sd, p, stride):
    """model/model.py:27-56 (BasicBlock) with the anti-aliased shortcut of
    FeatureExtractor._make_layer (model/model.py:89-103)."""
    out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], stride=stride, padding=1), sd, p + ".bn1"))
    out = _bn(F.conv2d(out, sd[p + ".conv2.weight"], padding=1), sd, p + ".bn2")
    if (p + ".downsample.1.weight") in sd:            # keys: downsample.0.filt, .1.weight (1x1), .2.* (BN)
        r = blur_downsample(x, stride)
        r = _bn(F.conv2d(r, sd[p + ".downsample.1.weight"]), sd, p + ".downsample.2")
    else:
        r = x
    return F.relu(out + r)


def feature_extractor(x, sd):
    """model/model.py:59-125 FeatureExtractor.do_forward: (1,3,H,W) -> (1,256,H/8,W/8)."""
    with torch.no_grad():
        x = F.relu(_bn(F.conv2d(x, sd["conv1.weight"], padding=1), sd, "bn1"))
        x = F.max_pool2d(x, kernel_size=2, stride=1)
        x = blur_downsample(x, 2)
        for layer, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
            x = basic_block(x, sd, layer + ".0", stride)
            x = basic_block(x, sd, layer + ".1", 1)
        return x


def corr_neigh(x, y, k=7):
    """model/model.py:129-160 CorrNeigh: 7x7 local correlation, channel i*k+j
    correlates x(r,c) with zero-padded y(r+i-k//2, c+j-k//2)."""
    p = k // 2
    n, c, h, w = x.shape
    yp = F.pad(y, (p, p, p, p))
    out = []
    for i in range(k):
        for j in range(k):
            out.append((x * yp[:, :, i:i + h, j:j + w]).sum(dim=1, keepdim=True))
    return torch.cat(out, dim=1)


def _trunk(corr, sd):
    x = F.relu(_bn(F.conv2d(corr, sd["conv1.weight"], padding=1), sd, "bn1"))
    x = F.relu(_bn(F.conv2d(x, sd["conv2.weight"], padding=1), sd, "bn2"))
    x = F.relu(_bn(F.conv2d(x, sd["conv3.weight"], padding=1), sd, "bn3"))
    return F.conv2d(x, sd["conv4.weight"], padding=1)


def net_flow_coarse(corr, sd, k=7):
    """model/model.py:167-249 NetFlowCoarse.do_forward with up8X=False.

    flowX = sum_p p * (j - k//2) / size(3) * 2, flowY = sum_p p * (i - k//2) / size(2) * 2
    for channel i*k+j (model/model.py:190-191,228-232)."""
    with torch.no_grad():
        n, c, d2, d3 = corr.shape
        p = F.softmax(_trunk(corr, sd), dim=1)
        r = k // 2
        gy = torch.arange(-r, r + 1).view(1, 1, -1, 1).expand(1, 1, k, k).contiguous().view(1, -1, 1, 1).float()
        gx = torch.arange(-r, r + 1).view(1, 1, 1, -1).expand(1, 1, k, k).contiguous().view(1, -1, 1, 1).float()
        flowX = torch.sum(p * gx, dim=1, keepdim=True) / d3 * 2
        flowY = torch.sum(p * gy, dim=1, keepdim=True) / d2 * 2
        return torch.cat((flowX, flowY), dim=1)


def net_matchability(corr, sd):
    """model/model.py:254-322 NetMatchability.do_forward with up8X=False."""
    with torch.no_grad():
        return torch.sigmoid(_trunk(corr, sd))


# --------------------------------------------------------------------------
# torchvision ResNet-50 conv1..layer3 (quick_start/coarseAlignFeatMatch.py:34-52;
# MoCo variant model/resnet50.py:107-168 has the same trunk)
# --------------------------------------------------------------------------
RESNET50_LAYERS = (("layer1", 64, 3, 1), ("layer2", 128, 4, 2), ("layer3", 256, 6, 2))


def bottleneck(x, sd, p, stride):
    out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
    out = F.relu(_bn(F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1), sd, p + ".bn2"))
    out = _bn(F.conv2d(out, sd[p + ".conv3.weight"]), sd, p + ".bn3")
    if (p + ".downsample.0.weight") in sd:
        r = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
    else:
        r = x
    return F.relu(out + r)


def resnet50_conv4(x, sd):
    """(1,3,H,W) normalised image -> (1,1024,H/16,W/16), post-ReLU."""
    with torch.no_grad():
        x = F.relu(_bn(F.conv2d(x, sd["conv1.weight"], stride=2, padding=3), sd, "bn1"))
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
        for layer, planes, blocks, stride in RESNET50_LAYERS:
            for b in range(blocks):
                x = bottleneck(x, sd, "%s.%d" % (layer, b), stride if b == 0 else 1)
        return x
