"""Host logic of the layer programs (program.py): the buffer-slot assignment that lets one C call run a whole network.
Compiled on the CPU (no kernel is launched): every layer's destination slot must be distinct from every tensor that is still
live (its own source / residual included), slots must be large enough, and the fp16 engine's element sizes must follow the
hand-over rules (fp32 only for the image, OUT_F32 outputs and TF32 layers)."""
import numpy as np
import pytest
import torch

from oracle import synth


def _programs(rf):
    from ransac_flow_b200.coarseAlignFeatMatch import ResNet50Conv4
    net = ResNet50Conv4(synth.resnet50_conv4_state(0), device="cpu")
    yield "resnet50 fp32", net.program, False, 3
    yield "resnet50 f16", net._build(64), True, 3
    yield "resnet50 split (conv3 + down-sampling fused)", net._build(64, fuse_downsample=True), "split", 3
    fe = rf.model.FeatureExtractor()
    fe.load_state_dict(synth.feature_extractor_state(0))
    fe.eval()
    yield "FeatureExtractor fp32", fe._fold_build(False), False, 3
    yield "FeatureExtractor f16", fe._fold_build(True), True, 3
    nf = rf.model.NetFlowCoarse(7)
    nf.load_state_dict(synth.net_flow_coarse_state(1))
    yield "NetFlowCoarse fp32", nf._fold_build(False), False, 64
    yield "NetFlowCoarse f16", nf._fold_build(True), True, 64
    nm = rf.model.NetMatchability(7)
    nm.load_state_dict(synth.net_matchability_state(2))
    yield "NetMatchability f16", nm._fold_build(True), True, 64


@pytest.mark.parametrize("hw", [[(96, 128)], [(192, 256), (96, 128), (48, 64), (96, 128)], [(33, 47), (480, 640)]])
def test_slot_assignment_never_aliases_live_tensors(rf, hw):
    for name, P, f16, cin in _programs(rf):
        if "Net" in name:
            hw_in = [(h // 8, w // 8) for h, w in hw]
        else:
            hw_in = hw
        split = f16 == "split"
        f16 = f16 is True
        c = P._compile(hw_in, torch.device("cpu"), f16, split)
        n_ops = len(P.ops)
        layers = c["layers"]
        # symbolic tensor -> slot, replayed in execution order with liveness from the topology
        last_use = {}
        for i, o in enumerate(P.ops):
            last_use[o[1]] = i
            if o[2] >= 0:
                last_use[o[2]] = i
            if i in P.dual:
                last_use[P.dual[i][0]] = i                       # the second input of a dual 1x1
        last_use[n_ops] = n_ops                                  # the output outlives the program
        slot_of = {0: 0}
        for i, o in enumerate(P.ops):
            L = layers[i]
            assert L.src == slot_of[o[1]] and (L.res == -1) == (o[2] < 0), name
            if o[2] >= 0:
                assert L.res == slot_of[o[2]], name
            if i in P.dual:
                assert (L.op, L.src2, L.Cin2, L.stride2) == (6, slot_of[P.dual[i][0]], P.chan[P.dual[i][0]], P.dual[i][2]), name
            else:
                assert L.src2 == -1, name
            live = {t for t, s in slot_of.items() if last_use.get(t, -1) >= i}
            for t in live:                                       # the destination must not overwrite anything still needed
                assert slot_of[t] != L.dst, (name, i, t)
            assert L.dst != 0, name                              # slot 0 is the caller's input
            slot_of[i + 1] = L.dst
        assert c["out_slot"] == slot_of[n_ops] and c["nslots"] <= 32
        # every slot holds its largest tenant
        for t, s in slot_of.items():
            if s == 0:
                continue
            px = sum(h * w for h, w in _hw_of(P, hw_in, t))
            esz = _esize(P, t, f16)
            assert c["bufs"][s].numel() >= px * P.chan[t] * esz, (name, t)
        assert c["out_dtype"] == (torch.float16 if ((f16 or split) and "Net" not in name) else torch.float32), name
        assert c["in_dtype"] == (torch.float16 if (f16 and "Net" in name) else torch.float32), name
        if split:                                                # three fused blocks: 3 layers fewer than the plain topology
            assert sum(1 for o in P.ops if o[0] == 6) == 3 and n_ops == 41


def _hw_of(P, hw, t):
    hws = [list(hw)]
    for o in P.ops:
        k, s, p = o[5], o[6], o[7]
        hws.append([((h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1) for h, w in hws[o[1]]])
    return hws[t]


def _esize(P, t, f16):
    if not f16:
        return 4
    if t == 0:
        return 4 if P.ops[0][0] in (3, 5) else 2                 # RF_OP_IM2COL / RF_OP_STEM7 read the fp32 image
    return 4 if P.flags.get(t - 1, 0) else 2


def test_fp16_handover_flags(rf):
    """Heads under the fp16 engine: conv3 writes fp32 (OUT_F32) for the TF32 conv4; everything before is fp16."""
    nf = rf.model.NetFlowCoarse(7)
    P = nf._fold_build(True)
    assert [P.flags.get(i, 0) for i in range(len(P.ops))] == [0, 0, 1, 2]
    assert [o[4] for o in P.ops] == [512, 256, 128, 49] and P.chan[0] == 64
    fe = rf.model.FeatureExtractor()
    Pf = fe._fold_build(True)
    assert all(Pf.flags.get(i, 0) == 0 for i in range(len(Pf.ops))) and Pf.chan[-1] == 256
    assert np.prod([o[6] for o in Pf.ops if o[0] in (0, 4) and o[6] > 1]) == 8      # total stride of the main path


def test_concat_k_folds_conv3_and_downsample_into_one_weight_matrix(rf):
    """FoldedConv.concat_k (host logic of the fused bottleneck tail): [W3 | Wd] with the summed folded-BN biases reproduces
    bn3(conv3(h)) + bn_d(conv_d(x[::s])) - checked here with plain torch on the CPU."""
    import torch.nn.functional as F
    from ransac_flow_b200.coarseAlignFeatMatch import _BN
    sd = synth.resnet50_conv4_state(0)
    p = "layer2.0"
    c3 = rf.model.FoldedConv(sd[p + ".conv3.weight"], _BN(sd, p + ".bn3"), 1, pad=0, device="cpu")
    ds = rf.model.FoldedConv(sd[p + ".downsample.0.weight"], _BN(sd, p + ".downsample.1"), 2, pad=0, device="cpu")
    f = rf.model.FoldedConv.concat_k(c3, ds)
    assert (f.cin, f.cin2, f.cout, f.k, f.pad) == (128, 256, 512, 1, 0) and f.w_split.shape == (2, 512, 384)
    g = torch.Generator().manual_seed(0)
    h, x = torch.randn(1, 128, 5, 7, generator=g), torch.randn(1, 256, 10, 13, generator=g)

    def bn(y, name):
        b = _BN(sd, name)
        return F.batch_norm(y, b.running_mean, b.running_var, b.weight, b.bias, False, 0.0, b.eps)

    ref = bn(F.conv2d(h, sd[p + ".conv3.weight"]), p + ".bn3") + bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=2), p + ".downsample.1")
    cat = torch.cat([h, x[:, :, ::2, ::2]], dim=1)                       # the kernel's K axis: conv2's output | the sampled block input
    got = torch.einsum("ok,nkhw->nohw", f._wt, cat) + f.bias.view(1, -1, 1, 1)
    assert (got - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    # the split planes carry the same matrix to 22 bits
    w22 = f.w_split[0].float() + f.w_split[1].float() / 2048.0
    assert (w22 - f._wt).abs().max().item() <= 2.0 ** -21 * f._wt.abs().max().item()
