"""Host logic of the weight images (CPU, no kernels run): the stage-1 training blobs (fused_nerf_train.build_pack_index /
dw_views against the sizes the library reports) and the product layers of the tcgen05 image (fused_train.tc_merge_layers /
tc_layer_offset against pnerf_palette_tc_weight_bytes and the products formed directly)."""
import torch

from palettenerf_b200 import _lib as L
from palettenerf_b200 import fused_nerf_train as NT
from palettenerf_b200 import fused_train as FT
from palettenerf_b200 import synthetic as S


def test_stage1_training_blobs_match_the_library_layout():
    m = S.build_nerf_model("cpu", seed=3)
    sd = dict(m.named_parameters())
    index, n_fwd = NT.build_pack_index({n: sd[n].shape for n in NT.WEIGHT_NAMES})
    assert n_fwd == 4 * L.lib.pnerf_nerf_train_wfwd_units()
    assert index.numel() - n_fwd == 4 * L.lib.pnerf_nerf_train_wbwd_units()
    flat = torch.cat([sd[n].detach().reshape(-1) for n in NT.WEIGHT_NAMES] + [torch.zeros(1)])
    zero = flat.numel() - 1
    assert int(index.max()) == zero and int(index.min()) == 0
    # every real weight of the forward layers appears exactly once in the forward blob (the rest is padding)
    fwd = index[:n_fwd]
    real = fwd[fwd != zero]
    assert real.numel() == zero and torch.equal(torch.sort(real).values, torch.arange(zero))
    # the transposed blob holds V2, V1, the geo half of V0, S1 and S0
    bwd = index[n_fwd:]
    realb = bwd[bwd != zero]
    n_v0_geo = 64 * 15
    want = sum(sd[n].numel() for n in NT.WEIGHT_NAMES) - sd["color_net.0.weight"].numel() + n_v0_geo
    assert realb.numel() == want and realb.unique().numel() == want
    # packed weight-gradient buffer -> parameter shapes; color_net.0 skips the column that faces the sigma logit
    dw = torch.arange(int(L.lib.pnerf_nerf_train_dw_floats()), dtype=torch.float32)
    g = NT.dw_views(dw)
    assert set(g) == set(NT.WEIGHT_NAMES)
    for n in NT.WEIGHT_NAMES:
        assert tuple(g[n].shape) == tuple(sd[n].shape), n
    v0 = dw[3072:3072 + 2048].view(64, 32)
    assert torch.equal(g["color_net.0.weight"], torch.cat([v0[:, :16], v0[:, 17:]], 1))


def test_tcgen05_image_product_layers_and_offsets():
    for pred_clip in (False, True):
        m = S.build_palette_model("cpu", seed=5, pred_clip=pred_clip)
        sd = dict(m.named_parameters())
        names = FT.WEIGHT_NAMES + (FT.CLIP_NAMES if pred_clip else [])
        index = FT.tc_pack_index({n: sd[n].shape for n in names}, pred_clip, m.opt.clip_dim)
        assert 2 * index.numel() == L.lib.pnerf_palette_tc_weight_bytes(int(pred_clip))
        flat = torch.cat([sd[n].detach().reshape(-1) for n in names] + [torch.zeros(1)])
        img = flat[index].clone()
        # the gather leaves the product layers (and the unused pad layer) zero
        for name, (n, k) in FT.TC_MERGED_SHAPES.items():
            off = FT.tc_layer_offset(name)
            assert img[off:off + n * k].abs().max().item() == 0.0, name
        FT.tc_merge_layers(m, img)
        d0 = sd["diff_net.0.weight"].detach() @ sd["sigma_net.1.weight"].detach()[1:16]
        off = FT.tc_layer_offset("d0")
        got = img[off:off + 64 * 64].reshape(8, 64, 8).permute(1, 0, 2).reshape(64, 64)      # [k-chunk][n][8] -> [n][k]
        assert torch.allclose(got, d0, atol=1e-6)
        heads = torch.cat([sd["offsets_radiance_net.weight"].detach(), sd["omega_net.0.weight"].detach()]) @ sd["basis_net.1.weight"].detach()
        off = FT.tc_layer_offset("b1")
        got = img[off:off + 32 * 64].reshape(8, 32, 8).permute(1, 0, 2).reshape(32, 64)
        assert torch.allclose(got[:17], heads, atol=1e-6) and got[17:].abs().max().item() == 0.0
        # layers behind the product layers keep their place: the first layer of the image is sigma_net.0 in [k-chunk][n][8] order
        s0 = img[:64 * 32].reshape(4, 64, 8).permute(1, 0, 2).reshape(64, 32)
        assert torch.equal(s0, sd["sigma_net.0.weight"].detach())
