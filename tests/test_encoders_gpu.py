"""GPU parity of the hash-grid / SH / frequency encoders and the HSV ops through the C ABI, vs the CPU oracle
(fp64 restatement) and vs the reference's own extensions when oracle/_ref is present.

Tolerances (written per check): fp32 tables 1e-5 (north_star), fp16 tables 1e-3 max-abs on O(1) outputs. Tables are
drawn from U(-1,1), not the reference's U(-1e-4,1e-4) init, so the checks are not vacuous."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_ref
from palettenerf_b200.gridencoder import GridEncoder, grid_encode
from palettenerf_b200.gridencoder.backend import _backend as GB
from palettenerf_b200.shencoder import SHEncoder
from palettenerf_b200.shencoder.backend import _backend as SB
from palettenerf_b200.freqencoder import FreqEncoder
from palettenerf_b200.palette.backend import rgb_to_hsv, hsv_to_rgb

pytestmark = pytest.mark.gpu


def _device_exp2(L, S, cuda):
    """exp2f(level * S) as CUDA's math library rounds it (see oracle._level_setup)"""
    return torch.exp2(torch.arange(L, device=cuda, dtype=torch.float32) * float(np.float32(S))).cpu().numpy()


def _points(B, D, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, D, generator=g)
    x[0] = 0.0; x[1] = 1.0                 # the closed ends of [0,1]
    x[2, 0] = -0.01; x[3, D - 1] = 1.01    # out of range -> zero rows, no gradient
    return x


@pytest.mark.parametrize("cfg", [
    dict(D=3, C=2, L=16, H=16, log2T=19, res=4096, gridtype="hash", dtype=torch.float32),
    dict(D=3, C=2, L=16, H=16, log2T=19, res=4096, gridtype="hash", dtype=torch.float16),
    dict(D=3, C=2, L=16, H=16, log2T=15, res=2048, gridtype="tiled", dtype=torch.float32),
    dict(D=3, C=2, L=6, H=16, log2T=19, res=512, gridtype="hash", dtype=torch.float16, align=True),
    dict(D=2, C=2, L=4, H=16, log2T=19, res=2048, gridtype="hash", dtype=torch.float32),   # bg encoder shape
    dict(D=3, C=4, L=8, H=16, log2T=14, res=1024, gridtype="hash", dtype=torch.float32),
    dict(D=3, C=1, L=5, H=8, log2T=12, res=256, gridtype="hash", dtype=torch.float32),
    dict(D=3, C=8, L=3, H=4, log2T=10, res=64, gridtype="hash", dtype=torch.float16),
])
def test_grid_encode_forward_backward(cuda, cfg):
    ref = load_ref("gridencoder")
    D, C, L, dt = cfg["D"], cfg["C"], cfg["L"], cfg["dtype"]
    align = cfg.get("align", False)
    enc = GridEncoder(input_dim=D, num_levels=L, level_dim=C, base_resolution=cfg["H"], log2_hashmap_size=cfg["log2T"],
                      desired_resolution=cfg["res"], gridtype=cfg["gridtype"], align_corners=align)
    g = torch.Generator().manual_seed(5)
    emb = (torch.rand(enc.embeddings.shape, generator=g) * 2 - 1).to(dt)
    B = 20011
    x = _points(B, D, 1)
    offsets = enc.offsets
    S = np.log2(enc.per_level_scale)

    e2 = _device_exp2(L, S, cuda)
    emb_c = emb.to(cuda).requires_grad_(True)
    out = grid_encode(x.to(cuda), emb_c, offsets.to(cuda), enc.per_level_scale, cfg["H"], False, enc.gridtype_id, align)
    assert out.shape == (B, L * C) and out.dtype == dt
    oout = oracle.grid_encode_forward(x.numpy(), emb.float().numpy(), offsets.numpy(), np.float32(S), cfg["H"],
                                      enc.gridtype_id, align, exp2_levels=e2)
    tol = 1e-3 if dt == torch.float16 else 1e-5
    err = np.abs(out.detach().float().cpu().numpy() - oout)
    # a point within 1 ulp of a cell boundary may pick the neighbouring cell with weight ~1e-7: harmless
    assert err.max() < tol, f"max-abs {err.max()}"
    assert np.abs(oout).max() > 0.3            # not vacuous
    assert not out[2].any() and not out[3].any()

    grad = torch.randn(B, L * C, generator=g).to(dt)
    out.backward(grad.to(cuda))
    og = oracle.grid_encode_backward(grad.float().numpy(), x.numpy(), emb.shape[0], offsets.numpy(), np.float32(S), cfg["H"],
                                     enc.gridtype_id, align, exp2_levels=e2)
    got = emb_c.grad.float().cpu().numpy()
    scale = np.abs(og).max()
    if dt == torch.float16:
        # fp16 atomics: each add rounds at the running sum's magnitude (2^-11 relative), order-dependent
        assert np.abs(got - og).max() < 2e-2 * scale
        assert np.abs(got - og).mean() < 1e-3 * scale
    else:
        assert np.abs(got - og).max() < 1e-5 * scale + 1e-6

    if ref is not None:
        rout = torch.empty(L, B, C, device=cuda, dtype=dt)
        ref.grid_encode_forward(x.to(cuda), emb.to(cuda), offsets.to(cuda), rout, B, D, C, L, float(S), cfg["H"], None,
                                enc.gridtype_id, align)
        rout = rout.permute(1, 0, 2).reshape(B, L * C)
        # the reference accumulates the 2^D corners in the table dtype (fp16: 8 roundings); ours once from fp32
        rtol_ref = 4e-3 if dt == torch.float16 else 1e-6
        assert (out.detach().float() - rout.float()).abs().max().item() < rtol_ref
        # reference-layout entry point ([L,B,C]) of the new kernel
        mine = torch.empty(L, B, C, device=cuda, dtype=dt)
        GB.grid_encode_forward(x.to(cuda), emb.to(cuda), offsets.to(cuda), mine, B, D, C, L, float(S), cfg["H"], None,
                               enc.gridtype_id, align)
        assert torch.equal(mine.permute(1, 0, 2).reshape(B, L * C), out.detach())
        rg = torch.zeros_like(emb, device=cuda)
        ref.grid_encode_backward(grad.view(B, L, C).permute(1, 0, 2).contiguous().to(cuda), x.to(cuda), emb.to(cuda),
                                 offsets.to(cuda), rg, B, D, C, L, float(S), cfg["H"], None, None, enc.gridtype_id, align)
        diff = (emb_c.grad.float() - rg.float()).abs().max().item()
        assert diff < (4e-2 if dt == torch.float16 else 1e-5) * scale + 1e-6


def test_grid_encode_input_gradient_and_float64(cuda):
    """dy_dx path (inputs.requires_grad) in fp64, the configuration of the reference's testing/test_hashgrid_grad.py:
    D=3, L=4, F=2, base 4, log2T=8; checked against the oracle's analytic dy_dx and by finite differences."""
    enc = GridEncoder(input_dim=3, num_levels=4, level_dim=2, base_resolution=4, log2_hashmap_size=8, per_level_scale=2)
    g = torch.Generator().manual_seed(0)
    emb = (torch.rand(enc.embeddings.shape, generator=g, dtype=torch.float64) * 2 - 1)
    x = torch.rand(64, 3, generator=g) * 0.9 + 0.05
    xc = x.to(cuda).requires_grad_(True)
    emb_c = emb.to(cuda).requires_grad_(True)
    out = grid_encode(xc, emb_c, enc.offsets.to(cuda), 2.0, 4, True, 0, False)
    assert out.dtype == torch.float64
    oout, ody = oracle.grid_encode_forward(x.numpy(), emb.numpy(), enc.offsets.numpy(), np.float32(1.0), 4, 0, False, True)
    np.testing.assert_allclose(out.detach().cpu().numpy(), oout, atol=1e-6)
    grad = torch.randn(64, 8, generator=g, dtype=torch.float64)
    out.backward(grad.to(cuda))
    ogx = np.einsum("blc,bldc->bd", grad.numpy().reshape(64, 4, 2), ody)
    np.testing.assert_allclose(xc.grad.cpu().numpy(), ogx, atol=1e-5)
    og = oracle.grid_encode_backward(grad.numpy(), x.numpy(), emb.shape[0], enc.offsets.numpy(), np.float32(1.0), 4, 0, False)
    np.testing.assert_allclose(emb_c.grad.cpu().numpy(), og, atol=1e-9)


def test_grid_module_amp_contract(cuda):
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, desired_resolution=4096).to(cuda)
    enc.embeddings.data.uniform_(-1, 1)
    x = (torch.rand(4096, 3, device=cuda) * 2 - 1) * 2
    with torch.autocast("cuda", dtype=torch.float16):
        y = enc(x, bound=2)
    assert y.dtype == torch.float16 and y.shape == (4096, 32)
    y.float().sum().backward()
    assert enc.embeddings.grad is not None and enc.embeddings.grad.dtype == torch.float32
    y32 = enc(x, bound=2)
    assert y32.dtype == torch.float32
    assert (y.float() - y32).abs().max().item() < 2e-3


def test_grid_unsupported_shapes_raise(cuda):
    with pytest.raises(RuntimeError, match="unsupported"):
        GB.grid_encode_forward(torch.zeros(4, 6, device=cuda), torch.zeros(64, 2, device=cuda),
                               torch.tensor([0, 64], dtype=torch.int32, device=cuda), torch.zeros(1, 4, 2, device=cuda), 4, 6, 2,
                               1, 1.0, 4, None, 0, False)
    with pytest.raises(RuntimeError, match="unsupported"):
        GB.grid_encode_forward(torch.zeros(4, 3, device=cuda), torch.zeros(64, 3, device=cuda),
                               torch.tensor([0, 64], dtype=torch.int32, device=cuda), torch.zeros(1, 4, 3, device=cuda), 4, 3, 3,
                               1, 1.0, 4, None, 0, False)


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7, 8])
def test_sh_encoder(cuda, degree):
    ref = load_ref("shencoder")
    g = torch.Generator().manual_seed(degree)
    d = torch.randn(5000, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    d[:100] *= 0.7   # the kernels use inputs as given (no normalisation)
    enc = SHEncoder(degree=degree)
    dc = d.to(cuda).requires_grad_(True)
    out = enc(dc)
    o, og = oracle.sh_encode(d.numpy(), degree, with_grad=True)
    # fp32 Horner evaluation of degree-7 polynomials with coefficients up to ~40
    np.testing.assert_allclose(out.detach().cpu().numpy(), o, rtol=1e-5, atol=2e-5)
    grad = torch.randn(5000, degree ** 2, generator=g)
    out.backward(grad.to(cuda))
    ogx = np.einsum("bc,bdc->bd", grad.double().numpy(), og)
    np.testing.assert_allclose(dc.grad.cpu().numpy(), ogx, rtol=1e-4, atol=1e-4 * max(1.0, np.abs(ogx).max()))
    if ref is not None:
        ro = torch.empty(5000, degree ** 2, device=cuda); rdy = torch.empty(5000, 3 * degree ** 2, device=cuda)
        ref.sh_encode_forward(d.to(cuda), ro, 5000, 3, degree, rdy)
        np.testing.assert_allclose(out.detach().cpu().numpy(), ro.cpu().numpy(), rtol=1e-5, atol=2e-5)
        mdy = torch.empty(5000, 3 * degree ** 2, device=cuda); mo = torch.empty(5000, degree ** 2, device=cuda)
        SB.sh_encode_forward(d.to(cuda), mo, 5000, 3, degree, mdy)
        np.testing.assert_allclose(mdy.cpu().numpy(), rdy.cpu().numpy(), rtol=1e-4, atol=1e-3)


def test_freq_encoder(cuda):
    ref = load_ref("freqencoder")
    g = torch.Generator().manual_seed(0)
    x = torch.rand(3000, 3, generator=g) * 2 - 1
    enc = FreqEncoder(3, 6)
    xc = x.to(cuda).requires_grad_(True)
    out = enc(xc)
    o = oracle.freq_encode(x.numpy(), 6)
    # __sinf: absolute error ~2^-21.4 for |arg| <= pi, growing with |arg| (up to 32+pi/2 here)
    np.testing.assert_allclose(out.detach().cpu().numpy(), o, atol=2e-5)
    grad = torch.randn(3000, 39, generator=g)
    out.backward(grad.to(cuda))
    xd = x.double().requires_grad_(True)
    od = torch.cat([xd] + [f(xd * 2.0 ** k) for k in range(6) for f in (torch.sin, torch.cos)], -1)
    od.backward(grad.double())
    np.testing.assert_allclose(xc.grad.cpu().numpy(), xd.grad.numpy(), rtol=1e-3, atol=2e-3)
    if ref is not None:
        ro = torch.empty(3000, 39, device=cuda)
        ref.freq_encode_forward(x.to(cuda), 3000, 3, 6, 39, ro)
        np.testing.assert_allclose(out.detach().cpu().numpy(), ro.cpu().numpy(), atol=1e-6)


def test_rgb_hsv(cuda):
    ref = load_ref("palette_func")
    g = torch.Generator().manual_seed(0)
    rgb = torch.rand(10000, 3, generator=g)
    rgb[0] = torch.tensor([0.5, 0.5, 0.5]); rgb[1] = 0.0; rgb[2] = torch.tensor([1.0, 0.0, 0.0]); rgb[3] = torch.tensor([0.2, 0.9, 0.9])
    hsv = rgb_to_hsv(rgb.to(cuda).view(100, 100, 3))
    assert hsv.shape == (100, 100, 3)
    o = oracle.rgb_to_hsv(rgb.numpy())
    np.testing.assert_allclose(hsv.view(-1, 3).cpu().numpy(), o, rtol=1e-5, atol=1e-3)   # H in degrees, S,V in [0,100]
    back = hsv_to_rgb(hsv)
    np.testing.assert_allclose(back.view(-1, 3).cpu().numpy(), rgb.numpy(), atol=2e-5)    # round trip
    np.testing.assert_allclose(back.view(-1, 3).cpu().numpy(), oracle.hsv_to_rgb(o), atol=2e-5)
    if ref is not None:
        rh = torch.empty(10000, 3, device=cuda)
        ref.rgb_to_hsv(10000, rgb.to(cuda), rh)
        np.testing.assert_allclose(hsv.view(-1, 3).cpu().numpy(), rh.cpu().numpy(), rtol=1e-4, atol=2e-2)  # ref: -use_fast_math
        rb = torch.empty(10000, 3, device=cuda)
        ref.hsv_to_rgb(10000, rh, rb)
        np.testing.assert_allclose(back.view(-1, 3).cpu().numpy(), rb.cpu().numpy(), atol=1e-4)
