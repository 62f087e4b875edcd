"""Golden-vector tests.

tests/golden/ref_python.npz   outputs of the reference's own Python (generated in the build container by
                              tests/golden/make_golden_cpu.py, which imports /root/reference in place)
tests/golden/ref_kernels.npz  outputs of the reference's own CUDA kernels on a B200 (tests/golden/make_golden_gpu.py,
                              which loads only oracle/_ref/*.so — the reference sources compiled as they are)

CPU suite (`-m "not gpu"`): the ORACLE is pinned against both files.
GPU suite (`-m gpu`): the CUDA path, called through the C-ABI bindings, is checked against the same fixtures.
Integer / index / coordinate outputs: bit-exact. Floating point: tolerance stated at each assert.
"""
import os

import numpy as np
import pytest

import oracle
from oracle import cpu_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    p = os.path.join(HERE, "golden", name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not generated yet")
    return np.load(p)


@pytest.fixture(scope="module")
def gp():
    return _load("ref_python.npz")


@pytest.fixture(scope="module")
def gk():
    return _load("ref_kernels.npz")


def table(n_entries, C):
    """hash-grid table of the fixtures (same integer formula as make_golden_gpu.table)"""
    i = np.arange(n_entries, dtype=np.uint64)[:, None]
    c = np.arange(C, dtype=np.uint64)[None, :]
    v = (i * np.uint64(2654435761) + c * np.uint64(40503) + np.uint64(12345)) % np.uint64(65536)
    return (v.astype(np.float64) / 32768.0 - 1.0).astype(np.float32)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def sorted_view(rays, arrs, M):
    r = rays[np.argsort(rays[:, 0], kind="stable")]
    outs = []
    for a in arrs:
        rows = [a[off:off + cnt] for _, off, cnt in r if cnt > 0 and off + cnt <= M]
        outs.append(np.concatenate(rows) if rows else np.zeros((0,) + a.shape[1:], a.dtype))
    return r, outs


def scatter_sorted(rays, M, *sorted_arrs):
    """inverse of sorted_view: place per-sample rows given in ray-id order into the layout described by `rays`"""
    outs = [np.zeros((M,) + a.shape[1:], a.dtype) for a in sorted_arrs]
    k = 0
    for _, off, c in rays[np.argsort(rays[:, 0], kind="stable")]:
        if c > 0 and off + c <= M:
            for o, a in zip(outs, sorted_arrs):
                o[off:off + c] = a[k:k + c]
            k += c
    return outs


# =====================================================================================================================
# CPU: oracle vs the reference's Python
# =====================================================================================================================
def test_oracle_freq_vs_reference_torch_encoder(gp):
    # reference class uses sin / cos, the CUDA kernel (and the oracle) sin(x + pi/2 in fp32): |diff| <= 2^f * 4.4e-8
    np.testing.assert_allclose(O.freq_encode(gp["freq_in"], 6), gp["freq_out_deg6"], atol=5e-6, rtol=0)


@pytest.mark.parametrize("deg", [1, 2, 3, 4, 5])
def test_oracle_sh_vs_reference_torch_formulae(gp, deg):
    # the reference's torch formulae use xx+yy+zz == 1 (test_shencoder.py:62); inputs are unit vectors rounded to fp32
    np.testing.assert_allclose(O.sh_encode(gp["sh_in"], deg), gp[f"sh_out_deg{deg}"], atol=2e-6, rtol=0)


def test_oracle_grid_offsets_vs_reference_module(gp):
    cfgs = [dict(input_dim=3, num_levels=16, base_resolution=16, log2_hashmap_size=19),
            dict(input_dim=3, num_levels=16, base_resolution=16, log2_hashmap_size=19),
            dict(input_dim=2, num_levels=8, base_resolution=8, log2_hashmap_size=14),
            dict(input_dim=3, num_levels=6, base_resolution=4, log2_hashmap_size=10, align_corners=True)]
    for i, c in enumerate(cfgs):
        got = O.grid_offsets(per_level_scale=float(gp[f"offsets_{i}_scale"]), **c)
        assert np.array_equal(got, gp[f"offsets_{i}"]), f"config {i}"
    assert int(gp["offsets_0"][-1]) == 6328848   # entries of the default table


def test_product_grid_module_offsets_vs_reference_module(gp):
    """host logic of the product (no kernel call): GridEncoder builds the same table layout and scale"""
    from palettenerf_b200.gridencoder import GridEncoder
    g = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=4096)
    assert np.array_equal(g.offsets.numpy(), gp["offsets_0"])
    assert g.per_level_scale == float(gp["offsets_0_scale"])
    assert tuple(g.embeddings.shape) == (6328848, 2)
    g = GridEncoder(input_dim=2, num_levels=8, level_dim=4, base_resolution=8, log2_hashmap_size=14, per_level_scale=1.7,
                    gridtype="tiled")
    assert np.array_equal(g.offsets.numpy(), gp["offsets_2"])


def test_product_trunc_exp_vs_reference(gp):
    import torch
    from palettenerf_b200.activation import trunc_exp
    v = torch.from_numpy(gp["trunc_exp_in"]).requires_grad_(True)
    y = trunc_exp(v)
    y.backward(torch.ones_like(y))
    assert np.array_equal(bits(y.detach().numpy()), bits(gp["trunc_exp_out"]))
    assert np.array_equal(bits(v.grad.numpy()), bits(gp["trunc_exp_grad"]))


@pytest.mark.parametrize("bpc", [3, 5])
def test_oracle_and_product_histogram_vs_reference_extension(gp, bpc):
    import ctypes
    from palettenerf_b200 import _lib as L
    bw, bc = O.compute_rgb_histogram(gp["hist_colors"], gp["hist_weights"], bpc)
    assert np.array_equal(bw, gp[f"hist_bin_weights_b{bpc}"])          # same fp64 accumulation order: exact
    assert np.array_equal(bc, gp[f"hist_bin_centers_b{bpc}"])
    # the product's host entry point (CPU pointers, no GPU needed)
    nb = 1 << (3 * bpc)
    c = np.ascontiguousarray(gp["hist_colors"], np.float32).reshape(-1)
    w = np.ascontiguousarray(gp["hist_weights"], np.float32)
    pw, pc = np.zeros(nb, np.float64), np.zeros((nb, 3), np.float32)
    st = L.lib.pnerf_compute_rgb_histogram(c.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p), w.shape[0],
                                           bpc, pw.ctypes.data_as(ctypes.c_void_p), pc.ctypes.data_as(ctypes.c_void_p))
    assert st == 0
    assert np.array_equal(pw, gp[f"hist_bin_weights_b{bpc}"]) and np.array_equal(pc, gp[f"hist_bin_centers_b{bpc}"])


# =====================================================================================================================
# CPU: oracle vs the reference's CUDA kernels
# =====================================================================================================================
def test_oracle_near_far_morton_packbits_vs_reference_kernels(gk):
    n, f = oracle.near_far_from_aabb(gk["rays_o"], gk["rays_d"], gk["aabb"], 0.2)
    assert np.array_equal(bits(n), bits(gk["nears"])) and np.array_equal(bits(f), bits(gk["fars"]))
    assert n[5] == np.finfo(np.float32).max                        # the ray that misses the box
    assert np.array_equal(oracle.morton3D(gk["morton_coords"]), gk["morton_idx"])
    assert np.array_equal(oracle.morton3D_invert(gk["morton_idx"]), gk["morton_back"])
    assert np.array_equal(gk["morton_back"], gk["morton_coords"])
    assert np.array_equal(oracle.packbits(gk["packbits_grid"], 0.5), gk["packbits_out"])


def test_oracle_scene_bitfield_vs_reference_packbits(gk, scene):
    """the synthetic scene generator + the oracle's packbits reproduce the bitfield the reference kernel packed"""
    assert np.float32(scene["thresh"]) == gk["scene_thresh"]
    assert np.array_equal(scene["bitfield"].numpy(), gk["scene_bitfield"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_march_rays_train_vs_reference_kernel(gk, tag):
    dt_gamma, max_steps, M = gk[f"march_{tag}_cfg"]
    max_steps, M = int(max_steps), int(M)
    x, d, dl, rays, cnt = oracle.march_rays_train(gk["rays_o"], gk["rays_d"], gk["scene_bitfield"], 2.0, float(dt_gamma),
                                                  max_steps, 2, 128, M, gk["nears"], gk["fars"], gk[f"march_{tag}_noises"])
    assert np.array_equal(cnt, gk[f"march_{tag}_counter"])
    r, (sx, sd, sl) = sorted_view(rays, [x, d, dl], M)
    assert np.array_equal(r[:, [0, 2]], gk[f"march_{tag}_rays_sorted_counts"])         # per-ray sample counts
    assert np.array_equal(bits(sx), bits(gk[f"march_{tag}_xyzs"]))                      # same fp32 bit patterns
    assert np.array_equal(bits(sl), bits(gk[f"march_{tag}_deltas"]))
    assert bool(gk[f"march_{tag}_dirs_ok"])
    assert int(cnt[0]) > 500


def _comp_layout(gk):
    """a deterministic layout for the composite fixtures: samples in ray-id order"""
    counts = gk["march_a_rays_sorted_counts"]
    rays = np.zeros((counts.shape[0], 3), np.int32)
    rays[:, 0], rays[:, 2] = counts[:, 0], counts[:, 1]
    rays[:, 1] = np.concatenate([[0], np.cumsum(counts[:, 1])[:-1]])
    return rays, int(counts[:, 1].sum())


def test_oracle_composite_train_vs_reference_kernels(gk):
    rays, m = _comp_layout(gk)
    sig, rgb, dl, T = gk["comp_sig"], gk["comp_rgb"], gk["march_a_deltas"], float(gk["comp_T"])
    ws, dep, img = oracle.composite_rays_train_forward(sig, rgb, dl, rays, T)
    # tolerance: __expf on the GPU vs libm expf in the oracle (2 ulp) through <= 256 sequential fp32 accumulations
    tol = dict(rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(ws, gk["comp_ws"], **tol)
    np.testing.assert_allclose(dep, gk["comp_depth"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(img, gk["comp_image"], **tol)
    gs, gr = oracle.composite_rays_train_backward(gk["comp_gws"], gk["comp_gimg"], sig, rgb, dl, rays, gk["comp_ws"],
                                                  gk["comp_image"], T)
    np.testing.assert_allclose(gr, gk["comp_grgb"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(gs, gk["comp_gsig"], rtol=1e-4, atol=2e-5)   # differences of O(1) sums times dt
    # the flex kernel drops a ray whose samples END at M (`>=`, raymarching.cu:601): give it one slot of slack
    fo = oracle.composite_rays_flex_train_forward(np.append(sig, 0).astype(np.float32), np.vstack([gk["comp_flex_in"], np.zeros((1, 5), np.float32)]),
                                                  np.vstack([dl, np.zeros((1, 2), np.float32)]), rays, T)
    np.testing.assert_allclose(fo, gk["comp_flex_out"], **tol)
    gi = oracle.composite_rays_flex_train_backward(gk["comp_gflex_out"], np.append(sig, 0).astype(np.float32),
                                                   np.vstack([dl, np.zeros((1, 2), np.float32)]), rays, 5, T)
    np.testing.assert_allclose(gi[:m], gk["comp_gflex_in"], **tol)
    sp = oracle.spread_ray_to_sample(gk["comp_gimg"], rays, m)
    assert np.array_equal(sp, gk["comp_spread"])


def test_oracle_inference_march_and_composite_vs_reference_kernels(gk):
    n_step, dt_gamma, max_steps, Mi = gk["inf_cfg"]
    n_step, max_steps, Mi = int(n_step), int(max_steps), int(Mi)
    alive = gk["inf_alive"]
    x, d, dl = oracle.march_rays(alive.shape[0], n_step, alive, gk["nears"], gk["rays_o"], gk["rays_d"], 2.0, gk["scene_bitfield"],
                                 2, 128, gk["nears"], gk["fars"], gk["inf_noises"], float(dt_gamma), max_steps, M=Mi)
    for a, b in ((x, gk["inf_xyzs"]), (d, gk["inf_dirs"]), (dl, gk["inf_deltas"])):
        assert np.array_equal(bits(a), bits(b))
    N = gk["rays_o"].shape[0]
    aux = oracle.composite_rays_flex(alive.shape[0], n_step, alive, gk["inf_sig"], gk["inf_flex"], dl, gk["inf_ws0"],
                                     np.zeros((N, 5), np.float32), 1e-2)
    al, rt, ws, dep, img = oracle.composite_rays(alive.shape[0], n_step, alive, gk["nears"], gk["inf_sig"], gk["inf_rgb"], dl,
                                                 gk["inf_ws0"], gk["inf_dep0"], gk["inf_img0"], 1e-2)
    tol = dict(rtol=2e-5, atol=2e-6)   # __expf vs expf, <= 4 accumulations
    assert np.array_equal(al, gk["inf_alive_out"])
    np.testing.assert_allclose(rt, gk["inf_rays_t_out"], **tol)
    np.testing.assert_allclose(ws, gk["inf_ws"], **tol)
    np.testing.assert_allclose(dep, gk["inf_depth"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(img, gk["inf_image"], **tol)
    np.testing.assert_allclose(aux, gk["inf_aux"], **tol)


def test_oracle_grid_encode_vs_reference_kernels(gk):
    L, C, Hb, log2T, pls = gk["grid_cfg"]
    L, C, Hb = int(L), int(C), int(Hb)
    offsets = gk["grid_offsets"]
    assert np.array_equal(offsets, O.grid_offsets(3, L, Hb, int(log2T), float(pls)))
    S = float(np.float32(np.log2(pls)))
    emb = table(int(offsets[-1]), C)
    out, dy = O.grid_encode_forward(gk["grid_x"], emb, offsets, S, Hb, with_dy_dx=True, exp2_levels=gk["grid_exp2_levels"])
    # fp32 kernel: 8 fp32 products summed per feature, O(1) table -> 1e-5 abs; fp16 kernel accumulates in half (gridencoder.cu:142)
    np.testing.assert_allclose(out, gk["grid_out_f32"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(out, gk["grid_out_f16"], atol=6e-3, rtol=0)
    assert np.all(gk["grid_out_f32"][2] == 0)                     # out-of-range point -> zero row
    B = gk["grid_x"].shape[0]
    np.testing.assert_allclose(dy.reshape(B, -1), gk["grid_dydx_f32"], atol=3e-4, rtol=1e-5)
    g = O.grid_encode_backward(gk["grid_grad"], gk["grid_x"], int(offsets[-1]), offsets, S, Hb, exp2_levels=gk["grid_exp2_levels"])
    np.testing.assert_allclose(g, gk["grid_gemb_f32"], atol=2e-5, rtol=1e-5)      # fp32 atomics, any order
    np.testing.assert_allclose(g, gk["grid_gemb_f16"], atol=3e-2, rtol=1e-2)      # half2 atomics
    gin = np.einsum("blc,bldc->bd", gk["grid_grad"].reshape(B, L, C).astype(np.float64), dy)
    np.testing.assert_allclose(gin, gk["grid_gin_f32"], atol=2e-3, rtol=1e-4)


def test_oracle_sh_freq_hsv_vs_reference_kernels(gk):
    for deg in (4, 8):
        out, grad = O.sh_encode(gk["sh_in"], deg, with_grad=True)
        np.testing.assert_allclose(out, gk[f"sh_out_{deg}"], atol=3e-6 * deg, rtol=1e-5)   # fp32 polynomial evaluation
        gin = np.einsum("bk,bdk->bd", gk[f"sh_grad_{deg}"].astype(np.float64), grad)
        np.testing.assert_allclose(gin, gk[f"sh_gin_{deg}"], atol=2e-4 * deg, rtol=1e-4)
    # freq: the reference is built with -use_fast_math (__sinf: abs error ~2^-21.4 in [-pi,pi], growing with |x| <= 32)
    np.testing.assert_allclose(O.freq_encode(gk["freq_in"], 6), gk["freq_out"], atol=2e-5, rtol=0)
    hsv = oracle.rgb_to_hsv(gk["hsv_rgb_in"])
    np.testing.assert_allclose(hsv, gk["hsv_out"], atol=2e-3, rtol=1e-5)                   # H in degrees, fast-math divides
    np.testing.assert_allclose(oracle.hsv_to_rgb(gk["hsv_out"]), gk["hsv_rgb_back"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(gk["hsv_rgb_back"], gk["hsv_rgb_in"], atol=1e-5, rtol=0)   # round trip of the reference


# =====================================================================================================================
# GPU: the CUDA path (through the C-ABI bindings) vs the same fixtures
# =====================================================================================================================
@pytest.mark.gpu
def test_cuda_raymarching_vs_reference_fixtures(cuda, gk):
    import torch
    from palettenerf_b200.raymarching.backend import _backend as B
    cu = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)  # noqa: E731
    o, d = cu(gk["rays_o"]), cu(gk["rays_d"])
    N = o.shape[0]
    nears, fars = torch.empty(N, device=cuda), torch.empty(N, device=cuda)
    B.near_far_from_aabb(o, d, cu(gk["aabb"]), N, 0.2, nears, fars)
    assert np.array_equal(bits(nears.cpu().numpy()), bits(gk["nears"])) and np.array_equal(bits(fars.cpu().numpy()), bits(gk["fars"]))
    idx = torch.empty(512, dtype=torch.int32, device=cuda)
    B.morton3D(cu(gk["morton_coords"]), 512, idx)
    assert np.array_equal(idx.cpu().numpy(), gk["morton_idx"])
    back = torch.empty(512, 3, dtype=torch.int32, device=cuda)
    B.morton3D_invert(idx, 512, back)
    assert np.array_equal(back.cpu().numpy(), gk["morton_back"])
    pb = torch.zeros(1024, dtype=torch.uint8, device=cuda)
    B.packbits(cu(gk["packbits_grid"]), 1024, 0.5, pb)
    assert np.array_equal(pb.cpu().numpy(), gk["packbits_out"])

    bitfield = cu(gk["scene_bitfield"])
    for tag in ("a", "b"):
        dt_gamma, max_steps, M = gk[f"march_{tag}_cfg"]
        max_steps, M = int(max_steps), int(M)
        xyzs, dirs, deltas = torch.zeros(M, 3, device=cuda), torch.zeros(M, 3, device=cuda), torch.zeros(M, 2, device=cuda)
        rays = torch.empty(N, 3, dtype=torch.int32, device=cuda)
        counter = torch.zeros(2, dtype=torch.int32, device=cuda)
        B.march_rays_train(o, d, bitfield, 2.0, float(dt_gamma), max_steps, N, 2, 128, M, nears, fars, xyzs, dirs, deltas, rays,
                           counter, cu(gk[f"march_{tag}_noises"]))
        assert np.array_equal(counter.cpu().numpy(), gk[f"march_{tag}_counter"])
        r, (sx, sd, sl) = sorted_view(rays.cpu().numpy(), [xyzs.cpu().numpy(), dirs.cpu().numpy(), deltas.cpu().numpy()], M)
        assert np.array_equal(r[:, [0, 2]], gk[f"march_{tag}_rays_sorted_counts"])
        assert np.array_equal(bits(sx), bits(gk[f"march_{tag}_xyzs"]))
        assert np.array_equal(bits(sl), bits(gk[f"march_{tag}_deltas"]))
        if tag != "a":
            continue
        # composites on this run's own layout; per-sample inputs come from the sorted-view fixtures
        rr = rays.cpu().numpy()
        sig, rgb, fin = scatter_sorted(rr, M, gk["comp_sig"], gk["comp_rgb"], gk["comp_flex_in"])
        T = float(gk["comp_T"])
        ws, dep, img = torch.empty(N, device=cuda), torch.empty(N, device=cuda), torch.empty(N, 3, device=cuda)
        B.composite_rays_train_forward(cu(sig), cu(rgb), deltas, rays, M, N, T, ws, dep, img)
        # same serial fp32 arithmetic and the same __expf as the reference kernel -> a few ulp at most
        tol = dict(rtol=2e-6, atol=2e-7)
        np.testing.assert_allclose(ws.cpu().numpy(), gk["comp_ws"], **tol)
        np.testing.assert_allclose(dep.cpu().numpy(), gk["comp_depth"], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(img.cpu().numpy(), gk["comp_image"], **tol)
        gsig, grgb = torch.zeros(M, device=cuda), torch.zeros(M, 3, device=cuda)
        B.composite_rays_train_backward(cu(gk["comp_gws"]), cu(gk["comp_gimg"]), cu(sig), cu(rgb), deltas, rays, cu(gk["comp_ws"]),
                                        cu(gk["comp_image"]), M, N, T, gsig, grgb)
        fout = torch.empty(N, 5, device=cuda)
        B.composite_rays_flex_train_forward(cu(sig), cu(fin), deltas, rays, M, N, 5, T, fout)
        gfin = torch.zeros(M, 5, device=cuda)
        B.composite_rays_flex_train_backward(cu(gk["comp_gflex_out"]), cu(sig), cu(fin), deltas, rays, fout, M, N, 5, T, gfin)
        spread = torch.zeros(M, 3, device=cuda)
        B.spread_ray_to_sample(cu(gk["comp_gimg"]), rays, M, N, 3, spread)
        _, (s_gsig, s_grgb, s_gfin, s_spread) = sorted_view(rr, [gsig.cpu().numpy(), grgb.cpu().numpy(), gfin.cpu().numpy(),
                                                                 spread.cpu().numpy()], M)
        np.testing.assert_allclose(s_grgb, gk["comp_grgb"], **tol)
        np.testing.assert_allclose(s_gsig, gk["comp_gsig"], rtol=2e-5, atol=5e-6)
        np.testing.assert_allclose(fout.cpu().numpy(), gk["comp_flex_out"], **tol)
        np.testing.assert_allclose(s_gfin, gk["comp_gflex_in"], **tol)
        assert np.array_equal(s_spread, gk["comp_spread"])


@pytest.mark.gpu
def test_cuda_inference_march_composite_vs_reference_fixtures(cuda, gk):
    import torch
    from palettenerf_b200.raymarching.backend import _backend as B
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)  # noqa: E731
    n_step, dt_gamma, max_steps, Mi = gk["inf_cfg"]
    n_step, max_steps, Mi = int(n_step), int(max_steps), int(Mi)
    alive = gk["inf_alive"]
    n_alive, N = alive.shape[0], gk["rays_o"].shape[0]
    xyzs, dirs, deltas = torch.zeros(Mi, 3, device=cuda), torch.zeros(Mi, 3, device=cuda), torch.zeros(Mi, 2, device=cuda)
    nears, fars = cu(gk["nears"]), cu(gk["fars"])
    B.march_rays(n_alive, n_step, cu(alive), nears.clone(), cu(gk["rays_o"]), cu(gk["rays_d"]), 2.0, float(dt_gamma), max_steps, 2,
                 128, cu(gk["scene_bitfield"]), nears, fars, xyzs, dirs, deltas, cu(gk["inf_noises"]))
    for a, b in ((xyzs, gk["inf_xyzs"]), (dirs, gk["inf_dirs"]), (deltas, gk["inf_deltas"])):
        assert np.array_equal(bits(a.cpu().numpy()), bits(b))
    al, rt = cu(alive), nears.clone()
    ws, dep, img, aux = cu(gk["inf_ws0"]), cu(gk["inf_dep0"]), cu(gk["inf_img0"]), torch.zeros(N, 5, device=cuda)
    B.composite_rays_flex(n_alive, n_step, 5, 1e-2, al, rt, cu(gk["inf_sig"]), cu(gk["inf_flex"]), deltas, ws, aux)
    B.composite_rays(n_alive, n_step, 1e-2, al, rt, cu(gk["inf_sig"]), cu(gk["inf_rgb"]), deltas, ws, dep, img)
    assert np.array_equal(al.cpu().numpy(), gk["inf_alive_out"])
    for a, b in ((rt, "inf_rays_t_out"), (ws, "inf_ws"), (dep, "inf_depth"), (img, "inf_image")):
        assert np.array_equal(bits(a.cpu().numpy()), bits(gk[b])), b          # same serial arithmetic: identical bits
    np.testing.assert_allclose(aux.cpu().numpy(), gk["inf_aux"], rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_cuda_encoders_vs_reference_fixtures(cuda, gk):
    import torch
    from palettenerf_b200.gridencoder.backend import _backend as GB
    from palettenerf_b200.shencoder.backend import _backend as SB
    from palettenerf_b200.freqencoder.backend import _backend as FB
    from palettenerf_b200.palette.backend import _backend as PB
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)  # noqa: E731
    L, C, Hb, log2T, pls = gk["grid_cfg"]
    L, C, Hb = int(L), int(C), int(Hb)
    offsets = cu(gk["grid_offsets"])
    S = float(np.float32(np.log2(pls)))
    emb32 = cu(table(int(gk["grid_offsets"][-1]), C))
    x = cu(gk["grid_x"])
    B = x.shape[0]
    for name, dt, tol_o, tol_g in (("f32", torch.float32, 1e-6, 1e-5), ("f16", torch.float16, 4e-3, 3e-2)):
        out = torch.empty(L, B, C, dtype=dt, device=cuda)
        dy = torch.empty(B, L * 3 * C, dtype=dt, device=cuda)
        GB.grid_encode_forward(x, emb32.to(dt), offsets, out, B, 3, C, L, S, Hb, dy, 0, False)
        got = out.permute(1, 0, 2).reshape(B, L * C).float().cpu().numpy()
        # fp32: same products, possibly a different summation order (<= 8 terms); fp16: the reference accumulates in
        # half (gridencoder.cu:142), this kernel in fp32 and rounds once -> two-sided tolerance of a few half ulps
        np.testing.assert_allclose(got, gk[f"grid_out_{name}"], atol=tol_o, rtol=0)
        gl = cu(gk["grid_grad"]).to(dt).view(B, L, C).permute(1, 0, 2).contiguous()
        gemb = torch.zeros(int(gk["grid_offsets"][-1]), C, dtype=dt, device=cuda)
        gin = torch.zeros(B, 3, dtype=dt, device=cuda)
        GB.grid_encode_backward(gl, x, emb32.to(dt), offsets, gemb, B, 3, C, L, S, Hb, dy, gin, 0, False)
        np.testing.assert_allclose(gemb.float().cpu().numpy(), gk[f"grid_gemb_{name}"], atol=tol_g, rtol=1e-2 if name == "f16" else 1e-5)
        np.testing.assert_allclose(gin.float().cpu().numpy(), gk[f"grid_gin_{name}"], atol=2e-3 if name == "f32" else 0.25,
                                   rtol=1e-4 if name == "f32" else 5e-2)
    for deg in (4, 8):
        out = torch.empty(64, deg * deg, device=cuda)
        dy = torch.empty(64, 3 * deg * deg, device=cuda)
        SB.sh_encode_forward(cu(gk["sh_in"]), out, 64, 3, deg, dy)
        np.testing.assert_allclose(out.cpu().numpy(), gk[f"sh_out_{deg}"], atol=2e-6 * deg, rtol=1e-5)
        gi = torch.zeros(64, 3, device=cuda)
        SB.sh_encode_backward(cu(gk[f"sh_grad_{deg}"]), cu(gk["sh_in"]), 64, 3, deg, dy, gi)
        np.testing.assert_allclose(gi.cpu().numpy(), gk[f"sh_gin_{deg}"], atol=1e-4 * deg, rtol=1e-4)
    fo = torch.empty(48, 39, device=cuda)
    FB.freq_encode_forward(cu(gk["freq_in"]), 48, 3, 6, 39, fo)
    np.testing.assert_allclose(fo.cpu().numpy(), gk["freq_out"], atol=2e-5, rtol=0)     # reference uses __sinf (fast math)
    fgi = torch.zeros(48, 3, device=cuda)
    FB.freq_encode_backward(cu(gk["freq_grad"]), fo, 48, 3, 6, 39, fgi)
    np.testing.assert_allclose(fgi.cpu().numpy(), gk["freq_gin"], atol=2e-3, rtol=1e-4)
    hsv = torch.empty(256, 3, device=cuda)
    PB.rgb_to_hsv(256, cu(gk["hsv_rgb_in"]), hsv)
    np.testing.assert_allclose(hsv.cpu().numpy(), gk["hsv_out"], atol=2e-3, rtol=1e-5)
    rgb = torch.empty(256, 3, device=cuda)
    PB.hsv_to_rgb(256, cu(gk["hsv_out"]), rgb)
    np.testing.assert_allclose(rgb.cpu().numpy(), gk["hsv_rgb_back"], atol=1e-5, rtol=0)
