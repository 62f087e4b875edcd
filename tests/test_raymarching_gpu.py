"""GPU parity of the ray-marching kernels through the C ABI: new CUDA path vs the CPU oracle (bit-exact for the
index/coordinate outputs) and, when oracle/_ref holds the reference's own extension, vs the reference kernels."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_ref
from palettenerf_b200 import synthetic as S
from palettenerf_b200.raymarching.backend import _backend as B, OCC_FLOATS
import palettenerf_b200.raymarching as rm
from palettenerf_b200.raymarching.raymarching import occupied_bounds as _occupied_bounds

pytestmark = pytest.mark.gpu


def _rays(n_side, az=35.0):
    return S.camera_rays(n_side, n_side, azimuth_deg=az)


def _sorted_rays(rays):
    r = rays.cpu().numpy() if torch.is_tensor(rays) else rays
    return r[np.argsort(r[:, 0], kind="stable")]


def _gather_samples(arr, rays_sorted, M):
    """concatenate each ray's rows in ray-id order (canonical view of a race-ordered layout)"""
    out = []
    for rid, off, cnt in rays_sorted:
        if cnt > 0 and off + cnt <= M:
            out.append(arr[off:off + cnt])
    return np.concatenate(out) if out else np.zeros((0,) + arr.shape[1:], arr.dtype)


def test_near_far_morton_packbits_bit_exact(cuda, scene):
    o, d = _rays(64)
    o[5] = torch.tensor([0.0, 0.0, 5.0]); d[5] = torch.tensor([0.0, 1.0, 0.0])   # misses the box
    d[6] = torch.tensor([0.0, 0.0, -1.0])                                          # axis-aligned: 1/0 = inf
    nears, fars = rm.near_far_from_aabb(o.to(cuda), d.to(cuda), scene["aabb"].to(cuda), 0.2)
    on, of = oracle.near_far_from_aabb(o.numpy(), d.numpy(), scene["aabb"].numpy(), 0.2)
    assert np.array_equal(nears.cpu().numpy().view(np.uint32), on.view(np.uint32))
    assert np.array_equal(fars.cpu().numpy().view(np.uint32), of.view(np.uint32))
    assert nears[5].item() == np.finfo(np.float32).max

    coords = torch.randint(0, 128, (100000, 3), dtype=torch.int32)
    idx = rm.morton3D(coords.to(cuda))
    assert np.array_equal(idx.cpu().numpy(), oracle.morton3D(coords.numpy()))
    back = rm.morton3D_invert(idx)
    assert torch.equal(back.cpu(), coords)
    assert rm.morton3D(torch.zeros(0, 3, dtype=torch.int32, device=cuda)).numel() == 0  # empty input

    grid = scene["grid"].clone()
    grid[0, :1000] = -1.0  # "untrained" cells never set a bit
    for thresh in (scene["thresh"], 0.0, 19.999):
        bf = rm.packbits(grid.to(cuda), thresh)
        assert np.array_equal(bf.cpu().numpy(), oracle.packbits(grid.numpy(), thresh))
    # ragged size (N not a multiple of 4 bytes) + unaligned view
    g = torch.rand(8 * 1021 + 8, device=cuda)[8:]
    out = torch.empty(1021, dtype=torch.uint8, device=cuda)
    B.packbits(g, 1021, 0.5, out)
    assert np.array_equal(out.cpu().numpy(), oracle.packbits(g.cpu().numpy(), 0.5))


@pytest.mark.parametrize("cfg", [
    dict(bound=2.0, C=2, dt_gamma=0.0, max_steps=1024, perturb=False),
    dict(bound=2.0, C=2, dt_gamma=1.0 / 128, max_steps=1024, perturb=True),
    dict(bound=1.5, C=2, dt_gamma=0.0, max_steps=512, perturb=True),     # non power-of-two bound: mip_bound = 1.5
    dict(bound=1.0, C=1, dt_gamma=0.0, max_steps=256, perturb=False),
])
def test_march_rays_train_bit_exact_vs_oracle_and_reference(cuda, scene, cfg):
    ref = load_ref("raymarching")
    o, d = _rays(96, az=20.0)
    N, H, C = o.shape[0], 128, cfg["C"]
    bound = cfg["bound"]
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32)
    bitfield = scene["bitfield"][: C * H ** 3 // 8].contiguous()
    nears, fars = oracle.near_far_from_aabb(o.numpy(), d.numpy(), aabb.numpy(), 0.2)
    noises = torch.rand(N, generator=torch.Generator().manual_seed(1)) if cfg["perturb"] else torch.zeros(N)
    M = N * 48

    def run(backend):
        xyzs = torch.zeros(M, 3, device=cuda); dirs = torch.zeros(M, 3, device=cuda); deltas = torch.zeros(M, 2, device=cuda)
        rays = torch.empty(N, 3, dtype=torch.int32, device=cuda); counter = torch.zeros(2, dtype=torch.int32, device=cuda)
        backend.march_rays_train(o.to(cuda), d.to(cuda), bitfield.to(cuda), bound, cfg["dt_gamma"], cfg["max_steps"], N, C,
                                 H, M, torch.from_numpy(nears).to(cuda), torch.from_numpy(fars).to(cuda), xyzs, dirs,
                                 deltas, rays, counter, noises.to(cuda))
        torch.cuda.synchronize()
        return xyzs.cpu().numpy(), dirs.cpu().numpy(), deltas.cpu().numpy(), rays.cpu().numpy(), counter.cpu().numpy()

    x, dr, dl, rays, cnt = run(B)
    # the one-walk schedule (t-list + occupied bounds, pnerf_march_rays_train_ws) must give the same bits
    for use_occ in (False, True):
        got = run(_WsBackend(use_occ, cuda))
        for a, b in zip(got, (x, dr, dl, rays, cnt)):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"one-walk march differs (occupied bounds: {use_occ})"
    ox, odr, odl, orays, ocnt = oracle.march_rays_train(o.numpy(), d.numpy(), bitfield.numpy(), bound, cfg["dt_gamma"],
                                                        cfg["max_steps"], C, H, M, nears, fars, noises.numpy())
    assert np.array_equal(cnt, ocnt) and cnt[1] == N and cnt[0] == rays[:, 2].sum()
    assert np.array_equal(rays, orays)                       # deterministic scan == oracle's ray order
    assert rays[:, 2].max() > 0
    for a, b in ((x, ox), (dr, odr), (dl, odl)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    used = int(min(cnt[0], M))
    assert not x[used:].any() and not dl[used:].any()       # tail stays zero

    if ref is not None:
        rx, rdr, rdl, rrays, rcnt = run(ref)
        assert np.array_equal(rcnt, cnt)
        rs, ns = _sorted_rays(rrays), _sorted_rays(rays)
        assert np.array_equal(rs[:, 0], ns[:, 0]) and np.array_equal(rs[:, 2], ns[:, 2])   # per-ray counts exact
        if cnt[0] <= M:  # no overflow: every ray was written by both
            for a, b in ((x, rx), (dr, rdr), (dl, rdl)):
                ga, gb = _gather_samples(a, ns, M), _gather_samples(b, rs, M)
                assert np.array_equal(ga.view(np.uint32), gb.view(np.uint32))


class _WsBackend:
    """pnerf_march_rays_train_ws behind the reference's march_rays_train signature"""

    def __init__(self, use_occ, dev):
        self.use_occ, self.dev = use_occ, dev

    def march_rays_train(self, o, d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays,
                         counter, noises):
        t_list = torch.full((N * max_steps,), float("nan"), device=self.dev)
        occ = None
        if self.use_occ:
            occ = torch.empty(OCC_FLOATS, device=self.dev)
            B.occupied_bounds(grid, C, H, bound, occ)
        B.march_rays_train_ws(o, d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays,
                              counter, noises, t_list, occ)


def _np_occupied_bounds(bitfield, C, H, bound):
    """numpy restatement of k_occupied_bounds (2x2x2 blocks, one cell of padding, sides at the bound opened)"""
    big = np.float32(3.402823466e+38)
    lo, hi = np.full(3, big, np.float32), np.full(3, -big, np.float32)
    per = H ** 3 // 8
    for level in range(C):
        nz = np.nonzero(bitfield[level * per:(level + 1) * per])[0].astype(np.uint32)
        if nz.size == 0:
            continue
        c = oracle.morton3D_invert((nz * 8).astype(np.int32)).astype(np.float32)
        mb = np.float32(min(2.0 ** level, bound))
        lo = np.minimum(lo, ((((c - 1) / np.float32(H)) * 2 - 1) * mb).min(axis=0))
        hi = np.maximum(hi, ((((c + 3) / np.float32(H)) * 2 - 1) * mb).max(axis=0))
    cell = np.float32(min(2.0 ** (C - 1), bound) * 2 / H)
    lo = np.where(lo <= -np.float32(bound) + cell, -big, lo)
    hi = np.where(hi >= np.float32(bound) - cell, big, hi)
    return np.concatenate([lo, hi]).astype(np.float32)


def test_occupied_bounds_vs_numpy(cuda, scene):
    bf = scene["bitfield"]
    occ = torch.empty(OCC_FLOATS, device=cuda)
    B.occupied_bounds(bf.to(cuda), 2, 128, 2.0, occ)
    occ = occ[:6]
    exp = _np_occupied_bounds(bf.numpy(), 2, 128, 2.0)
    assert np.allclose(occ.cpu().numpy(), exp, rtol=1e-6, atol=1e-6)
    assert (occ[:3] < -0.2).all() and (occ[3:] > 0.2).all() and (occ.abs() < 1.0).all()   # lego-shaped solid, well inside
    # empty grid: lo > hi on every axis, and the wrapper cache follows in-place writes of the bitfield
    z = torch.zeros_like(bf).to(cuda)
    o1 = _occupied_bounds(z, 2, 128, 2.0)
    assert (o1[:3] > o1[3:6]).all()
    z[12345] = 255
    o2 = _occupied_bounds(z, 2, 128, 2.0)
    assert (o2[:3] < o2[3:6]).all() and o2 is not o1


@pytest.mark.parametrize("seed,dt_gamma", [(0, 0.0), (1, 1.0 / 128), (2, 0.0)])
def test_one_walk_march_on_scattered_occupancy_touching_the_bound(cuda, seed, dt_gamma):
    """adversarial grid: sparse random cells in both cascades INCLUDING the faces of the scene bound (where positions are
    clamped) — the occupied-bounds clip and the t-list writer must not change a single bit of the march"""
    g = torch.Generator().manual_seed(seed)
    H, C, bound = 128, 2, 2.0
    bf = torch.zeros(C * H ** 3 // 8, dtype=torch.uint8)
    idx = torch.randint(0, bf.numel(), (400,), generator=g)
    bf[idx] = torch.randint(1, 256, (400,), generator=g).to(torch.uint8)
    if seed == 2:   # cluster only: tight bounds, most of every ray is tail
        bf.zero_()
        coords = torch.randint(60, 70, (300, 3), generator=g).int()
        codes = torch.from_numpy(oracle.morton3D(coords.numpy())).long()
        bf[codes // 8] = 255
    o, d = _rays(64, az=75.0)
    N = o.shape[0]
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = oracle.near_far_from_aabb(o.numpy(), d.numpy(), aabb, 0.05)
    noises = torch.rand(N, generator=g)
    M = N * 256

    def run(backend):
        xyzs = torch.zeros(M, 3, device=cuda); dirs = torch.zeros(M, 3, device=cuda); deltas = torch.zeros(M, 2, device=cuda)
        rays = torch.empty(N, 3, dtype=torch.int32, device=cuda); counter = torch.zeros(2, dtype=torch.int32, device=cuda)
        backend.march_rays_train(o.to(cuda), d.to(cuda), bf.to(cuda), bound, dt_gamma, 1024, N, C, H, M,
                                 torch.from_numpy(nears).to(cuda), torch.from_numpy(fars).to(cuda), xyzs, dirs, deltas, rays,
                                 counter, noises.to(cuda))
        return [t.cpu().numpy() for t in (xyzs, dirs, deltas, rays, counter)]

    base = run(B)
    assert base[4][0] > 0
    for use_occ in (False, True):
        for a, b in zip(run(_WsBackend(use_occ, cuda)), base):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_march_rays_train_overflow_drops_rays_like_reference(cuda, scene):
    o, d = _rays(48)
    N = o.shape[0]
    nears, fars = oracle.near_far_from_aabb(o.numpy(), d.numpy(), scene["aabb"].numpy(), 0.2)
    M = 4096  # far too small
    xyzs = torch.zeros(M, 3, device=cuda); dirs = torch.zeros(M, 3, device=cuda); deltas = torch.zeros(M, 2, device=cuda)
    rays = torch.empty(N, 3, dtype=torch.int32, device=cuda); counter = torch.zeros(2, dtype=torch.int32, device=cuda)
    B.march_rays_train(o.to(cuda), d.to(cuda), scene["bitfield"].to(cuda), 2.0, 0.0, 1024, N, 2, 128, M,
                       torch.from_numpy(nears).to(cuda), torch.from_numpy(fars).to(cuda), xyzs, dirs, deltas, rays, counter,
                       torch.zeros(N, device=cuda))
    ox, _, odl, orays, ocnt = oracle.march_rays_train(o.numpy(), d.numpy(), scene["bitfield"].numpy(), 2.0, 0.0, 1024, 2, 128,
                                                      M, nears, fars, np.zeros(N, np.float32))
    assert counter.cpu().numpy()[0] > M
    assert np.array_equal(counter.cpu().numpy(), ocnt) and np.array_equal(rays.cpu().numpy(), orays)
    assert np.array_equal(xyzs.cpu().numpy(), ox) and np.array_equal(deltas.cpu().numpy(), odl)


def test_march_rays_train_wrapper_matches_reference_contract(cuda, scene):
    o, d = _rays(32)
    nears, fars = rm.near_far_from_aabb(o.to(cuda), d.to(cuda), scene["aabb"].to(cuda), 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=cuda)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o.to(cuda), d.to(cuda), 2.0, scene["bitfield"].to(cuda), 2, 128, nears,
                                                   fars, counter, -1, False, 128, True, 0.0, 1024)
    m = counter[0].item()
    assert xyzs.shape[0] == m + 128 - m % 128 and dirs.shape == xyzs.shape and deltas.shape == (xyzs.shape[0], 2)
    assert rays.shape == (o.shape[0], 3) and rays.dtype == torch.int32
    assert (deltas[m:] == 0).all() and (deltas[:m, 0] > 0).all()


@pytest.mark.parametrize("n_step", [1, 4, 8])
def test_inference_march_and_composite_vs_oracle(cuda, scene, n_step):
    ref = load_ref("raymarching")
    o, d = _rays(80, az=60.0)
    N = o.shape[0]
    nears, fars = oracle.near_far_from_aabb(o.numpy(), d.numpy(), scene["aabb"].numpy(), 0.2)
    g = torch.Generator().manual_seed(n_step)
    alive = torch.randperm(N, generator=g)[: N // 2].int()
    n_alive = alive.shape[0]
    rays_t = torch.from_numpy(nears.copy())
    noises = torch.rand(n_alive, generator=g)
    M = n_alive * n_step + 128

    def run_march(backend):
        xyzs = torch.zeros(M, 3, device=cuda); dirs = torch.zeros(M, 3, device=cuda); deltas = torch.zeros(M, 2, device=cuda)
        backend.march_rays(n_alive, n_step, alive.to(cuda), rays_t.to(cuda), o.to(cuda), d.to(cuda), 2.0, 1.0 / 256, 1024, 2,
                           128, scene["bitfield"].to(cuda), torch.from_numpy(nears).to(cuda),
                           torch.from_numpy(fars).to(cuda), xyzs, dirs, deltas, noises.to(cuda))
        return xyzs, dirs, deltas

    xyzs, dirs, deltas = run_march(B)
    ox, odr, odl = oracle.march_rays(n_alive, n_step, alive.numpy(), rays_t.numpy(), o.numpy(), d.numpy(), 2.0,
                                     scene["bitfield"].numpy(), 2, 128, nears, fars, noises.numpy(), 1.0 / 256, 1024, M=M)
    for a, b in ((xyzs, ox), (dirs, odr), (deltas, odl)):
        assert np.array_equal(a.cpu().numpy().view(np.uint32), b.view(np.uint32))
    assert (deltas[:, 0] > 0).any()
    if ref is not None:
        rx, rd, rl = run_march(ref)
        assert torch.equal(rx, xyzs) and torch.equal(rd, dirs) and torch.equal(rl, deltas)

    # composite the marched samples with synthetic sigmas / colours; two rounds to exercise carried state
    sig = torch.rand(M, generator=g) * 60.0
    rgb = torch.rand(M, 3, generator=g)
    ws0 = torch.rand(N, generator=g) * 0.5; dep0 = torch.rand(N, generator=g); img0 = torch.rand(N, 3, generator=g)
    T_thresh = 1e-2

    def run_comp(backend):
        al, rt = alive.clone().to(cuda), rays_t.clone().to(cuda)
        ws, dep, img = ws0.clone().to(cuda), dep0.clone().to(cuda), img0.clone().to(cuda)
        aux = torch.zeros(N, 5, device=cuda)
        backend.composite_rays_flex(n_alive, n_step, 5, T_thresh, al, rt, sig.to(cuda), torch.rand(M, 5, generator=torch.Generator().manual_seed(3)).to(cuda), deltas, ws, aux)
        backend.composite_rays(n_alive, n_step, T_thresh, al, rt, sig.to(cuda), rgb.to(cuda), deltas, ws, dep, img)
        return al, rt, ws, dep, img, aux

    al, rt, ws, dep, img, aux = run_comp(B)
    inp5 = torch.rand(M, 5, generator=torch.Generator().manual_seed(3))
    oaux = oracle.composite_rays_flex(n_alive, n_step, alive.numpy(), sig.numpy(), inp5.numpy(), odl, ws0.numpy(),
                                      np.zeros((N, 5), np.float32), T_thresh)
    oal, ort, ows, odep, oimg = oracle.composite_rays(n_alive, n_step, alive.numpy(), rays_t.numpy(), sig.numpy(), rgb.numpy(),
                                                      odl, ws0.numpy(), dep0.numpy(), img0.numpy(), T_thresh)
    # tolerance: __expf (2 ulp + range reduction) vs libm expf, <= 8 sequential fp32 accumulations
    tol = dict(rtol=2e-5, atol=2e-6)
    assert np.array_equal(al.cpu().numpy(), oal)            # same rays terminate
    np.testing.assert_allclose(rt.cpu().numpy(), ort, **tol)
    np.testing.assert_allclose(ws.cpu().numpy(), ows, **tol)
    np.testing.assert_allclose(dep.cpu().numpy(), odep, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(img.cpu().numpy(), oimg, **tol)
    np.testing.assert_allclose(aux.cpu().numpy(), oaux, **tol)
    assert (al.cpu().numpy() == -1).any() and (al.cpu().numpy() >= 0).any()
    if ref is not None:  # same serial arithmetic as the reference kernel -> identical bits
        ral, rrt, rws, rdep, rimg, raux = run_comp(ref)
        assert torch.equal(ral, al) and torch.equal(rrt, rt) and torch.equal(rws, ws) and torch.equal(rdep, dep)
        assert torch.equal(rimg, img)
        np.testing.assert_allclose(raux.cpu().numpy(), aux.cpu().numpy(), rtol=1e-6, atol=1e-7)


def test_sph_from_ray_and_spread(cuda, scene):
    ref = load_ref("raymarching")
    o, d = _rays(32)
    c = rm.sph_from_ray((o * 0.1).to(cuda), d.to(cuda), 3.0).cpu().numpy()
    oo, dd = (o * 0.1).double().numpy(), d.double().numpy()
    A = (dd * dd).sum(1); Bh = (oo * dd).sum(1); Cc = (oo * oo).sum(1) - 9.0
    t = (-Bh + np.sqrt(Bh * Bh - A * Cc)) / A
    p = oo + t[:, None] * dd
    theta = np.arctan2(np.sqrt(p[:, 0] ** 2 + p[:, 2] ** 2), p[:, 1]); phi = np.arctan2(p[:, 2], p[:, 0])
    np.testing.assert_allclose(c, np.stack([2 * theta / np.pi - 1, phi / np.pi], 1), atol=2e-6)
    if ref is not None:
        rc = torch.empty(o.shape[0], 2, device=cuda)
        ref.sph_from_ray((o * 0.1).to(cuda), d.to(cuda), 3.0, o.shape[0], rc)
        np.testing.assert_allclose(c, rc.cpu().numpy(), atol=1e-6)

    N = 300
    g = torch.Generator().manual_seed(0)
    counts = torch.randint(0, 40, (N,), generator=g); counts[::7] = 0
    offs = torch.cumsum(counts, 0) - counts
    perm = torch.randperm(N, generator=g)
    rays = torch.stack([perm, offs, counts], 1).int()
    Mtot = int(counts.sum())
    inp = torch.rand(N, 3, generator=g)
    for M in (Mtot, Mtot - 25):  # second case: clipped at M
        out = torch.zeros(M, 3, device=cuda)
        rm.spread_ray_to_sample(inp.to(cuda), rays.to(cuda), out)
        assert np.array_equal(out.cpu().numpy(), oracle.spread_ray_to_sample(inp.numpy(), rays.numpy(), M))
