#!/usr/bin/env python
"""Golden vectors from the reference's own CUDA KERNELS, generated on a GPU box.

TEST INFRASTRUCTURE. Run on a B200 (e.g. `gpurun -- python tests/golden/make_golden_gpu.py gpurun_out/golden`):
it loads ONLY the reference extensions that oracle/build_ref.py compiled from /root/reference sources into
oracle/_ref/ (none of this repository's kernels), feeds them small seeded inputs and stores inputs + outputs in
<outdir>/ref_kernels.npz. The file is then committed as tests/golden/ref_kernels.npz; the CPU suite checks the oracle
against it (tests/test_golden.py, `-m "not gpu"`), the GPU suite checks the CUDA path against it.

The reference has no stored known-answer vectors for this path; these fixtures are "outputs of the
reference itself", which is what pins the oracle. Inputs are built with numpy's PCG64 (stable across platforms) or
integer formulas, never with torch's device RNG.

Layout notes
  * march_rays_train assigns sample slots through an atomicAdd race (raymarching.cu:413-416), so its outputs are stored
    in the canonical ray-id-sorted view: `rays` sorted by id, and each ray's rows concatenated in that order.
  * hash-grid tables are not stored: entry (i, c) = ((i*2654435761 + c*40503 + 12345) mod 2^16) / 2^15 - 1, exact in
    fp16/fp32 (`table()` below; tests rebuild it).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_ref  # noqa: E402
from palettenerf_b200 import synthetic as S  # noqa: E402  (host-side scene generator only; no kernels)


def table(n_entries, C):
    i = np.arange(n_entries, dtype=np.uint64)[:, None]
    c = np.arange(C, dtype=np.uint64)[None, :]
    v = (i * np.uint64(2654435761) + c * np.uint64(40503) + np.uint64(12345)) % np.uint64(65536)
    return (v.astype(np.float64) / 32768.0 - 1.0).astype(np.float32)


def sorted_view(rays, arrs, M):
    r = rays[np.argsort(rays[:, 0], kind="stable")]
    outs = []
    for a in arrs:
        rows = [a[off:off + cnt] for _, off, cnt in r if cnt > 0 and off + cnt <= M]
        outs.append(np.concatenate(rows) if rows else np.zeros((0,) + a.shape[1:], a.dtype))
    return r, outs


def main(outdir):
    dev = torch.device("cuda:0")
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    rm, ge, sh, fe, pf = (load_ref(n) for n in ("raymarching", "gridencoder", "shencoder", "freqencoder", "palette_func"))
    assert all(m is not None for m in (rm, ge, sh, fe, pf)), "oracle/_ref/*.so missing: run oracle/build_ref.py first"
    rng = np.random.default_rng(7)
    G = {}

    # ------------------------------------------------------------------ scene (host-side generator)
    grid = S.density_grid()
    thresh = min(grid.clamp(min=0).mean().item(), S.LEGO["density_thresh"])
    G["scene_thresh"] = np.float32(thresh)
    bitfield = torch.zeros(grid.numel() // 8, dtype=torch.uint8, device=dev)
    rm.packbits(grid.to(dev), bitfield.numel(), float(thresh), bitfield)
    G["scene_bitfield"] = bitfield.cpu().numpy()          # 512 KiB, compresses to a few KB

    # ------------------------------------------------------------------ near_far / morton / packbits
    o, d = S.camera_rays(12, 12, azimuth_deg=35.0)
    o, d = o.numpy().copy(), d.numpy().copy()
    o[5] = (0, 0, 5); d[5] = (0, 1, 0)        # misses the box
    d[6] = (0, 0, -1)                         # axis-aligned (1/0 = inf)
    o[7] = (0.5, 0.25, 0.1)                   # origin inside the box
    N = o.shape[0]
    aabb = np.array([-2, -2, -2, 2, 2, 2], np.float32)
    nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
    rm.near_far_from_aabb(cu(o), cu(d), cu(aabb), N, 0.2, nears, fars)
    G.update(rays_o=o, rays_d=d, aabb=aabb, nears=nears.cpu().numpy(), fars=fars.cpu().numpy())

    coords = rng.integers(0, 1024, size=(512, 3)).astype(np.int32)
    idx = torch.empty(512, dtype=torch.int32, device=dev)
    rm.morton3D(cu(coords), 512, idx)
    back = torch.empty(512, 3, dtype=torch.int32, device=dev)
    rm.morton3D_invert(idx, 512, back)
    G.update(morton_coords=coords, morton_idx=idx.cpu().numpy(), morton_back=back.cpu().numpy())

    pg = rng.uniform(-0.2, 1.0, size=8192).astype(np.float32)
    pg[::17] = 0.5          # equal to the threshold: strict '>' leaves the bit clear
    pg[::29] = -1.0         # "untrained" cells
    pb = torch.zeros(1024, dtype=torch.uint8, device=dev)
    rm.packbits(cu(pg), 1024, 0.5, pb)
    G.update(packbits_grid=pg, packbits_out=pb.cpu().numpy())

    # ------------------------------------------------------------------ march_rays_train (+ composites on its samples)
    for tag, dt_gamma, max_steps in (("a", 0.0, 256), ("b", 1.0 / 128, 1024)):
        noises = rng.uniform(0, 1, size=N).astype(np.float32)
        M = N * max_steps if max_steps <= 256 else N * 128
        xyzs, dirs, deltas = (torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev))
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        rm.march_rays_train(cu(o), cu(d), bitfield, 2.0, dt_gamma, max_steps, N, 2, 128, M, nears, fars, xyzs, dirs, deltas,
                            rays, counter, cu(noises))
        cnt = counter.cpu().numpy()
        m = int(cnt[0])
        assert m <= M
        r_sorted, (sx, sd, sl) = sorted_view(rays.cpu().numpy(), [xyzs.cpu().numpy(), dirs.cpu().numpy(), deltas.cpu().numpy()], M)
        G.update({f"march_{tag}_noises": noises, f"march_{tag}_cfg": np.array([dt_gamma, max_steps, M], np.float64),
                  f"march_{tag}_counter": cnt, f"march_{tag}_rays_sorted_counts": r_sorted[:, [0, 2]],
                  f"march_{tag}_xyzs": sx, f"march_{tag}_deltas": sl})
        # dirs are per-sample copies of rays_d: store the verdict instead of the rows
        want = np.concatenate([np.repeat(d[rid][None], c, axis=0) for rid, _, c in r_sorted if c > 0] or [np.zeros((0, 3), np.float32)])
        G[f"march_{tag}_dirs_ok"] = np.array(np.array_equal(sd, want))

        if tag == "a":
            # composite on the reference's own (race-ordered) layout; outputs are per ray id, hence layout independent
            sig = np.zeros(M, np.float32); rgb = np.zeros((M, 3), np.float32); flex_in = np.zeros((M, 5), np.float32)
            # per-sample inputs are defined in the SORTED view and scattered into the raced layout
            ssig = rng.uniform(0, 40, size=m).astype(np.float32)
            srgb = rng.uniform(0, 1, size=(m, 3)).astype(np.float32)
            sflex = rng.uniform(-1, 1, size=(m, 5)).astype(np.float32)
            rr = rays.cpu().numpy()
            order = rr[np.argsort(rr[:, 0], kind="stable")]
            k = 0
            for _, off, c in order:
                if c > 0 and off + c <= M:
                    sig[off:off + c] = ssig[k:k + c]; rgb[off:off + c] = srgb[k:k + c]; flex_in[off:off + c] = sflex[k:k + c]
                    k += c
            assert k == m
            T_thresh = 1e-2
            ws, dep, img = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev)
            rm.composite_rays_train_forward(cu(sig), cu(rgb), deltas, rays, M, N, T_thresh, ws, dep, img)
            gws = rng.normal(size=N).astype(np.float32); gimg = rng.normal(size=(N, 3)).astype(np.float32)
            gsig, grgb = torch.zeros(M, device=dev), torch.zeros(M, 3, device=dev)
            rm.composite_rays_train_backward(cu(gws), cu(gimg), cu(sig), cu(rgb), deltas, rays, ws, img, M, N, T_thresh, gsig, grgb)
            fout = torch.empty(N, 5, device=dev)
            rm.composite_rays_flex_train_forward(cu(sig), cu(flex_in), deltas, rays, M, N, 5, T_thresh, fout)
            gfo = rng.normal(size=(N, 5)).astype(np.float32)
            gfin = torch.zeros(M, 5, device=dev)
            rm.composite_rays_flex_train_backward(cu(gfo), cu(sig), cu(flex_in), deltas, rays, fout, M, N, 5, T_thresh, gfin)
            spread = torch.zeros(M, 3, device=dev)
            rm.spread_ray_to_sample(cu(gimg), rays, M, N, 3, spread)
            # per-ray inputs/outputs are indexed by RAY ID inside the kernels (raymarching.cu:519,527), so they are layout
            # independent as they are; per-sample outputs go through the sorted view
            _, (s_gsig, s_grgb, s_gfin, s_spread) = sorted_view(rr, [gsig.cpu().numpy(), grgb.cpu().numpy(), gfin.cpu().numpy(),
                                                                     spread.cpu().numpy()], M)
            G.update(comp_sig=ssig, comp_rgb=srgb, comp_flex_in=sflex, comp_T=np.float32(T_thresh),
                     comp_ws=ws.cpu().numpy(), comp_depth=dep.cpu().numpy(), comp_image=img.cpu().numpy(),
                     comp_gws=gws, comp_gimg=gimg, comp_gsig=s_gsig, comp_grgb=s_grgb, comp_flex_out=fout.cpu().numpy(),
                     comp_gflex_out=gfo, comp_gflex_in=s_gfin, comp_spread=s_spread)

    # ------------------------------------------------------------------ inference march + composites (one iteration)
    n_step = 4
    alive = rng.permutation(N)[: N // 2].astype(np.int32)
    n_alive = alive.shape[0]
    Mi = n_alive * n_step + 128
    inoise = rng.uniform(0, 1, size=n_alive).astype(np.float32)
    xyzs, dirs, deltas = (torch.zeros(Mi, 3, device=dev), torch.zeros(Mi, 3, device=dev), torch.zeros(Mi, 2, device=dev))
    rays_t = nears.clone()
    rm.march_rays(n_alive, n_step, cu(alive), rays_t, cu(o), cu(d), 2.0, 1.0 / 256, 1024, 2, 128, bitfield, nears, fars,
                  xyzs, dirs, deltas, cu(inoise))
    isig = rng.uniform(0, 60, size=Mi).astype(np.float32)
    irgb = rng.uniform(0, 1, size=(Mi, 3)).astype(np.float32)
    iflex = rng.uniform(0, 1, size=(Mi, 5)).astype(np.float32)
    ws0 = rng.uniform(0, 0.5, size=N).astype(np.float32); dep0 = rng.uniform(0, 1, size=N).astype(np.float32)
    img0 = rng.uniform(0, 1, size=(N, 3)).astype(np.float32)
    al, rt = cu(alive), rays_t.clone()
    ws, dep, img, aux = cu(ws0), cu(dep0), cu(img0), torch.zeros(N, 5, device=dev)
    rm.composite_rays_flex(n_alive, n_step, 5, 1e-2, al, rt, cu(isig), cu(iflex), deltas, ws, aux)
    rm.composite_rays(n_alive, n_step, 1e-2, al, rt, cu(isig), cu(irgb), deltas, ws, dep, img)
    G.update(inf_alive=alive, inf_noises=inoise, inf_cfg=np.array([n_step, 1.0 / 256, 1024, Mi], np.float64),
             inf_xyzs=xyzs.cpu().numpy(), inf_dirs=dirs.cpu().numpy(), inf_deltas=deltas.cpu().numpy(),
             inf_sig=isig, inf_rgb=irgb, inf_flex=iflex, inf_ws0=ws0, inf_dep0=dep0, inf_img0=img0,
             inf_alive_out=al.cpu().numpy(), inf_rays_t_out=rt.cpu().numpy(), inf_ws=ws.cpu().numpy(),
             inf_depth=dep.cpu().numpy(), inf_image=img.cpu().numpy(), inf_aux=aux.cpu().numpy())

    # ------------------------------------------------------------------ hash grid fwd / bwd (fp32 and fp16)
    from oracle import cpu_oracle as O  # offsets formula only (checked against the reference in ref_python.npz)
    L, C, Hb, log2T, pls = 6, 2, 4, 10, 1.5
    offsets = O.grid_offsets(3, L, Hb, log2T, pls)
    Sf = float(np.float32(np.log2(pls)))
    B = 160
    x = rng.uniform(0, 1, size=(B, 3)).astype(np.float32)
    x[0] = (0, 0, 0); x[1] = (1, 1, 1); x[2] = (1.0001, 0.5, 0.5)   # the last one is out of range -> zero output row
    emb = table(int(offsets[-1]), C)
    grad = rng.normal(size=(B, L * C)).astype(np.float32)
    # exp2f(level * S) as the device math library rounds it (the oracle takes it as an input, see cpu_oracle._level_setup)
    exp2_levels = torch.exp2(torch.arange(L, dtype=torch.float32, device=dev) * Sf).cpu().numpy()
    G.update(grid_cfg=np.array([L, C, Hb, log2T, pls], np.float64), grid_offsets=np.asarray(offsets, np.int32), grid_x=x,
             grid_grad=grad, grid_exp2_levels=exp2_levels)
    for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
        out = torch.empty(L, B, C, dtype=dt, device=dev)
        dy = torch.empty(B, L * 3 * C, dtype=dt, device=dev)
        ge.grid_encode_forward(cu(x), cu(emb).to(dt), cu(np.asarray(offsets, np.int32)), out, B, 3, C, L, Sf, Hb, dy, 0, False)
        gl = cu(grad).to(dt).view(B, L, C).permute(1, 0, 2).contiguous()
        gemb = torch.zeros(int(offsets[-1]), C, dtype=dt, device=dev)
        gin = torch.zeros(B, 3, dtype=dt, device=dev)
        ge.grid_encode_backward(gl, cu(x), cu(emb).to(dt), cu(np.asarray(offsets, np.int32)), gemb, B, 3, C, L, Sf, Hb, dy, gin,
                                0, False)
        G[f"grid_out_{name}"] = out.permute(1, 0, 2).reshape(B, L * C).float().cpu().numpy()
        G[f"grid_gemb_{name}"] = gemb.float().cpu().numpy()
        G[f"grid_gin_{name}"] = gin.float().cpu().numpy()
        G[f"grid_dydx_{name}"] = dy.float().cpu().numpy()

    # ------------------------------------------------------------------ SH / freq / hsv
    dd = rng.normal(size=(64, 3)); dd = (dd / np.linalg.norm(dd, axis=-1, keepdims=True)).astype(np.float32)
    G["sh_in"] = dd
    for deg in (4, 8):
        out = torch.empty(64, deg * deg, device=dev)
        dy = torch.empty(64, 3 * deg * deg, device=dev)
        sh.sh_encode_forward(cu(dd), out, 64, 3, deg, dy)
        g = rng.normal(size=(64, deg * deg)).astype(np.float32)
        gi = torch.zeros(64, 3, device=dev)
        sh.sh_encode_backward(cu(g), cu(dd), 64, 3, deg, dy, gi)
        G.update({f"sh_out_{deg}": out.cpu().numpy(), f"sh_grad_{deg}": g, f"sh_gin_{deg}": gi.cpu().numpy()})

    fx = rng.uniform(-1, 1, size=(48, 3)).astype(np.float32)
    deg = 6
    Cf = 3 + 3 * 2 * deg
    fo = torch.empty(48, Cf, device=dev)
    fe.freq_encode_forward(cu(fx), 48, 3, deg, Cf, fo)
    fg = rng.normal(size=(48, Cf)).astype(np.float32)
    fgi = torch.zeros(48, 3, device=dev)
    fe.freq_encode_backward(cu(fg), fo, 48, 3, deg, Cf, fgi)
    G.update(freq_in=fx, freq_out=fo.cpu().numpy(), freq_grad=fg, freq_gin=fgi.cpu().numpy())

    col = rng.uniform(0, 1, size=(256, 3)).astype(np.float32)
    col[0] = (0, 0, 0); col[1] = (1, 1, 1); col[2] = (0.5, 0.5, 0.5); col[3] = (1, 0, 0); col[4] = (0, 1, 0); col[5] = (0, 0, 1)
    hsv = torch.empty(256, 3, device=dev)
    pf.rgb_to_hsv(256, cu(col), hsv)
    rgb2 = torch.empty(256, 3, device=dev)
    pf.hsv_to_rgb(256, hsv, rgb2)
    G.update(hsv_rgb_in=col, hsv_out=hsv.cpu().numpy(), hsv_rgb_back=rgb2.cpu().numpy())

    torch.cuda.synchronize()
    os.makedirs(outdir, exist_ok=True)
    path = os.path.join(outdir, "ref_kernels.npz")
    np.savez_compressed(path, **G)
    print(f"wrote {path}: {len(G)} arrays, {os.path.getsize(path)} bytes; gpu = {torch.cuda.get_device_name(0)}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
