"""PaletteNetwork — the palette-decomposition field, drop-in for palette/network.py:10-308 of the reference.

Three hash grids (`encoder`, `encoder_palette`, `encoder_clip`), SH(4) view encoding and six small MLPs:
  sigma_net 32-64-16 (ReLU) -> sigma = exp(h0), geo = h[1:16] (detached)
  diff_net  15-64-64-3 (sigmoid)                on geo
  color_net 31-64-64-3 (sigmoid)                on SH(d) ++ geo          ("view-dependent" colour)
  basis_net 35-64-15   (ELU)                    on grid_palette(x) ++ diffuse.detach()
  offsets_radiance_net Linear 15 -> 3*Nb+1 (bias), omega_net Linear 15 -> Nb + softplus, then (+0.05) / sum
  clip_net  32-64-clip_dim (ReLU)               on grid_clip(x), only with opt.pred_clip
`forward(x, d)` returns (sigma, clip_feat, omega, offsets_radiance, view_dep, diffuse) — public API (palette/gui.py:161).
State-dict keys equal the reference's so its checkpoints load.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..activation import trunc_exp
from ..encoding import get_encoder
from ..nerf.network import mlp, run_mlp
from .renderer import PaletteRenderer


class PaletteNetwork(PaletteRenderer):
    def __init__(self, opt, encoding="hashgrid", encoding_dir="sphere_harmonics", encoding_bg="hashgrid", num_layers=2,
                 hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, num_layers_bg=2, hidden_dim_bg=64,
                 bound=1, **kwargs):
        super().__init__(opt, bound, **kwargs)
        self.num_layers, self.hidden_dim, self.geo_feat_dim = num_layers, hidden_dim, geo_feat_dim
        self.num_layers_color, self.hidden_dim_color = num_layers_color, hidden_dim_color
        self.num_basis = opt.num_basis
        res = 2048 * bound
        self.encoder, self.in_dim = get_encoder(encoding, desired_resolution=res)
        self.encoder_palette, self.in_dim_palette = get_encoder(encoding, desired_resolution=res)
        self.encoder_clip, self.in_dim_clip = get_encoder(encoding, desired_resolution=res)  # always allocated (ref :33)
        self.encoder_dir, self.in_dim_dir = get_encoder(encoding_dir)

        hidden = [hidden_dim] * (num_layers - 1)
        hidden_c = [hidden_dim] * (num_layers_color - 1)
        self.sigma_net = mlp([self.in_dim] + hidden + [1 + geo_feat_dim])
        self.color_net = mlp([self.in_dim_dir + geo_feat_dim] + hidden_c + [3])   # name kept for stage-1 checkpoints
        self.diff_net = mlp([geo_feat_dim] + hidden_c + [3])
        self.basis_net = mlp([self.in_dim_palette + 3] + hidden + [geo_feat_dim])
        self.offsets_radiance_net = nn.Linear(geo_feat_dim, self.num_basis * 3 + 1)
        self.omega_net = nn.Sequential(nn.Linear(geo_feat_dim, self.num_basis, bias=False), nn.Softplus())
        if opt.pred_clip:
            self.clip_net = mlp([self.in_dim_clip] + hidden + [opt.clip_dim])
        if self.bg_radius > 0:
            self.num_layers_bg, self.hidden_dim_bg = num_layers_bg, hidden_dim_bg
            self.encoder_bg, self.in_dim_bg = get_encoder(encoding_bg, input_dim=2, num_levels=4, log2_hashmap_size=19,
                                                          desired_resolution=2048)
            self.bg_net = mlp([self.in_dim_bg + self.in_dim_dir] + [hidden_dim_bg] * (num_layers_bg - 1) + [3])
        else:
            self.bg_net = None

    # -- field queries --------------------------------------------------------------------------------
    def density(self, x):
        h = run_mlp(self.sigma_net, self.encoder(x, bound=self.bound))
        return {"sigma": trunc_exp(h[..., 0]), "geo_feat": h[..., 1:]}

    def _palette_heads(self, x, d, geo_feat):
        geo = geo_feat.detach()
        diffuse = torch.sigmoid(run_mlp(self.diff_net, geo))
        view_dep = torch.sigmoid(run_mlp(self.color_net, torch.cat([self.encoder_dir(d), geo], dim=-1)))
        feat = torch.cat([self.encoder_palette(x, bound=self.bound), diffuse.detach()], dim=-1)
        feat = run_mlp(self.basis_net, feat, act=F.elu)
        offsets_radiance = self.offsets_radiance_net(feat)
        omega = self.omega_net(feat) + 0.05
        omega = omega / omega.sum(dim=-1, keepdim=True)
        return omega, offsets_radiance, view_dep, diffuse

    def color(self, x, d, mask=None, geo_feat=None, **kwargs):
        if mask is None:
            return self._palette_heads(x, d, geo_feat)
        n, nb, kw = x.shape[0], self.num_basis, dict(dtype=x.dtype, device=x.device)
        omega = F.softmax(torch.zeros(n, nb, **kw), dim=-1)
        offsets_radiance = torch.zeros(n, 3 * nb, **kw)   # (sic) the reference's masked fallback has 3*Nb columns
        view_dep, diffuse = torch.zeros(n, 3, **kw), torch.zeros(n, 3, **kw)
        if mask.any():
            o, r, v, f = self._palette_heads(x[mask], d[mask], geo_feat[mask])
            omega[mask], view_dep[mask], diffuse[mask] = o.to(omega.dtype), v.to(view_dep.dtype), f.to(diffuse.dtype)
            offsets_radiance[mask] = r.to(offsets_radiance.dtype)
        return omega, offsets_radiance, view_dep, diffuse

    def forward(self, x, d):
        den = self.density(x)
        sigma = den["sigma"]
        if self.opt.pred_clip:
            clip_feat = run_mlp(self.clip_net, self.encoder_clip(x, bound=self.bound))
        else:
            clip_feat = torch.zeros(sigma.shape[0], self.opt.clip_dim, dtype=sigma.dtype, device=sigma.device)
        omega, offsets_radiance, view_dep, diffuse = self.color(x, d, geo_feat=den["geo_feat"].detach())
        return sigma, clip_feat, omega, offsets_radiance, view_dep, diffuse

    def background(self, x, d):
        h = torch.cat([self.encoder_dir(d), self.encoder_bg(x)], dim=-1)
        return torch.sigmoid(run_mlp(self.bg_net, h))

    def get_params(self, lr):
        # basis_net is absent on purpose: the reference never steps it (palette/network.py:283-308)
        groups = [self.encoder.parameters(), self.encoder_palette.parameters(), self.encoder_clip.parameters(),
                  self.sigma_net.parameters(), self.encoder_dir.parameters(), self.color_net.parameters(),
                  self.diff_net.parameters(), self.offsets_radiance_net.parameters(), self.omega_net.parameters(),
                  self.basis_color]
        if self.opt.use_initialization_from_rgbxy:
            groups.append(self.hist_weights)
        if self.bg_radius > 0:
            groups += [self.encoder_bg.parameters(), self.bg_net.parameters()]
        if self.opt.pred_clip:
            groups.append(self.clip_net.parameters())
        return [{"params": g, "lr": lr} for g in groups]
