"""NeRFNetwork — stage-1 Instant-NGP field (hash grid -> sigma MLP 32-64-16, SH(4) + 15 geo -> colour MLP
31-64-64-3), drop-in for nerf/network.py:9-206 (same ctor arguments and state_dict keys `encoder.*`,
`sigma_net.{i}.weight`, `color_net.{i}.weight`, optional `encoder_bg.*` / `bg_net.*`)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..activation import trunc_exp
from ..encoding import get_encoder
from .renderer import NeRFRenderer


def mlp(dims):
    """bias-free Linear stack as an nn.ModuleList (activation applied by the caller)"""
    return nn.ModuleList([nn.Linear(i, o, bias=False) for i, o in zip(dims[:-1], dims[1:])])


def run_mlp(layers, h, act=F.relu):
    last = len(layers) - 1
    for i, layer in enumerate(layers):
        h = layer(h)
        if i != last:
            h = act(h, inplace=True)
    return h


class NeRFNetwork(NeRFRenderer):
    def __init__(self, encoding="hashgrid", encoding_dir="sphere_harmonics", encoding_bg="hashgrid", num_layers=2,
                 hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, num_layers_bg=2, hidden_dim_bg=64,
                 bound=1, **kwargs):
        super().__init__(bound, **kwargs)
        self.num_layers, self.hidden_dim, self.geo_feat_dim = num_layers, hidden_dim, geo_feat_dim
        self.num_layers_color, self.hidden_dim_color = num_layers_color, hidden_dim_color
        self.encoder, self.in_dim = get_encoder(encoding, desired_resolution=2048 * bound)
        self.sigma_net = mlp([self.in_dim] + [hidden_dim] * (num_layers - 1) + [1 + geo_feat_dim])
        self.encoder_dir, self.in_dim_dir = get_encoder(encoding_dir)
        # NB: like the reference, hidden layers of the colour net use `hidden_dim` (nerf/network.py:59)
        self.color_net = mlp([self.in_dim_dir + geo_feat_dim] + [hidden_dim] * (num_layers_color - 1) + [3])
        if self.bg_radius > 0:
            self.num_layers_bg, self.hidden_dim_bg = num_layers_bg, hidden_dim_bg
            self.encoder_bg, self.in_dim_bg = get_encoder(encoding_bg, input_dim=2, num_levels=4, log2_hashmap_size=19,
                                                          desired_resolution=2048)
            self.bg_net = mlp([self.in_dim_bg + self.in_dim_dir] + [hidden_dim_bg] * (num_layers_bg - 1) + [3])
        else:
            self.bg_net = None

    def density(self, x):
        h = run_mlp(self.sigma_net, self.encoder(x, bound=self.bound))
        return {"sigma": trunc_exp(h[..., 0]), "geo_feat": h[..., 1:]}

    def color(self, x, d, mask=None, geo_feat=None, **kwargs):
        if mask is not None:
            rgbs = torch.zeros(mask.shape[0], 3, dtype=x.dtype, device=x.device)
            if not mask.any():
                return rgbs
            d, geo_feat = d[mask], geo_feat[mask]
        h = torch.sigmoid(run_mlp(self.color_net, torch.cat([self.encoder_dir(d), geo_feat], dim=-1)))
        if mask is None:
            return h
        rgbs[mask] = h.to(rgbs.dtype)
        return rgbs

    def forward(self, x, d):
        den = self.density(x)
        return den["sigma"], self.color(x, d, geo_feat=den["geo_feat"])

    def background(self, x, d):
        h = torch.cat([self.encoder_dir(d), self.encoder_bg(x)], dim=-1)
        return torch.sigmoid(run_mlp(self.bg_net, h))

    def get_params(self, lr):
        groups = [self.encoder, self.sigma_net, self.encoder_dir, self.color_net]
        if self.bg_radius > 0:
            groups += [self.encoder_bg, self.bg_net]
        return [{"params": g.parameters(), "lr": lr} for g in groups]
