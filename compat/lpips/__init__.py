"""stand-in for lpips: no pretrained network is available offline; the 'loss' is identically zero"""
import torch
import torch.nn as nn


class LPIPS(nn.Module):
    def __init__(self, net="alex", **kwargs):
        super().__init__()
        self.net = net

    def forward(self, x, y, normalize=False, **kwargs):
        return torch.zeros(x.shape[0], 1, 1, 1, dtype=x.dtype, device=x.device)
