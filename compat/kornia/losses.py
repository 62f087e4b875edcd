import torch
import torch.nn.functional as F


def _gauss(window_size, sigma, dtype, device):
    x = torch.arange(window_size, dtype=dtype, device=device) - window_size // 2
    g = torch.exp(-(x ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def ssim_loss(img1, img2, window_size=11, max_val=1.0, eps=1e-12, reduction="mean"):
    """dissimilarity (1 - SSIM) / 2 with a separable Gaussian window (sigma 1.5), like kornia's"""
    c = img1.shape[1]
    g = _gauss(window_size, 1.5, img1.dtype, img1.device)
    k = (g[:, None] * g[None, :])[None, None].repeat(c, 1, 1, 1)
    pad = window_size // 2
    f = lambda x: F.conv2d(F.pad(x, (pad,) * 4, mode="reflect"), k, groups=c)  # noqa: E731
    mu1, mu2 = f(img1), f(img2)
    s11, s22, s12 = f(img1 * img1) - mu1 ** 2, f(img2 * img2) - mu2 ** 2, f(img1 * img2) - mu1 * mu2
    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    ssim = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 ** 2 + mu2 ** 2 + c1) * (s11 + s22 + c2) + eps)
    loss = torch.clamp((1 - ssim) / 2, 0, 1)
    return loss.mean() if reduction == "mean" else loss.sum() if reduction == "sum" else loss
