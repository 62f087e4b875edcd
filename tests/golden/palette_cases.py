"""Inputs shared by tests/golden/make_golden_palette.py (which runs the REFERENCE on them) and the parity tests (which run
the oracle and the CUDA path on them). Everything is derived from seeds on the CPU, so both sides see the same bits.

Cases (SURVEY §8d "second parity run"): hash tables U(-0.5, 0.5) instead of the near-zero reference init (otherwise every
tolerance is vacuous), `density_scale` 1 (rays cross the whole solid) and 40 (alpha ~ 0.13 per sample: rays terminate
on T < T_thresh inside the solid), with and without the semantic-feature branch (`--pred_clip`).
"""
import contextlib
import types

import numpy as np
import torch

MODEL_CASES = {
    "noclip": dict(seed=1, pred_clip=False, table_scale=0.5),
    "clip": dict(seed=1, pred_clip=True, table_scale=0.5),
}
DENSITY_SCALES = (1.0, 40.0)
EVAL_SIDE = 32            # 32 x 32 view
TRAIN_RAYS = 1024
FWD_SAMPLES = 2048
RENDER_KW = dict(dt_gamma=0.0, max_steps=1024, T_thresh=1e-4)
LAMBDAS = dict(lambda_sparsity=2e-4, lambda_smooth=4e-3, lambda_patchsmooth=0.0, lambda_view_dep=0.1, lambda_offsets=0.03,
               lambda_weight=0.05, lambda_palette=0.001)       # main_palette.py:83-89 defaults
GRAD_SCALE = 1024.0       # static loss scale of the fp16 runs (the reference trains under GradScaler)
TABLE_GRAD_SAMPLES = 4096


def make_opt(pred_clip, **kw):
    """the argparse fields of main_palette.py that the model / train_step read"""
    opt = types.SimpleNamespace(num_basis=4, clip_dim=16, pred_clip=pred_clip, test=True, use_initialization_from_rgbxy=False,
                                color_space="srgb", smooth_sigma_xyz=0.005, smooth_sigma_color=0.2, smooth_sigma_clip=0.0,
                                random_size=0, patch_size=1, **LAMBDAS, **RENDER_KW)
    for k, v in kw.items():
        setattr(opt, k, v)
    return opt


def build_model(case, device="cpu"):
    """this repository's model for a case (palettenerf_b200.synthetic.build_palette_model on the CPU RNG)"""
    from palettenerf_b200 import synthetic as S
    c = MODEL_CASES[case]
    return S.build_palette_model(device, seed=c["seed"], pred_clip=c["pred_clip"], table_scale=c["table_scale"])


def eval_rays():
    from palettenerf_b200 import synthetic as S
    return S.camera_rays(EVAL_SIDE, EVAL_SIDE)


def train_rays():
    from palettenerf_b200 import synthetic as S
    return S.training_rays(TRAIN_RAYS, H=200, W=200, seed=3, n_views=4)


def train_targets(pred_clip):
    g = torch.Generator().manual_seed(77)
    gt = torch.rand(1, TRAIN_RAYS, 3, generator=g)
    feat = torch.randn(1, TRAIN_RAYS, 16, generator=g) * 0.3 if pred_clip else None
    return gt, feat


def table_grad_indices(n_entries):
    g = torch.Generator().manual_seed(99)
    return torch.randint(0, n_entries, (TABLE_GRAD_SAMPLES,), generator=g)


def hash_uniform(t):
    """U[0,1) noise that is a pure function of the BITS of each row of a float32 [M,3] tensor (exact integer arithmetic,
    so every device and both implementations agree): the smooth-loss jitter of a sample must not depend on the row the
    sample landed in — the reference's march hands out rows by an atomic race (raymarching.cu:405-411)."""
    b = t.detach().contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    h = ((b[:, 0] * 73856093) ^ (b[:, 1] * 19349663) ^ (b[:, 2] * 83492791)) & 0xFFFFFFFF
    cols = []
    for j in range(3):
        v = (h + (j + 1) * 0x9E3779B9) & 0xFFFFFFFF
        v = ((v ^ (v >> 15)) * 0x2C1B3C6D) & 0xFFFFFFFF
        v = ((v ^ (v >> 12)) * 0x297A2D39) & 0xFFFFFFFF
        v = v ^ (v >> 15)
        cols.append((v & 0xFFFFFF).to(torch.float32) / 16777216.0)
    return torch.stack(cols, dim=-1)


class FixedRandom(contextlib.AbstractContextManager):
    """Replaces torch.rand / torch.rand_like by deterministic streams for the duration of one model call, so the reference
    and this repository draw the same numbers:
      torch.rand(N)        (march noise, raymarching.py:213-216; indexed by ray) -> call k returns the first N values of
                           torch.rand(2^21, generator=seed(1000 + k)) on the CPU;
      torch.rand_like(xyz) (smooth-loss jitter, palette/renderer.py:362; indexed by sample ROW, and rows are assigned by
                           an atomic race in the reference) -> hash_uniform(xyz): a function of the sample, not its row."""
    POOL = 1 << 21

    def __init__(self):
        self.calls = 0
        self._orig = None

    def _draw(self, shape, device, dtype):
        n = int(np.prod(shape)) if len(shape) else 1
        if n > self.POOL:
            raise RuntimeError(f"FixedRandom: {n} values requested, pool is {self.POOL}")
        real_rand = self._orig[0] if self._orig else torch.rand
        pool = real_rand(self.POOL, generator=torch.Generator().manual_seed(1000 + self.calls))
        self.calls += 1
        return pool[:n].reshape(shape).to(device=device, dtype=dtype or torch.float32)

    def __enter__(self):
        self._orig = (torch.rand, torch.rand_like)

        def rand(*size, dtype=None, device=None, generator=None, **kw):
            if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
                size = tuple(size[0])
            return self._draw(tuple(size), device or "cpu", dtype)

        def rand_like(t, dtype=None, device=None, **kw):
            if t.dim() == 2 and t.shape[1] == 3 and t.dtype == torch.float32:
                return hash_uniform(t)
            return self._draw(tuple(t.shape), device or t.device, dtype or t.dtype)

        torch.rand, torch.rand_like = rand, rand_like
        return self

    def __exit__(self, *exc):
        torch.rand, torch.rand_like = self._orig
        return False
