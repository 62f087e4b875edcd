"""stand-in for kornia: only kornia.losses.ssim_loss is used (SSIMMeter of the reference)"""
from . import losses  # noqa: F401
