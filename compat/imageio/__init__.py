"""stand-in for imageio: imwrite through cv2, mimwrite is a no-op (no video encoder in this image)"""
import numpy as np


def imwrite(path, image, **kwargs):
    import cv2
    a = np.asarray(image)
    if a.ndim == 3 and a.shape[-1] == 3:
        a = a[..., ::-1]
    elif a.ndim == 3 and a.shape[-1] == 4:
        a = a[..., [2, 1, 0, 3]]
    cv2.imwrite(str(path), a)


def imread(path, **kwargs):
    import cv2
    a = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if a is not None and a.ndim == 3:
        a = a[..., ::-1] if a.shape[-1] == 3 else a[..., [2, 1, 0, 3]]
    return a


def mimwrite(path, frames, **kwargs):
    return None
