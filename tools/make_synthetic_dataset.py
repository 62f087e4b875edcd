#!/usr/bin/env python
"""Write a tiny Blender-format (NeRF-synthetic) dataset for the reference's unmodified mains: transforms_{train,val,test}.json
+ RGBA PNGs of an analytic scene (a coloured box with a sphere, the solid of palettenerf_b200/synthetic.py, flat-shaded by a
ray cast on the CPU). usage: python tools/make_synthetic_dataset.py <out_dir> [side=64] [n_train=12]"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pose_blender(radius, az_deg, el_deg):
    """camera-to-world in the Blender / NeRF-synthetic convention (camera looks along -z, y up)"""
    az, el = math.radians(az_deg), math.radians(el_deg)
    pos = np.array([radius * math.cos(el) * math.cos(az), radius * math.cos(el) * math.sin(az), radius * math.sin(el)])
    fwd = -pos / np.linalg.norm(pos)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up); right /= np.linalg.norm(right)
    cam_up = np.cross(right, fwd)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, cam_up, -fwd, pos
    return m


def render(pose, side, angle_x):
    """RGBA image: ray-box / ray-sphere intersection of the synthetic solid, colour from the hit normal"""
    f = 0.5 * side / math.tan(0.5 * angle_x)
    j, i = np.meshgrid(np.arange(side) + 0.5, np.arange(side) + 0.5, indexing="ij")
    d = np.stack([(i - side / 2) / f, -(j - side / 2) / f, -np.ones_like(i)], -1)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    d = d @ pose[:3, :3].T
    o = np.broadcast_to(pose[:3, 3], d.shape)
    lo, hi = np.array([-0.55, -0.40, -0.30]) * 1.6, np.array([0.55, 0.40, 0.30]) * 1.6
    with np.errstate(divide="ignore", invalid="ignore"):
        t0, t1 = (lo - o) / d, (hi - o) / d
    tn, tf = np.minimum(t0, t1).max(-1), np.maximum(t0, t1).min(-1)
    hit = (tn < tf) & (tf > 0)
    p = o + tn[..., None] * d
    n = np.abs(p / (hi + 1e-9))
    axis = n.argmax(-1)
    palette = np.array([[0.85, 0.70, 0.15], [0.20, 0.25, 0.65], [0.75, 0.20, 0.20]])
    rgb = palette[axis] * (0.6 + 0.4 * np.clip(-(d * np.sign(p))[np.arange(side)[:, None], np.arange(side)[None, :], axis], 0, 1))[..., None]
    img = np.zeros((side, side, 4))
    img[..., :3] = np.where(hit[..., None], rgb, 0)
    img[..., 3] = hit
    return (img * 255).astype(np.uint8)


def main(out, side=64, n_train=12):
    import cv2
    os.makedirs(out, exist_ok=True)
    angle_x = 0.6911
    rng = np.random.default_rng(0)
    for split, n in (("train", n_train), ("val", 2), ("test", 3)):
        os.makedirs(os.path.join(out, split), exist_ok=True)
        frames = []
        for k in range(n):
            pose = pose_blender(4.031, 360.0 * k / n + (0 if split == "train" else 17), 20 + 25 * rng.random())
            img = render(pose, side, angle_x)
            name = f"./{split}/r_{k}"
            cv2.imwrite(os.path.join(out, split, f"r_{k}.png"), img[..., [2, 1, 0, 3]])
            frames.append({"file_path": name, "rotation": 0.0, "transform_matrix": pose.tolist()})
        json.dump({"camera_angle_x": angle_x, "frames": frames}, open(os.path.join(out, f"transforms_{split}.json"), "w"), indent=1)
    print(f"[make_synthetic_dataset] {out}: {n_train} train / 2 val / 3 test views of {side}x{side}")


if __name__ == "__main__":
    main(sys.argv[1], *(int(a) for a in sys.argv[2:]))
