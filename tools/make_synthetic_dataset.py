#!/usr/bin/env python
"""Write a synthetic dataset in the layout the reference's scripts read (schema of nerf/load_flame.py:40-211
and real_to_nerf.py:1435-1483 in the reference): transforms_{train,val,test}.json with camera_angle_x, intrinsics
[fx, fy, cx, cy] (cx, cy relative) and per frame {file_path, bbox (relative), transform_matrix 4x4, expression[76]};
PNG frames; bg/00050.png; index_map.npy [N, 2].  There is no network for real data, so this is what the launcher
(4d-facial-avatars_b200/run_reference_script.py) and the loader tests run on.

    python tools/make_synthetic_dataset.py OUT_DIR [--size 64] [--train 8] [--val 2] [--test 12] [--seed 42]
"""
import argparse
import json
import math
import os

import numpy as np


def _pose(rng):
    yaw, pitch = (rng.uniform(-1.0, 1.0, size=2) * math.radians(15.0)).tolist()
    ry = np.array([[math.cos(yaw), 0.0, math.sin(yaw)], [0.0, 1.0, 0.0], [-math.sin(yaw), 0.0, math.cos(yaw)]])
    rx = np.array([[1.0, 0.0, 0.0], [0.0, math.cos(pitch), -math.sin(pitch)], [0.0, math.sin(pitch), math.cos(pitch)]])
    m = np.eye(4)
    m[:3, :3] = ry @ rx
    m[:3, 3] = [0.0, 0.0, 0.5]   # the authors scale scenes so that the mean camera distance is 0.5 (near 0.2 / far 0.8)
    return m


def _write_png(path, img_u8):
    import cv2
    os.makedirs(os.path.dirname(path), exist_ok=True)
    cv2.imwrite(path, img_u8[..., ::-1])  # RGB -> BGR


def write_dataset(out_dir, size=64, n_train=8, n_val=2, n_test=12, seed=42):
    rng = np.random.default_rng(seed)
    h = w = int(size)
    intrinsics = [1200.0 * w / 512.0, 1200.0 * h / 512.0, 0.5, 0.5]
    camera_angle_x = 2.0 * math.atan(0.5 * w / intrinsics[0])
    bg = (rng.uniform(0.0, 1.0, size=(h, w, 3)) * 255.0).astype(np.uint8)
    _write_png(os.path.join(out_dir, "bg", "00050.png"), bg)
    yy, xx = np.mgrid[0:h, 0:w]
    index = 0
    for split, count in (("train", n_train), ("val", n_val), ("test", n_test)):
        frames = []
        for k in range(count):
            expr = rng.normal(0.0, 0.5, size=76)
            cx, cy = rng.uniform(0.4, 0.6, size=2)
            rad = rng.uniform(0.2, 0.3)
            # a coloured disc over the background: something with a bounding box for the importance sampler
            mask = ((xx / w - cx) ** 2 + (yy / h - cy) ** 2) < rad ** 2
            img = bg.copy()
            img[mask] = (rng.uniform(0.2, 1.0, size=3) * 255.0).astype(np.uint8)
            rel = f"./{split}/f_{index:04d}"
            _write_png(os.path.join(out_dir, rel + ".png"), img)
            bbox = [max(0.0, cy - rad), min(1.0, cy + rad), max(0.0, cx - rad), min(1.0, cx + rad)]  # rows then columns, relative
            frames.append({"file_path": rel, "bbox": [float(v) for v in bbox], "transform_matrix": _pose(rng).tolist(),
                           "expression": [float(v) for v in expr]})
            index += 1
        with open(os.path.join(out_dir, f"transforms_{split}.json"), "w") as fp:
            json.dump({"camera_angle_x": camera_angle_x, "intrinsics": intrinsics, "frames": frames}, fp)
    n_total = n_train + n_val + n_test
    np.save(os.path.join(out_dir, "index_map.npy"), np.stack([np.arange(n_total), np.arange(n_total)], axis=1))
    return dict(height=h, width=w, intrinsics=intrinsics, frames=n_total)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("out_dir")
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--train", type=int, default=8)
    ap.add_argument("--val", type=int, default=2)
    ap.add_argument("--test", type=int, default=12)
    ap.add_argument("--seed", type=int, default=42)
    a = ap.parse_args()
    print(write_dataset(a.out_dir, a.size, a.train, a.val, a.test, a.seed))
