"""Reader for the NeRFace dataset layout (transforms_{split}.json + frames + bg/), the schema of
nerf/load_flame.py:40-211.  I/O only; not on the render hot path."""
import json
import os

import numpy as np
import torch


def _read_image(path):
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 3:
        img = img[..., ::-1] if img.shape[2] == 3 else img[..., [2, 1, 0, 3]]
    return img


def load_flame_data(basedir, half_res=False, testskip=1, debug=False, expressions=True, load_frontal_faces=False,
                    load_bbox=True, test=False):
    """Returns (imgs, poses, render_poses, [H, W, intrinsics], i_split, expressions, frontal_imgs, bboxs) like the
    reference; with test=True only the test split's poses/expressions are read and images are skipped."""
    splits = ["test"] if test else ["train", "val", "test"]
    metas = {}
    for s in splits:
        with open(os.path.join(basedir, f"transforms_{s}.json"), "r") as fp:
            metas[s] = json.load(fp)
    all_imgs, all_poses, all_expr, all_bbox, counts = [], [], [], [], [0]
    for s in splits:
        meta = metas[s]
        skip = 1 if (s == "train" or testskip == 0) else testskip
        imgs, poses, exprs, bboxs = [], [], [], []
        for frame in meta["frames"][::skip]:
            if not test:
                imgs.append(_read_image(os.path.join(basedir, frame["file_path"] + ".png")))
            poses.append(np.array(frame["transform_matrix"]))
            exprs.append(np.array(frame["expression"]) if expressions else None)
            bboxs.append(np.array(frame["bbox"]) if (load_bbox and "bbox" in frame) else np.array([0.0, 1.0, 0.0, 1.0]))
        if not test:
            all_imgs.append((np.array(imgs) / 255.0).astype(np.float32))
        all_poses.append(np.array(poses).astype(np.float32))
        all_expr.append(np.array(exprs).astype(np.float32))
        all_bbox.append(np.array(bboxs).astype(np.float32))
        counts.append(counts[-1] + len(poses))
    i_split = [np.arange(counts[i], counts[i + 1]) for i in range(len(splits))]
    poses = np.concatenate(all_poses, 0)
    exprs = np.concatenate(all_expr, 0)
    bboxs = np.concatenate(all_bbox, 0)
    meta = metas[splits[0]]
    if test:
        first = _read_image(os.path.join(basedir, meta["frames"][0]["file_path"] + ".png")) \
            if os.path.exists(os.path.join(basedir, meta["frames"][0]["file_path"] + ".png")) else None
        H, W = (first.shape[:2] if first is not None else (512, 512))
        imgs = None
    else:
        imgs = np.concatenate(all_imgs, 0)
        H, W = imgs[0].shape[:2]
    intrinsics = np.array(meta["intrinsics"]) if "intrinsics" in meta else np.array(
        [0.5 * W / np.tan(0.5 * float(meta["camera_angle_x"]))] * 2 + [0.5, 0.5])
    if half_res:
        H, W = H // 2, W // 2
        intrinsics = intrinsics.copy()
        intrinsics[:2] = intrinsics[:2] * 0.5
        if imgs is not None:
            import cv2
            imgs = np.stack([cv2.resize(im, (W, H), interpolation=cv2.INTER_AREA) for im in imgs], 0)
    bboxs = bboxs.copy()
    bboxs[:, 0:2] *= H
    bboxs[:, 2:4] *= W
    bboxs = np.floor(bboxs)
    render_poses = torch.from_numpy(poses[:1].copy())
    imgs_t = torch.from_numpy(imgs) if imgs is not None else None
    return (imgs_t, torch.from_numpy(poses), render_poses, [int(H), int(W), intrinsics], i_split,
            torch.from_numpy(exprs), None, torch.from_numpy(bboxs).int())
