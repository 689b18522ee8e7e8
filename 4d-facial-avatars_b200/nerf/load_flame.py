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


def _orbit_pose(theta_deg, phi_deg, radius):
    """Camera-to-world of the 40-view orbit the reference returns as render_poses (nerf/load_flame.py:33-38, 125-131):
    translate along z, pitch by phi, yaw by theta, then the Blender axis swap."""
    th, ph = np.deg2rad(theta_deg), np.deg2rad(phi_deg)
    t = np.eye(4, dtype=np.float32)
    t[2, 3] = radius
    rx = np.eye(4, dtype=np.float32)
    rx[1, 1] = rx[2, 2] = np.cos(ph)
    rx[1, 2], rx[2, 1] = -np.sin(ph), np.sin(ph)
    ry = np.eye(4, dtype=np.float32)
    ry[0, 0] = ry[2, 2] = np.cos(th)
    ry[0, 2], ry[2, 0] = -np.sin(th), np.sin(th)
    swap = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    return swap @ (ry @ (rx @ t))


def load_flame_data(basedir, half_res=False, testskip=1, debug=False, expressions=True, load_frontal_faces=False,
                    load_bbox=True, test=False):
    """Returns (imgs, poses, render_poses, [H, W, intrinsics], i_split, expressions, frontal_imgs, bboxs) like the
    reference; with test=True only the test split is read (images included: eval_transformed_rays.py:494 indexes them)."""
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
            imgs.append(_read_image(os.path.join(basedir, frame["file_path"] + ".png")))
            poses.append(np.array(frame["transform_matrix"]))
            exprs.append(np.array(frame["expression"]) if expressions else None)
            bboxs.append(np.array(frame["bbox"]) if (load_bbox and "bbox" in frame) else np.array([0.0, 1.0, 0.0, 1.0]))
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
    render_poses = torch.stack([torch.from_numpy(_orbit_pose(a, -30.0, 4.0)) for a in np.linspace(-180, 180, 40 + 1)[:-1]], 0)
    imgs_t = torch.from_numpy(imgs) if imgs is not None else None
    return (imgs_t, torch.from_numpy(poses), render_poses, [int(H), int(W), intrinsics], i_split,
            torch.from_numpy(exprs), None, torch.from_numpy(bboxs).int())
