"""Small tensor helpers the reference's scripts import from `nerf` (nerf/nerf_helpers.py).  Only get_ray_bundle is
on the render path, and the fused kernel can also generate rays itself (NfbRays.pose / intrinsics)."""
import math

import torch


def img2mse(img_src, img_tgt):
    return torch.nn.functional.mse_loss(img_src, img_tgt)


def mse2psnr(mse):
    if mse == 0:
        mse = 1e-5
    return -10.0 * math.log10(mse)


def get_minibatches(inputs, chunksize=1024 * 8):
    return [inputs[i:i + chunksize] for i in range(0, inputs.shape[0], chunksize)]


def meshgrid_xy(tensor1, tensor2):
    """np.meshgrid(..., indexing='xy') for two 1-D tensors."""
    ii, jj = torch.meshgrid(tensor1, tensor2, indexing="ij")
    return ii.transpose(-1, -2), jj.transpose(-1, -2)


def get_ray_bundle(height, width, intrinsics, tform_cam2world, center=(0.5, 0.5)):
    """Camera rays through every pixel (nerf_helpers.py:68-123): origins and UN-normalised directions, [H, W, 3].
    intrinsics = [fx, fy, cx, cy] with cx, cy relative to the image size; a scalar focal means [f, f, .5, .5]."""
    pose = tform_cam2world
    if not hasattr(intrinsics, "__len__") or len(intrinsics) < 4:
        f = float(intrinsics if not hasattr(intrinsics, "__len__") else intrinsics[0])
        intrinsics = [f, f, 0.5, 0.5]
    fx, fy, cx, cy = (float(v) for v in intrinsics[:4])
    col = torch.arange(width, dtype=pose.dtype, device=pose.device).view(1, width).expand(height, width)
    row = torch.arange(height, dtype=pose.dtype, device=pose.device).view(height, 1).expand(height, width)
    cam = torch.stack(((col - width * cx) / fx, -(row - height * cy) / fy, -torch.ones_like(col)), dim=-1)
    ray_directions = torch.sum(cam[..., None, :] * pose[:3, :3], dim=-1)
    ray_origins = pose[:3, -1].expand(ray_directions.shape)
    return ray_origins, ray_directions


def positional_encoding(tensor, num_encoding_functions=6, include_input=True, log_sampling=True):
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)] (nerf_helpers.py:195-239).  The fused
    kernel computes this internally; this torch version exists for callers that use the function directly."""
    parts = [tensor] if include_input else []
    if log_sampling:
        bands = 2.0 ** torch.linspace(0.0, num_encoding_functions - 1, num_encoding_functions,
                                      dtype=tensor.dtype, device=tensor.device)
    else:
        bands = torch.linspace(1.0, 2.0 ** (num_encoding_functions - 1), num_encoding_functions,
                               dtype=tensor.dtype, device=tensor.device)
    for f in bands:
        parts += [torch.sin(tensor * f), torch.cos(tensor * f)]
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)


def get_embedding_function(num_encoding_functions=6, include_input=True, log_sampling=True):
    fn = lambda x: positional_encoding(x, num_encoding_functions, include_input, log_sampling)  # noqa: E731
    fn.num_encoding_functions, fn.include_input, fn.log_sampling = num_encoding_functions, include_input, log_sampling
    return fn


def dump_rays(*args, **kwargs):
    raise NotImplementedError("dump_rays (PLY debugging aid) is outside the render path")
