"""CPU oracle for the NeRFace per-ray render path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement (torch, CPU, FP32) of the reference's hot
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``4d-facial-avatars_b200``) never does: it fails loudly if the CUDA library is missing.

Pinning: the reference ships no tests or golden vectors, so this
oracle is pinned against the reference itself, executed in the build container by
``oracle/make_golden.py`` (which imports ``/root/reference`` read-only); the outputs
are committed under ``tests/golden/`` and ``tests/test_oracle_golden.py`` checks this
file against them.

Reference anchors (relative to nerface_code/nerf-pytorch/nerf/):
  ray_bundle          nerf_helpers.py:68-123   (get_ray_bundle)
  posenc              nerf_helpers.py:195-239  (positional_encoding)
  mlp_forward         models.py:236-261        (ConditionalBlendshapePaperNeRFModel.forward)
  composite           volume_rendering_utils.py:7-75 + nerf_helpers.py:44-65
  resample            nerf_helpers.py:344-387  (sample_pdf_2)
  render_chunk        train_utils.py:36-162    (predict_and_render_radiance) + :9-33 (run_network)
  run_one_iter        train_utils.py:165-290   (run_one_iter_of_nerf)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor

# state_dict key order of the paper model (models.py:218-233); (out, in) shapes.
PAPER_MODEL_SHAPES: List[Tuple[str, Tuple[int, ...]]] = []
for _i, _in in enumerate([171, 256, 256, 427, 256, 256]):
    PAPER_MODEL_SHAPES += [(f"layers_xyz.{_i}.weight", (256, _in)), (f"layers_xyz.{_i}.bias", (256,))]
PAPER_MODEL_SHAPES += [("fc_feat.weight", (256, 256)), ("fc_feat.bias", (256,)),
                       ("fc_alpha.weight", (1, 256)), ("fc_alpha.bias", (1,))]
for _i, _in in enumerate([280, 128, 128, 128]):
    PAPER_MODEL_SHAPES += [(f"layers_dir.{_i}.weight", (128, _in)), (f"layers_dir.{_i}.bias", (128,))]
PAPER_MODEL_SHAPES += [("fc_rgb.weight", (3, 128)), ("fc_rgb.bias", (3,))]

DIM_XYZ, DIM_DIR, DIM_EXPR, DIM_LATENT = 63, 24, 76, 32


def random_init_params(seed: int, stress: bool = False) -> Dict[str, Tensor]:
    """torch.nn.Linear default init (kaiming-uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)) for
    both weight and bias), drawn in state_dict order from one seeded CPU generator.  ``stress``
    applies SURVEY.md §8(d)'s opaque-stress scaling so compositing/resampling see opaque rays."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    fan_in = 1
    for name, shape in PAPER_MODEL_SHAPES:
        if name.endswith("weight"):
            fan_in = shape[1]
        bound = 1.0 / math.sqrt(fan_in)
        out[name] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound
    if stress:
        for name in out:
            if name.endswith("weight"):
                out[name] = out[name] * 2.0
        out["fc_alpha.weight"] = out["fc_alpha.weight"] * 20.0  # x2 above, x40 in total
        out["fc_alpha.bias"] = out["fc_alpha.bias"] + 5.0
    return out


# --------------------------------------------------------------------------------------
# a1: rays
def ray_bundle(height: int, width: int, intrinsics: Sequence[float], pose: Tensor) -> Tuple[Tensor, Tensor]:
    """Per-pixel origin/direction, row-major [H, W, 3]; directions are NOT normalised."""
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    col = torch.arange(width, dtype=pose.dtype).view(1, width).expand(height, width)
    row = torch.arange(height, dtype=pose.dtype).view(height, 1).expand(height, width)
    cam = torch.stack(((col - width * cx) / fx, -(row - height * cy) / fy, -torch.ones_like(col)), dim=-1)
    rot = pose[:3, :3]
    rd = (cam[..., None, :] * rot).sum(dim=-1)
    ro = pose[:3, -1].expand(rd.shape)
    return ro, rd


# a5: positional encoding
def posenc(x: Tensor, num_freqs: int, include_input: bool) -> Tensor:
    parts = [x] if include_input else []
    freqs = 2.0 ** torch.linspace(0.0, num_freqs - 1, num_freqs, dtype=x.dtype)
    for f in freqs:
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)


# a6: the conditional MLP
def mlp_forward(p: Dict[str, Tensor], x: Tensor, expr: Tensor, latent: Tensor) -> Tensor:
    """x: [rows, 63+24] -> [rows, 4] = (rgb raw, sigma raw)."""
    lin = torch.nn.functional.linear
    xyz, dirs = x[..., :DIM_XYZ], x[..., DIM_XYZ:]
    rows = xyz.shape[0]
    cond = torch.cat(((expr * 1 / 3).repeat(rows, 1), latent.repeat(rows, 1)), dim=1)
    initial = torch.cat((xyz, cond), dim=1)
    h = initial
    for i in range(6):
        inp = torch.cat((initial, h), dim=-1) if i == 3 else h
        h = torch.relu(lin(inp, p[f"layers_xyz.{i}.weight"], p[f"layers_xyz.{i}.bias"]))
    feat = lin(h, p["fc_feat.weight"], p["fc_feat.bias"])
    sigma = lin(feat, p["fc_alpha.weight"], p["fc_alpha.bias"])
    g = torch.relu(lin(torch.cat((feat, dirs), dim=-1), p["layers_dir.0.weight"], p["layers_dir.0.bias"]))
    for i in (1, 2):  # layers_dir.3 exists in the state_dict but is never applied
        g = torch.relu(lin(g, p[f"layers_dir.{i}.weight"], p[f"layers_dir.{i}.bias"]))
    rgb = lin(g, p["fc_rgb.weight"], p["fc_rgb.bias"])
    return torch.cat((rgb, sigma), dim=-1)


def mlp_activations(p: Dict[str, Tensor], x: Tensor, expr: Tensor, latent: Tensor) -> List[Tensor]:
    """Post-activation tensors in the order of the kernel's tensor-core steps 0..8 (nfb_layout.h): the six
    layers_xyz outputs, then the three layers_dir outputs.  Test aid for the layer probe (NfbDebug.act_dump)."""
    lin = torch.nn.functional.linear
    xyz, dirs = x[..., :DIM_XYZ], x[..., DIM_XYZ:]
    rows = xyz.shape[0]
    cond = torch.cat(((expr * 1 / 3).repeat(rows, 1), latent.repeat(rows, 1)), dim=1)
    initial = torch.cat((xyz, cond), dim=1)
    h, acts = initial, []
    for i in range(6):
        inp = torch.cat((initial, h), dim=-1) if i == 3 else h
        h = torch.relu(lin(inp, p[f"layers_xyz.{i}.weight"], p[f"layers_xyz.{i}.bias"]))
        acts.append(h)
    feat = lin(h, p["fc_feat.weight"], p["fc_feat.bias"])
    g = torch.relu(lin(torch.cat((feat, dirs), dim=-1), p["layers_dir.0.weight"], p["layers_dir.0.bias"]))
    acts.append(g)
    for i in (1, 2):
        g = torch.relu(lin(g, p[f"layers_dir.{i}.weight"], p[f"layers_dir.{i}.bias"]))
        acts.append(g)
    return acts


# a8 + a7: compositing
def exclusive_cumprod(t: Tensor) -> Tensor:
    c = torch.cumprod(t, dim=-1)
    c = torch.roll(c, 1, dims=-1)
    c[..., 0] = 1.0
    return c


def composite(raw: Tensor, z: Tensor, rd: Tensor, noise_std: float = 0.0, noise: Optional[Tensor] = None,
              white_bkgd: bool = False, has_bg: bool = False):
    """raw [N,S,4] (last sample's rgb already overwritten with the background when has_bg)."""
    far_gap = torch.full_like(z[..., :1], 1e10)
    delta = torch.cat((z[..., 1:] - z[..., :-1], far_gap), dim=-1) * rd[..., None, :].norm(p=2, dim=-1)
    if has_bg:
        col = torch.cat((torch.sigmoid(raw[:, :-1, :3]), raw[:, -1:, :3]), dim=1)
    else:
        col = torch.sigmoid(raw[..., :3])
    sig_in = raw[..., 3]
    if noise_std > 0.0:
        sig_in = sig_in + noise * noise_std
    sigma = torch.relu(sig_in).clone()
    sigma[:, -1] += 1e-6
    alpha = 1.0 - torch.exp(-sigma * delta)
    w = alpha * exclusive_cumprod(1.0 - alpha + 1e-10)
    rgb = (w[..., None] * col).sum(dim=-2)
    depth = (w * z).sum(dim=-1)
    acc = w.sum(dim=-1)
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])
    return rgb, disp, acc, w, depth


# a9: inverse-CDF resampling
def resample(bins: Tensor, weights: Tensor, num: int, det: bool, u: Optional[Tensor] = None) -> Tensor:
    weights = weights + 1e-5
    pdf = weights / weights.sum(dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    cdf = torch.cat((torch.zeros_like(cdf[..., :1]), cdf), dim=-1)
    if det:
        u = torch.linspace(0.0, 1.0, steps=num, dtype=weights.dtype).expand(list(cdf.shape[:-1]) + [num])
    u = u.contiguous()
    idx = torch.searchsorted(cdf.contiguous(), u, right=True)
    lo = (idx - 1).clamp(min=0)
    hi = idx.clamp(max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    bin_lo, bin_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    den = cdf_hi - cdf_lo
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    return bin_lo + (u - cdf_lo) / den * (bin_hi - bin_lo)


@dataclass
class Sampling:
    num_coarse: int = 64
    num_fine: int = 128
    perturb: bool = False
    noise_std: float = 0.0
    white_bkgd: bool = False
    chunksize: int = 65536


@dataclass
class Noise:
    """Explicit noise for one ray chunk, in the reference's draw order:
    t_rand[N,Nc] (train_utils.py:75), n_c[N,Nc] (volume_rendering_utils.py:44),
    u[N,Nf] (nerf_helpers.py:363), n_f[N,Nc+Nf]."""
    t_rand: Optional[Tensor] = None
    n_c: Optional[Tensor] = None
    u: Optional[Tensor] = None
    n_f: Optional[Tensor] = None


def draw_noise(n: int, s: Sampling, gen: Optional[torch.Generator] = None) -> Noise:
    """Draw the four tensors exactly when and in the order the reference would."""
    kw = dict(dtype=torch.float32, generator=gen)
    out = Noise()
    if s.perturb:
        out.t_rand = torch.rand((n, s.num_coarse), **kw)
    if s.noise_std > 0.0:
        out.n_c = torch.randn((n, s.num_coarse), **kw)
    if s.num_fine > 0:
        if s.perturb:  # det = (perturb == 0.0)
            out.u = torch.rand((n, s.num_fine), **kw)
        if s.noise_std > 0.0:
            out.n_f = torch.randn((n, s.num_coarse + s.num_fine), **kw)
    return out


def _encode(pts: Tensor, dir_cols: Tensor) -> Tensor:
    flat = pts.reshape(-1, 3)
    dirs = dir_cols[:, None, :].expand(pts.shape).reshape(-1, 3)
    return torch.cat((posenc(flat, 10, True), posenc(dirs, 4, False)), dim=-1)


def _mlp_rows(p, x: Tensor, expr: Tensor, latent: Tensor, rows_per_call: int) -> Tensor:
    """run_network feeds the MLP `chunksize` ROWS at a time (train_utils.py:20); the GEMM batch shape
    changes FP32 rounding at the 1e-5 level on opaque weights, so the oracle keeps the same split."""
    return torch.cat([mlp_forward(p, x[i:i + rows_per_call], expr, latent)
                      for i in range(0, x.shape[0], rows_per_call)], dim=0)


def render_chunk(rays: Tensor, pc: Dict[str, Tensor], pf: Optional[Dict[str, Tensor]], s: Sampling,
                 expr: Tensor, latent: Tensor, bg: Optional[Tensor], noise: Noise,
                 dir_cols: Optional[Tensor] = None, extras: Optional[dict] = None):
    """rays [N,8] = (o, d, near, far).  dir_cols [N,3] is what the direction encoder sees; the
    reference feeds it columns 5..7 of the ray batch = (d_z, near, far) (train_utils.py:14)."""
    n = rays.shape[0]
    ro, rd = rays[:, :3], rays[:, 3:6]
    near, far = rays[:, 6:7], rays[:, 7:8]
    if dir_cols is None:
        dir_cols = rays[:, 5:8]
    t = torch.linspace(0.0, 1.0, s.num_coarse, dtype=rays.dtype)
    z = (near * (1.0 - t) + far * t).expand(n, s.num_coarse)
    if s.perturb:
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat((mid, z[:, -1:]), dim=-1)
        lower = torch.cat((z[:, :1], mid), dim=-1)
        z = lower + (upper - lower) * noise.t_rand
    pts = ro[:, None, :] + rd[:, None, :] * z[:, :, None]
    raw = _mlp_rows(pc, _encode(pts, dir_cols), expr, latent, s.chunksize).reshape(n, s.num_coarse, 4).clone()
    if bg is not None:
        raw[:, -1, :3] = bg
    rgb_c, disp_c, acc_c, w, _ = composite(raw, z, rd, s.noise_std, noise.n_c, s.white_bkgd, bg is not None)
    if extras is not None:
        extras.update(z_coarse=z, raw_coarse=raw, w_coarse=w)
    rgb_f = disp_f = acc_f = None
    if s.num_fine > 0:
        zmid = 0.5 * (z[:, 1:] + z[:, :-1])
        zs = resample(zmid, w[:, 1:-1], s.num_fine, det=not s.perturb, u=noise.u).detach()
        z, _ = torch.sort(torch.cat((z, zs), dim=-1), dim=-1)
        pts = ro[:, None, :] + rd[:, None, :] * z[:, :, None]
        raw = _mlp_rows(pf, _encode(pts, dir_cols), expr, latent, s.chunksize).reshape(n, z.shape[1], 4).clone()
        if bg is not None:
            raw[:, -1, :3] = bg
        rgb_f, disp_f, acc_f, w, _ = composite(raw, z, rd, s.noise_std, noise.n_f, s.white_bkgd, bg is not None)
        if extras is not None:
            extras.update(z_fine=z, raw_fine=raw, w_fine=w, z_samples=zs)
    return rgb_c, disp_c, acc_c, rgb_f, disp_f, acc_f, w[:, -1]


def run_one_iter(ro: Tensor, rd: Tensor, pc, pf, s: Sampling, near: float, far: float, expr: Tensor,
                 latent: Tensor, bg: Optional[Tensor] = None, mode: str = "validation",
                 noise_per_chunk: Optional[List[Noise]] = None, gen: Optional[torch.Generator] = None,
                 rd_ablation: Optional[Tensor] = None):
    """Chunked driver with the reference's return conventions (train_utils.py:165-290)."""
    shape3, shape1 = rd.shape, rd.shape[:-1]
    o, d = ro.reshape(-1, 3), rd.reshape(-1, 3)
    n = d.shape[0]
    rays = torch.cat((o, d, near * torch.ones_like(d[:, :1]), far * torch.ones_like(d[:, :1])), dim=-1)
    abl0 = None
    if rd_ablation is not None:  # every chunk takes chunk 0 of the ablation bundle (train_utils.py:81-82)
        abl0 = rd_ablation.reshape(-1, 3)[: s.chunksize]
    outs = []
    for ci, start in enumerate(range(0, n, s.chunksize)):
        chunk = rays[start:start + s.chunksize]
        noise = noise_per_chunk[ci] if noise_per_chunk is not None else draw_noise(chunk.shape[0], s, gen)
        dir_cols = None
        if abl0 is not None:
            if abl0.shape[0] != chunk.shape[0]:
                raise RuntimeError("ablation chunk 0 and ray chunk differ in length (reference raises here too)")
            dir_cols = torch.cat((abl0[:, 2:3], chunk[:, 6:8]), dim=-1)
        bgc = bg[start:start + s.chunksize] if bg is not None else None
        outs.append(render_chunk(chunk, pc, pf, s, expr, latent, bgc, noise, dir_cols))
    cols = [torch.cat(c, dim=0) if c[0] is not None else None for c in zip(*outs)]
    if mode == "validation":
        shapes = [shape3, shape1, shape1]
        if pf is not None:
            shapes = shapes + shapes + [shape1]
            return tuple(c.view(sh) if c is not None else None for c, sh in zip(cols, shapes))
        return tuple([c.view(sh) for c, sh in zip(cols, shapes)] + [None, None, None])  # 6-tuple quirk
    return tuple(cols)


# --------------------------------------------------------------------------------------
# synthetic inputs
def synthetic_frame(frame_index: int, height: int, width: int):
    """expression -> latent -> pose angles -> background, from Generator(42 + frame_index)."""
    g = torch.Generator().manual_seed(42 + frame_index)
    expr = torch.randn(76, generator=g) * 0.5
    latent = torch.randn(32, generator=g) * 0.1
    ang = (torch.rand(2, generator=g) * 2.0 - 1.0) * math.radians(15.0)
    bg = torch.rand((height, width, 3), generator=g)
    yaw, pitch = float(ang[0]), float(ang[1])
    ry = torch.tensor([[math.cos(yaw), 0.0, math.sin(yaw)], [0.0, 1.0, 0.0], [-math.sin(yaw), 0.0, math.cos(yaw)]])
    rx = torch.tensor([[1.0, 0.0, 0.0], [0.0, math.cos(pitch), -math.sin(pitch)], [0.0, math.sin(pitch), math.cos(pitch)]])
    pose = torch.eye(4)
    pose[:3, :3] = ry @ rx
    pose[:3, 3] = torch.tensor([0.0, 0.0, 0.5])
    intrinsics = [1200.0 * width / 512.0, 1200.0 * height / 512.0, 0.5, 0.5]
    return dict(expr=expr, latent=latent, pose=pose.float(), bg=bg, intrinsics=intrinsics)
