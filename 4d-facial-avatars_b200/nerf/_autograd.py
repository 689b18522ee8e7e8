"""Gradients for training (train_transformed_rays.py:389 `loss.backward()`).

Forward: the fused sm_100a kernel (same launch as evaluation) — it also returns the sample depths it used.
Backward (interim, round 1): the path is re-evaluated on the device as a differentiable torch graph at exactly those
depths and noise values, and torch.autograd produces the parameter / latent-code gradients (cuBLAS GEMMs).  The
resampled depths carry no gradient, as in the reference (`z_samples.detach()`, train_utils.py:124).  A fused
tcgen05 backward is the planned replacement."""
import torch

from ._engine import PARAM_ORDER


def _posenc(x, n_freq, include_input):
    parts = [x] if include_input else []
    for k in range(n_freq):
        parts += [torch.sin(x * float(2 ** k)), torch.cos(x * float(2 ** k))]
    return torch.cat(parts, dim=-1)


def _mlp(p, x, expr, latent):
    F = torch.nn.functional
    xyz, dirs = x[..., :63], x[..., 63:]
    rows = xyz.shape[0]
    cond = torch.cat(((expr * 1 / 3).reshape(1, -1).expand(rows, -1), latent.reshape(1, -1).expand(rows, -1)), dim=1)
    initial = torch.cat((xyz, cond), dim=1)
    h = initial
    for i in range(6):
        h = F.relu(F.linear(torch.cat((initial, h), dim=-1) if i == 3 else h, p[f"layers_xyz.{i}.weight"], p[f"layers_xyz.{i}.bias"]))
    feat = F.linear(h, p["fc_feat.weight"], p["fc_feat.bias"])
    sigma = F.linear(feat, p["fc_alpha.weight"], p["fc_alpha.bias"])
    g = F.relu(F.linear(torch.cat((feat, dirs), dim=-1), p["layers_dir.0.weight"], p["layers_dir.0.bias"]))
    g = F.relu(F.linear(g, p["layers_dir.1.weight"], p["layers_dir.1.bias"]))
    g = F.relu(F.linear(g, p["layers_dir.2.weight"], p["layers_dir.2.bias"]))
    return torch.cat((F.linear(g, p["fc_rgb.weight"], p["fc_rgb.bias"]), sigma), dim=-1)


def _composite(raw, z, rd, noise_std, noise, white_bkgd, bg):
    n, s = z.shape
    delta = torch.cat((z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], 1e10)), dim=-1) * rd.norm(p=2, dim=-1, keepdim=True)
    col = torch.sigmoid(raw[..., :3])
    if bg is not None:
        col = torch.cat((col[:, :-1], bg[:, None, :]), dim=1)
    sig_in = raw[..., 3]
    if noise_std > 0.0:
        sig_in = sig_in + noise * noise_std
    last = torch.zeros(s, device=z.device, dtype=z.dtype)
    last[-1] = 1e-6
    sigma = torch.relu(sig_in) + last
    alpha = 1.0 - torch.exp(-sigma * delta)
    trans = torch.cumprod(1.0 - alpha + 1e-10, dim=-1)
    trans = torch.cat((torch.ones_like(trans[:, :1]), trans[:, :-1]), dim=-1)
    w = alpha * trans
    rgb = (w[..., None] * col).sum(dim=-2)
    depth = (w * z).sum(dim=-1)
    acc = w.sum(dim=-1)
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])
    return rgb, disp, acc, w


def _pass(p, z, rays, dir_cols, expr, latent, noise_std, noise, white_bkgd, bg):
    ro, rd = rays[:, :3], rays[:, 3:6]
    n, s = z.shape
    pts = ro[:, None, :] + rd[:, None, :] * z[:, :, None]
    x = torch.cat((_posenc(pts.reshape(-1, 3), 10, True), _posenc(dir_cols[:, None, :].expand(n, s, 3).reshape(-1, 3), 4, False)), dim=-1)
    raw = _mlp(p, x, expr, latent).reshape(n, s, 4)
    return _composite(raw, z, rd, noise_std, noise, white_bkgd, bg)


class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, rays, args, has_fine, expr, latent, *params):
        out = eng.render(rays[:, :3], rays[:, 3:6], debug=True, **args)
        ctx.eng_args = (args, has_fine)
        ctx.save_for_backward(rays, expr, latent, out["z_coarse"], out["z_fine"] if has_fine else rays.new_zeros(1), *params)
        ctx.set_materialize_grads(False)
        res = (out["rgb_coarse"], out["disp_coarse"], out["acc_coarse"],
               out.get("rgb_fine"), out.get("disp_fine"), out.get("acc_fine"), out["w_last"])
        ctx.n_out = [r is not None for r in res]
        return tuple(r if r is not None else rays.new_zeros(0) for r in res)

    @staticmethod
    def backward(ctx, *gouts):
        args, has_fine = ctx.eng_args
        rays, expr, latent, z_c, z_f, *params = ctx.saved_tensors
        npar = len(PARAM_ORDER)
        with torch.enable_grad():
            leaves = [t.detach().requires_grad_(True) for t in params]
            lat = latent.detach().requires_grad_(True)
            pc = dict(zip(PARAM_ORDER, leaves[:npar]))
            pf = dict(zip(PARAM_ORDER, leaves[npar:2 * npar])) if has_fine else None
            dir_cols = torch.cat((args["dir_z"].reshape(-1, 1) if args.get("dir_z") is not None else rays[:, 5:6],
                                  torch.full_like(rays[:, :1], args["near"]), torch.full_like(rays[:, :1], args["far"])), dim=-1)
            nz = args.get("noise") or {}
            bg = args.get("background")
            rgb_c, disp_c, acc_c, w = _pass(pc, z_c.detach(), rays, dir_cols, expr, lat, args["noise_std"], nz.get("n_c"),
                                            args["white_bkgd"], bg)
            outs = [rgb_c, disp_c, acc_c, None, None, None, w[:, -1]]
            if has_fine:
                rgb_f, disp_f, acc_f, w = _pass(pf, z_f.detach(), rays, dir_cols, expr, lat, args["noise_std"], nz.get("n_f"),
                                                args["white_bkgd"], bg)
                outs[3:7] = [rgb_f, disp_f, acc_f, w[:, -1]]
            pairs = [(o, g) for o, g in zip(outs, gouts) if o is not None and g is not None]
            if not pairs:
                return (None,) * (6 + len(params))
            used = [t for t in leaves + [lat]]
            grads = torch.autograd.grad([o for o, _ in pairs], used, [g for _, g in pairs], allow_unused=True)
        gpar, glat = grads[:-1], grads[-1]
        return (None, None, None, None, None, glat) + tuple(gpar)


def render_with_grad(eng, rays, model_coarse, model_fine, expressions, latent_code, args):
    sd_c = dict(model_coarse.named_parameters())
    params = [sd_c[k] for k in PARAM_ORDER]
    has_fine = model_fine is not None
    if has_fine:
        sd_f = dict(model_fine.named_parameters())
        params += [sd_f[k] for k in PARAM_ORDER]
    res = _RenderFn.apply(eng, rays, args, has_fine, expressions, latent_code, *params)
    return tuple(r if r.numel() > 0 else None for r in res)
