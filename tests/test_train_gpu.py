"""Training path: fused forward + fused backward (nfb_render_backward) through the drop-in API, against autograd through
the CPU oracle (same noise, same depths).  The backward carries gradients and activations as FP16 tensor-core operands
(2^-11 relative each) with FP32 accumulation, so parameter gradients are compared at 1e-2 of each tensor's largest
magnitude (measured: typically 3e-4 .. 2e-3; tests/test_backward_gpu.py checks every stage separately)."""
import pytest
import torch

import nerface_oracle as O

pytestmark = pytest.mark.gpu


def _models(nerf, dev, stress=False):
    out = []
    for seed in (100, 101):
        m = nerf.models.ConditionalBlendshapePaperNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4,
                                                            include_input_xyz=True, include_input_dir=False)
        m.load_state_dict(O.random_init_params(seed, stress))
        out.append(m.to(dev))
    return out


@pytest.mark.parametrize("stress", [False, True])
def test_gradients_match_oracle_autograd(built_lib, stress, monkeypatch):
    import nerf
    from nerf import train_utils
    dev = torch.device("cuda", 0)
    nerf.set_precision("exact")
    n, nc, nf = 48, 64, 64
    fr = O.synthetic_frame(21, 6, 8)
    ro, rd = O.ray_bundle(6, 8, fr["intrinsics"], fr["pose"])
    ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()
    bg = fr["bg"].reshape(-1, 3)
    target = torch.rand(n, 3, generator=torch.Generator().manual_seed(5))
    s = O.Sampling(nc, nf, True, 0.1, False, 2048)
    noise = O.draw_noise(n, s, torch.Generator().manual_seed(77))
    monkeypatch.setattr(train_utils, "_draw_noise", lambda m, opts, device, has_fine: {
        "t_rand": noise.t_rand.to(device), "n_c": noise.n_c.to(device), "u": noise.u.to(device), "n_f": noise.n_f.to(device)})

    # --- oracle: CPU autograd
    pc = {k: v.clone().requires_grad_(True) for k, v in O.random_init_params(100, stress).items()}
    pf = {k: v.clone().requires_grad_(True) for k, v in O.random_init_params(101, stress).items()}
    latent_codes = torch.zeros(4, 32)
    latent_codes[2] = fr["latent"]
    lat_ref = latent_codes.clone().requires_grad_(True)
    rays = torch.cat((ro, rd, torch.full((n, 1), 0.2), torch.full((n, 1), 0.8)), dim=-1)
    ref = O.render_chunk(rays, pc, pf, s, fr["expr"], lat_ref[2], bg, noise)
    loss_ref = ((ref[0] - target) ** 2).mean() + ((ref[3] - target) ** 2).mean() + 0.005 * lat_ref[2].norm()
    loss_ref.backward()

    # --- ours: fused forward on the GPU through the drop-in API, gradients via loss.backward()
    mc, mf = _models(nerf, dev, stress)
    lat = latent_codes.clone().to(dev).requires_grad_(True)
    blk = dict(num_coarse=nc, num_fine=nf, perturb=True, lindisp=False, radiance_field_noise_std=0.1,
               white_background=False, chunksize=2048)
    cfg = nerf.CfgNode(dict(nerf=dict(use_viewdirs=True, train=blk), dataset=dict(no_ndc=True, near=0.2, far=0.8)))
    out = nerf.run_one_iter_of_nerf(6, 8, fr["intrinsics"], mc, mf, ro.to(dev), rd.to(dev), cfg, mode="train",
                                    expressions=fr["expr"].to(dev), background_prior=bg.to(dev), latent_code=lat[2])
    tgt = target.to(dev)
    loss = ((out[0] - tgt) ** 2).mean() + ((out[3] - tgt) ** 2).mean() + 0.005 * lat[2].norm()
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss_ref)) < 1e-5
    assert mc.layers_dir[3].weight.grad is None and mf.layers_dir[3].bias.grad is None  # unused layer: no grad, as in the reference
    worst = 0.0
    for model, ref_p in ((mc, pc), (mf, pf)):
        for k, p in model.named_parameters():
            if k.startswith("layers_dir.3"):
                continue
            g_ref = ref_p[k].grad
            err = float((p.grad.cpu() - g_ref).abs().max())
            scale = float(g_ref.abs().max()) + 1e-12
            worst = max(worst, err / scale)
            assert err <= 1e-2 * scale + 1e-7, (k, err, scale)
    g_lat = lat.grad.cpu()
    assert float(g_lat[[0, 1, 3]].abs().max()) == 0.0  # only the indexed row receives gradient
    assert float((g_lat[2] - lat_ref.grad[2]).abs().max()) <= 1e-2 * float(lat_ref.grad[2].abs().max()) + 1e-7
    print(f"stress={stress}: worst relative gradient error {worst:.2e}")
    nerf.set_precision("fast")


def test_adam_step_changes_outputs_and_repacks(built_lib):
    """The packed weight streams must follow optimizer updates (parameter _version tracking)."""
    import nerf
    dev = torch.device("cuda", 0)
    mc, mf = _models(nerf, dev)
    fr = O.synthetic_frame(3, 4, 8)
    ro, rd = O.ray_bundle(4, 8, fr["intrinsics"], fr["pose"])
    blk = dict(num_coarse=64, num_fine=64, perturb=False, lindisp=False, radiance_field_noise_std=0.0,
               white_background=False, chunksize=2048)
    cfg = nerf.CfgNode(dict(nerf=dict(use_viewdirs=True, train=blk, validation=blk), dataset=dict(no_ndc=True, near=0.2, far=0.8)))
    lat = torch.zeros(32, device=dev, requires_grad=True)
    opt = torch.optim.Adam(list(mc.parameters()) + list(mf.parameters()) + [lat], lr=5e-4)
    kw = dict(expressions=fr["expr"].to(dev), background_prior=fr["bg"].reshape(-1, 3).to(dev))
    tgt = torch.rand(32, 3, device=dev)
    losses = []
    for _ in range(3):
        out = nerf.run_one_iter_of_nerf(4, 8, fr["intrinsics"], mc, mf, ro.reshape(-1, 3).to(dev), rd.reshape(-1, 3).to(dev), cfg,
                                        mode="train", latent_code=lat, **kw)
        loss = ((out[0] - tgt) ** 2).mean() + ((out[3] - tgt) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[2] < losses[0]
    with torch.no_grad():  # evaluation after training uses the updated weights
        got = nerf.run_one_iter_of_nerf(4, 8, fr["intrinsics"], mc, mf, ro.to(dev), rd.to(dev), cfg, mode="validation", latent_code=lat, **kw)
        pc = {k: v.detach().cpu() for k, v in mc.state_dict().items()}
        pf = {k: v.detach().cpu() for k, v in mf.state_dict().items()}
        ref = O.run_one_iter(ro, rd, pc, pf, O.Sampling(64, 64), 0.2, 0.8, fr["expr"], lat.detach().cpu(), fr["bg"].reshape(-1, 3), "validation")
    assert float((got[3].cpu() - ref[3]).abs().max()) < 1e-4
