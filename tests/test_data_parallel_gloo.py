"""nerf.parallel.data_parallel on CPU (gloo, world_size 2) around a small differentiable stand-in with the signature of
run_one_iter_of_nerf: validation frames are assembled from row shards, and a training step's parameter / latent gradients
after the averaging all-reduce equal the single-process full-batch gradients — including the ray-independent regulariser."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_run(height, width, focal_length, model_coarse, model_fine, ray_origins, ray_directions, options, mode="train",
              encode_position_fn=None, encode_direction_fn=None, expressions=None, background_prior=None, latent_code=None,
              ray_directions_ablation=None):
    shape1 = ray_directions.shape[:-1]
    ro, rd = ray_origins.reshape(-1, 3), ray_directions.reshape(-1, 3)
    x = torch.cat((ro, rd, latent_code.reshape(1, -1).expand(ro.shape[0], -1)), dim=-1)
    hc, hf = torch.tanh(model_coarse(x)), torch.tanh(model_fine(x))
    if background_prior is not None:
        hc = hc + 0.1 * torch.cat((background_prior.reshape(-1, 3), background_prior.reshape(-1, 3)[:, :2]), dim=-1)
    outs = [hc[:, :3], hc[:, 3], hc[:, 4], hf[:, :3], hf[:, 3], hf[:, 4], hf[:, 4] * 0.5]
    if mode == "validation":
        outs = [o.reshape(shape1 + o.shape[1:]) for o in outs]
    return tuple(outs)


def _models():
    torch.manual_seed(0)
    return torch.nn.Linear(6 + 4, 5), torch.nn.Linear(6 + 4, 5)


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "4d-facial-avatars_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf import parallel
    run = parallel.data_parallel(_fake_run)
    g = torch.Generator().manual_seed(1)
    mc, mf = _models()
    lat = torch.randn(4, generator=g).requires_grad_(True)
    # ---- validation: ragged rows (7 rows over 2 ranks)
    H, W = 7, 5
    ro, rd = torch.randn(H, W, 3, generator=g), torch.randn(H, W, 3, generator=g)
    bg = torch.rand(H * W, 3, generator=g)
    with torch.no_grad():
        got = run(H, W, 1.0, mc, mf, ro, rd, None, mode="validation", background_prior=bg, latent_code=lat)
        ref = _fake_run(H, W, 1.0, mc, mf, ro, rd, None, mode="validation", background_prior=bg, latent_code=lat)
    ok_val = all(torch.allclose(a, b) and a.shape == b.shape for a, b in zip(got, ref))
    # ---- train: 16 rays, loss over the FULL batch as the unmodified script computes it
    n = 16
    ro, rd = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
    bg = torch.rand(n, 3, generator=g)
    tgt = torch.rand(n, 3, generator=g)

    def loss_of(outs):
        return ((outs[0] - tgt) ** 2).mean() + ((outs[3] - tgt) ** 2).mean() + 0.005 * lat.norm()
    params = list(mc.parameters()) + list(mf.parameters()) + [lat]
    loss_ref = loss_of(_fake_run(0, 0, 1.0, mc, mf, ro, rd, None, mode="train", background_prior=bg, latent_code=lat))
    g_ref = torch.autograd.grad(loss_ref, params)
    outs = run(0, 0, 1.0, mc, mf, ro, rd, None, mode="train", background_prior=bg, latent_code=lat)
    loss = loss_of(outs)
    loss.backward()
    parallel.allreduce_gradients(params, average=True)
    ok_loss = abs(float(loss) - float(loss_ref)) < 1e-6 and all(o.shape[0] == n for o in outs)
    ok_grad = all(torch.allclose(p.grad, gr, atol=1e-6, rtol=1e-5) for p, gr in zip(params, g_ref))
    q.put((rank, ok_val, ok_loss, ok_grad))
    dist.destroy_process_group()


def test_data_parallel_wrapper_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] and r[3] for r in res), res


def test_world_size_one_is_a_passthrough():
    sys.path.insert(0, os.path.join(ROOT, "4d-facial-avatars_b200"))
    from nerf import parallel
    mc, mf = _models()
    lat = torch.zeros(4)
    ro, rd = torch.randn(6, 3), torch.randn(6, 3)
    a = parallel.data_parallel(_fake_run)(0, 0, 1.0, mc, mf, ro, rd, None, mode="train", latent_code=lat)
    b = _fake_run(0, 0, 1.0, mc, mf, ro, rd, None, mode="train", latent_code=lat)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
