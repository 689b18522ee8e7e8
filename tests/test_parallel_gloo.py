"""N>1 host logic on CPU: world_size-2 gloo processes exercise row sharding + all-gather assembly and the flat
gradient-bucket all-reduce (the GPU path uses the same functions over NCCL)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, height, q):
    sys.path.insert(0, os.path.join(ROOT, "4d-facial-avatars_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf import parallel
    W = 5

    def fake_render(begin, rows):  # pixel value = global pixel index, 3 channels
        idx = torch.arange(begin * W, (begin + rows) * W, dtype=torch.float32).view(rows, W, 1)
        return idx.expand(rows, W, 3).contiguous()

    frame = parallel.render_frame_sharded(fake_render, height)
    ok_frame = torch.equal(frame, torch.arange(height * W, dtype=torch.float32).view(height, W, 1).expand(height, W, 3))
    # gradient bucket: rank r holds grads filled with r+1; one parameter without grad
    ps = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2))]
    ps[0].grad = torch.full((3, 4), float(rank + 1))
    ps[1].grad = torch.full((7,), float(10 * (rank + 1)))
    n = parallel.allreduce_gradients(ps, average=True)
    mean = sum(range(1, world + 1)) / world
    ok_grad = (n == 19 and torch.allclose(ps[0].grad, torch.full((3, 4), mean)) and
               torch.allclose(ps[1].grad, torch.full((7,), 10 * mean)) and ps[2].grad is None)
    begin, per = parallel.shard_batch(2048, world, rank)
    q.put((rank, ok_frame, ok_grad, begin, per))
    dist.destroy_process_group()


def _run(height):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, height, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_gloo_world2_even_rows():
    res = _run(8)
    assert all(r[1] and r[2] for r in res)
    assert [(r[3], r[4]) for r in res] == [(0, 1024), (1024, 1024)]


def test_gloo_world2_ragged_rows():
    res = _run(7)  # 4 + 3 rows
    assert all(r[1] and r[2] for r in res)


def test_shard_rows_cover_exactly():
    sys.path.insert(0, os.path.join(ROOT, "4d-facial-avatars_b200"))
    from nerf import parallel
    for h in (1, 7, 512, 1024):
        for w in (1, 2, 3, 4, 8):
            blocks = [parallel.shard_rows(h, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and sum(b[1] for b in blocks) == h
            assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
