"""Data-parallel helpers for the render path.  Rays are independent, so there is no collective
inside the path: evaluation shards contiguous pixel rows across ranks and all-gathers the output tiles; training
shards the ray batch and all-reduces one flat FP32 gradient bucket.  One process per GPU, torch.distributed
(NCCL on GPUs; the same code runs on gloo/CPU tensors in the tests)."""
import torch
import torch.distributed as dist


def shard_rows(height: int, world: int, rank: int):
    """Contiguous row block [begin, begin+rows) of rank `rank`; the first height % world ranks get one extra row."""
    base, extra = divmod(height, world)
    rows = base + (1 if rank < extra else 0)
    begin = rank * base + min(rank, extra)
    return begin, rows


def gather_rows(local: torch.Tensor, height: int, group=None) -> torch.Tensor:
    """All-gather per-rank row blocks [rows_r, W, C] (ragged when height % world != 0) into [height, W, C]."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    shapes = [shard_rows(height, world, r)[1] for r in range(world)]
    max_rows = max(shapes)
    pad = local
    if local.shape[0] < max_rows:
        pad = torch.cat((local, local.new_zeros((max_rows - local.shape[0],) + tuple(local.shape[1:]))), dim=0)
    out = local.new_empty((world * max_rows,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = [out[r * max_rows:r * max_rows + shapes[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


def render_frame_sharded(render_rows, height: int, group=None) -> torch.Tensor:
    """render_rows(row_begin, rows) -> [rows, W, C] on this rank; returns the assembled [height, W, C] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    begin, rows = shard_rows(height, world, rank)
    local = render_rows(begin, rows)
    return gather_rows(local, height, group) if world > 1 else local


def shard_batch(n: int, world: int, rank: int):
    """Equal contiguous shards of a ray batch (n must divide evenly so that mean-of-means == global mean)."""
    if n % world:
        raise ValueError(f"ray batch of {n} does not split evenly over {world} ranks")
    per = n // world
    return rank * per, per


def allreduce_gradients(params, group=None, average=True):
    """One all-reduce over a single flat FP32 bucket holding every existing .grad (parameters whose grad is None,
    e.g. layers_dir.3, are skipped on every rank alike)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel()
