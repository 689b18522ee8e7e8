"""Run one of the reference's CLI scripts (train_transformed_rays.py / eval_transformed_rays.py) UNMODIFIED with the
drop-in `nerf` package of this repo first on sys.path.  Provides minimal stand-ins for `imageio` and
`matplotlib` when those packages are absent (the scripts import them before `nerf`).

    python 4d-facial-avatars_b200/run_reference_script.py <script.py> [script args...]

Data-parallel, one process per GPU of one node, the script body still unmodified:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        4d-facial-avatars_b200/run_reference_script.py <script.py> [script args...]

Under torchrun (WORLD_SIZE > 1) the launcher initialises NCCL, pins the rank's GPU, replaces `nerf.run_one_iter_of_nerf` by
its ray-sharding wrapper (nerf/parallel.py: data_parallel), averages the parameter gradients in a flat bucket before
every optimizer step (a global optimizer pre-step hook), and lets only rank 0 write images / checkpoints / summaries.
"""
import os
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def _ensure_imageio():
    try:
        import imageio  # noqa: F401
        return
    except ImportError:
        pass
    import numpy as np
    m = types.ModuleType("imageio")

    def imread(path):
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(path)
        if img.ndim == 3:
            img = img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
        return img

    def imwrite(path, img):
        import cv2
        img = np.asarray(img)
        if img.ndim == 3 and img.shape[2] >= 3:
            img = img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
        cv2.imwrite(path, img)

    m.imread, m.imwrite, m.imsave = imread, imwrite, imwrite
    sys.modules["imageio"] = m


def _ensure_matplotlib():
    try:
        import matplotlib  # noqa: F401
        return
    except ImportError:
        pass

    class _Anything:
        def __call__(self, *a, **k):
            return self

        def __getattr__(self, name):
            return self

    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.use = lambda *a, **k: None
    for name in ("figure", "savefig", "close", "imshow", "show", "plot", "axis", "subplots", "Axes"):
        setattr(plt, name, _Anything())
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt


def _enable_data_parallel():
    """One process per GPU under torchrun: shard rays inside run_one_iter_of_nerf, average gradients before optimizer steps,
    rank-0-only output files."""
    import torch
    import torch.distributed as dist
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import nerf
    from nerf import parallel
    nerf.run_one_iter_of_nerf = parallel.data_parallel(nerf.run_one_iter_of_nerf)
    nerf.train_utils.run_one_iter_of_nerf = nerf.run_one_iter_of_nerf

    def _avg_grads(optimizer, args, kwargs):
        params = [p for g in optimizer.param_groups for p in g["params"]]
        parallel.allreduce_gradients(params, average=True)
    from torch.optim.optimizer import register_optimizer_step_pre_hook
    register_optimizer_step_pre_hook(_avg_grads)
    if rank != 0:  # only rank 0 writes files
        import imageio
        imageio.imwrite = imageio.imsave = lambda *a, **k: None
        torch.save = lambda *a, **k: None
        try:
            from torch.utils import tensorboard
            for name in ("add_scalar", "add_image", "add_images", "add_histogram"):
                setattr(tensorboard.SummaryWriter, name, lambda *a, **k: None)
        except Exception:
            pass
    return world


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = os.path.abspath(sys.argv[1])
    _ensure_imageio()
    _ensure_matplotlib()
    sys.path.insert(0, HERE)  # the drop-in `nerf` wins over the one next to the script
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        _enable_data_parallel()
    sys.argv = [script] + sys.argv[2:]
    os.chdir(os.path.dirname(script))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
