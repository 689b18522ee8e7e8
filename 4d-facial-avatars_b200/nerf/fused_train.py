"""Fused training step: what train_transformed_rays.py:336-400 does per iteration — render in train
mode, mse(rgb_coarse) + mse(rgb_fine) + 10 * 0.0005 * ||latent||, loss.backward(), optimizer.step(), zero_grad(), LR decay —
as kernel launches of libnfb on ONE flat FP32 parameter bucket, with no torch.autograd graph, no torch.optim and no
gradient copies:

    set_frame (1 launch) -> training forward (1) -> loss gradient (1) -> backward (writes into the flat gradient bucket)
    -> [one NCCL all-reduce of that bucket when the batch is sharded over ranks] -> Adam + zero_grad (1) -> re-pack (2: fold, pack)

The models keep their reference `state_dict` (their parameters become views of the bucket), so checkpoints, `.parameters()`
and the drop-in `run_one_iter_of_nerf` keep working on the same objects.  torch is used for memory, the noise draws (in the
reference's order) and torch.distributed."""
import torch
import torch.distributed as dist

from . import _engine
from ._engine import PARAM_ORDER


class FusedTrainer:
    def __init__(self, model_coarse, model_fine, n_latent, lr=5e-4, lr_decay_steps=250000, lr_decay_factor=0.1,
                 betas=(0.9, 0.999), eps=1e-8, num_coarse=64, num_fine=64, perturb=True, noise_std=0.1, near=0.2, far=0.8,
                 latent_reg=0.005, white_bkgd=False, latent_codes=None, precision=None):
        dev = next(model_coarse.parameters()).device
        self.eng = _engine.renderer_for(dev)
        self.dev, self.mc, self.mf = dev, model_coarse, model_fine
        self.lr0, self.decay_steps, self.decay_factor = float(lr), float(lr_decay_steps), float(lr_decay_factor)
        self.betas, self.eps = betas, eps
        self.opts = dict(near=float(near), far=float(far), num_coarse=int(num_coarse), num_fine=int(num_fine) if model_fine is not None else 0,
                         perturb=bool(perturb), noise_std=float(noise_std), white_bkgd=bool(white_bkgd), precision=precision)
        self.latent_reg = float(latent_reg)
        self.iter = 0  # optimizer steps taken so far

        # ---- flat bucket: [coarse 26 tensors | fine 26 tensors | pad to 256 | latent table n_latent x 32]
        models = [model_coarse] + ([model_fine] if model_fine is not None else [])
        tensors = [dict(m.named_parameters())[k] for m in models for k in PARAM_ORDER]
        n_mlp = sum(t.numel() for t in tensors)
        self.lat_off = (n_mlp + 255) // 256 * 256
        n_total = self.lat_off + n_latent * 32
        self.params = torch.zeros(n_total, device=dev, dtype=torch.float32)
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        off = 0
        self._views, self._gviews = [], []
        for t in tensors:
            n = t.numel()
            view = self.params[off:off + n].view(t.shape)
            view.copy_(t.detach().to(device=dev, dtype=torch.float32))
            t.data = view  # the module's parameter now IS a slice of the bucket
            self._views.append(view)
            self._gviews.append(self.grads[off:off + n].view(t.shape))
            off += n
        self.latent_codes = self.params[self.lat_off:].view(n_latent, 32)
        if latent_codes is not None:
            self.latent_codes.copy_(latent_codes.detach().to(dev))
        npar = len(PARAM_ORDER)
        skip = [k.startswith("layers_dir.3") for k in PARAM_ORDER]  # unused by the forward (models.py:257): no gradient
        self._pc, self._pf = self._views[:npar], (self._views[npar:] if model_fine is not None else None)
        self._gc = [None if s else g for s, g in zip(skip, self._gviews[:npar])]
        self._gf = [None if s else g for s, g in zip(skip, self._gviews[npar:])] if model_fine is not None else None
        self.loss = torch.zeros(4, device=dev, dtype=torch.float32)
        self._g_rgb = {}
        self._own_engine()

    def _own_engine(self):
        """The device's renderer holds ONE set of packed weights: (re-)pack this trainer's if someone else's are in place (another
        trainer, or other models rendered through the drop-in API since the last step)."""
        if self.eng.packed_owner is not self:
            self.eng.repack(self._pc, self._pf)
            self.eng.mark_synced(self.mc, self.mf)
            self.eng.packed_owner = self

    def lr(self):
        """Learning rate of step number self.iter (1-based).  The reference assigns lr0 * factor ** (i / decay) AFTER the
        optimizer step of loop index i (train_transformed_rays.py:393-399), so loop index i >= 1 runs at exponent (i - 1) / decay
        and loop index 0 at lr0."""
        i = self.iter - 1  # the reference's loop index of the step being taken
        return self.lr0 if i <= 0 else self.lr0 * self.decay_factor ** ((i - 1) / self.decay_steps)

    def _draw_noise(self, n):
        """rand[N,Nc], randn[N,Nc], rand[N,Nf], randn[N,Nc+Nf] — the reference's draw order for one chunk (train chunksize =
        num_random_rays in the shipped YAML, so a batch is one chunk)."""
        o, kw = self.opts, dict(device=self.dev, dtype=torch.float32)
        nc, nf = o["num_coarse"], o["num_fine"]
        out = dict(t_rand=None, n_c=None, u=None, n_f=None)
        if o["perturb"]:
            out["t_rand"] = torch.rand((n, nc), **kw)
        if o["noise_std"] > 0.0:
            out["n_c"] = torch.randn((n, nc), **kw)
        if nf > 0:
            if o["perturb"]:
                out["u"] = torch.rand((n, nf), **kw)
            if o["noise_std"] > 0.0:
                out["n_f"] = torch.randn((n, nc + nf), **kw)
        return out if (o["perturb"] or o["noise_std"] > 0.0) else None

    def gradients(self, ray_origins, ray_directions, target, expressions, latent_index, background=None, world=1, n_total=None,
                  noise=None, events=None, group=None):
        """Forward, loss and backward of this rank's rays ([n,3] CUDA tensors) into the flat gradient bucket; world > 1: the rays
        are one of `world` equal shards of a batch of n_total rays and the bucket is SUM-all-reduced (one collective), after
        which every rank holds the whole batch's gradient.  Returns the device tensor [mse_coarse, mse_fine] of THIS shard's
        share (sum over ranks = batch loss).  `events`: optional (before_collective, after_collective) CUDA events."""
        eng, o = self.eng, self.opts
        self._own_engine()
        n = ray_origins.shape[0]
        n_total = n * world if n_total is None else n_total
        row = self.latent_codes[latent_index]
        eng.set_frame(expressions, row)
        if noise is None:
            noise = self._draw_noise(n)
        out = eng.render(ray_origins, ray_directions, o["near"], o["far"], o["num_coarse"], o["num_fine"], perturb=o["perturb"],
                         noise_std=o["noise_std"], white_bkgd=o["white_bkgd"], background=background, noise=noise,
                         precision=o["precision"], train=True)
        g = self._g_rgb.get(n)
        if g is None:
            g = self._g_rgb[n] = (torch.empty((n, 3), device=self.dev), torch.empty((n, 3), device=self.dev))
        self.loss.zero_()
        has_fine = o["num_fine"] > 0
        eng.loss_mse_grad(out["rgb_coarse"], out["rgb_fine"] if has_fine else None, _engine._f32c(target, self.dev), n_total,
                          g[0], g[1] if has_fine else None, self.loss)
        glat = self.grads[self.lat_off + 32 * latent_index:self.lat_off + 32 * latent_index + 32]
        eng.backward_into((g[0], None, None, g[1] if has_fine else None, None, None, None), self._pc, self._pf, self._gc, self._gf, glat)
        if events is not None:
            events[0].record()
        if world > 1:
            dist.all_reduce(self.grads, group=group)  # ONE collective over the flat bucket (sum: the loss is pre-divided by n_total)
        if events is not None:
            events[1].record()
        self._reg_row = latent_index
        return self.loss[:2]

    def update(self):
        """Adam over the bucket (+ the latent regulariser's gradient on the last frame's row, + zero_grad), then the re-pack."""
        eng = self.eng
        self.iter += 1
        eng.adam_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.lr(), self.iter, self.betas, self.eps,
                      reg_offset=self.lat_off + 32 * self._reg_row if self.latent_reg > 0.0 else -1, reg_weight=self.latent_reg)
        eng.repack(self._pc, self._pf)
        eng.mark_synced(self.mc, self.mf)
        eng.packed_owner = self

    # ---- the whole iteration as ONE CUDA graph (launch-bound at small per-rank batches: ~20 kernels of 3-800 us)
    def capture(self, n, has_background=True, world=1, n_total=None, group=None):
        """Capture gradients() + update() for batches of exactly n rays on this rank into a CUDA graph.  Everything that changes
        from step to step lives in device memory: the inputs (static buffers filled by step_graph), the latent row index, and the
        optimizer's step counter / learning rate (nfb_adam_step_dev).  The noise is drawn inside the graph (torch's graph-safe
        Philox state), in the reference's order.  With world > 1 the NCCL all-reduce of the flat bucket is part of the graph."""
        import ctypes as C
        from . import _capi as capi
        dev, eng, o = self.dev, self.eng, self.opts
        n_total = n * world if n_total is None else n_total
        z = lambda *shape, dt=torch.float32: torch.zeros(shape, device=dev, dtype=dt)  # noqa: E731
        sb = dict(ro=z(n, 3), rd=z(n, 3), tgt=z(n, 3), bg=z(n, 3) if has_background else None, expr=z(76), idx=z(1, dt=torch.int64),
                  lat=z(32), glat=z(32), g0=z(n, 3), g1=z(n, 3))
        sb["rd"][:, 2] = -1.0  # a valid ray for the warm-up
        st = capi.NfbAdamDev(step=self.iter, pad=0, lr0=self.lr0, decay_factor=self.decay_factor, decay_steps=self.decay_steps,
                             beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, grad_scale=1.0, reg_weight=self.latent_reg,
                             table_offset=self.lat_off if self.latent_reg > 0.0 else -1, row=sb["idx"].data_ptr(),
                             lr_over_bc1=0.0, sqrt_bc2=1.0, reg_offset=-1)
        sb["adam"] = torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8).to(dev)
        has_fine = o["num_fine"] > 0
        table_grads = self.grads[self.lat_off:].view(-1, 32)

        def forward_backward():
            sb["lat"].copy_(self.latent_codes.index_select(0, sb["idx"])[0])
            eng.set_frame(sb["expr"], sb["lat"])
            out = eng.render(sb["ro"], sb["rd"], o["near"], o["far"], o["num_coarse"], o["num_fine"], perturb=o["perturb"],
                             noise_std=o["noise_std"], white_bkgd=o["white_bkgd"], background=sb["bg"], noise=self._draw_noise(n),
                             precision=o["precision"], train=True)
            self.loss.zero_()
            eng.loss_mse_grad(out["rgb_coarse"], out["rgb_fine"] if has_fine else None, sb["tgt"], n_total, sb["g0"],
                              sb["g1"] if has_fine else None, self.loss)
            eng.backward_into((sb["g0"], None, None, sb["g1"] if has_fine else None, None, None, None), self._pc, self._pf, self._gc,
                              self._gf, sb["glat"])
            table_grads.index_add_(0, sb["idx"], sb["glat"][None])
            return out

        self._own_engine()
        forward_backward()          # eager warm-up: sizes the library's training buffers (cudaMalloc is not capturable)
        self.grads.zero_()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            keep = forward_backward()
            if world > 1:
                dist.all_reduce(self.grads, group=group)
            eng.adam_step_dev(self.params, self.grads, self.exp_avg, self.exp_avg_sq, sb["adam"])
            eng.repack(self._pc, self._pf)
        self._graph = dict(graph=graph, sb=sb, n=n, keep=keep)
        return self

    def step_graph(self, ray_origins, ray_directions, target, expressions, latent_index, background=None):
        """One optimizer step by replaying the captured graph (capture() first): copies the step's inputs into the static buffers —
        on the current stream, so the caller may keep them on the device or in pinned host memory — and replays."""
        g = self._graph
        sb = g["sb"]
        if ray_origins.shape[0] != g["n"]:
            raise ValueError(f"the graph was captured for {g['n']} rays per step")
        self._own_engine()
        sb["ro"].copy_(ray_origins, non_blocking=True)
        sb["rd"].copy_(ray_directions, non_blocking=True)
        sb["tgt"].copy_(target, non_blocking=True)
        if sb["bg"] is not None:
            sb["bg"].copy_(background, non_blocking=True)
        sb["expr"].copy_(expressions.reshape(-1), non_blocking=True)
        sb["idx"].fill_(int(latent_index))
        g["graph"].replay()
        self.iter += 1
        self.eng.mark_synced(self.mc, self.mf)
        self.eng.packed_owner = self
        return self.loss[:2]

    def step(self, *args, **kwargs):
        """One optimizer step: gradients(...) then update().  Returns gradients()'s loss tensor."""
        loss = self.gradients(*args, **kwargs)
        self.update()
        return loss
