#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: rays/sec at 512x512 with 64 coarse + 128 fine samples per ray.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
    python bench.py --impl reference ...                       (the reference algorithm on the host cores)

One step = one full 512x512 frame per GPU (BASELINE config 2; with N GPUs config 5: N concurrent frames with
different expression codes, all-gathered into one [N,512,512,3] video tensor) => weak scaling, value = N*H*W*K / time.
Synthetic per-frame pose / expression / latent / background and random-init weights.

`value`  : inputs already resident in HBM; per step nfb_set_frame + the fused render kernel (+ NCCL all-gather
           when N>1), timed with CUDA events per step (L2 flushed between steps, outside the events).
`e2e`    : the same frame through the C-ABI host entry nfb_render_frame_host: pinned host expression / latent /
           background in, 11 floats per ray out, copies inside the timed region.
`roofline`: dominant kernel = the render kernel (fast mode: nfb::v6::render2_kernel, two tiles in flight); achieved = algorithmic FLOP per launch (1,100,032 FLOP per MLP
           evaluation x (2*Nc+Nf) evaluations per ray x rays) / its CUDA-event time; peak from MEASURED_PEAKS.json.
`cpu_baseline`: the oracle (a port of the reference, oracle/nerface_oracle.py) on the host cores for a 64x64 crop.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

NEAR, FAR = 0.2, 0.8
ALGO_FLOP_PER_EVAL = 1100032


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("NFB_PRECISION", "fast"), choices=["fast", "exact"])
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--num-coarse", type=int, default=64)
    ap.add_argument("--num-fine", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = max((float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_frame_crop(frame_index, H, W, crop, nc, nf, threads):
    """The reference algorithm (oracle port) on a crop x crop pixel block of the synthetic frame, CPU."""
    import nerface_oracle as O
    torch.set_num_threads(threads)
    fr = O.synthetic_frame(frame_index, H, W)
    pc, pf = O.random_init_params(100), O.random_init_params(101)
    ro, rd = O.ray_bundle(H, W, fr["intrinsics"], fr["pose"])
    r0, c0 = (H - crop) // 2, (W - crop) // 2
    ro, rd = ro[r0:r0 + crop, c0:c0 + crop].contiguous(), rd[r0:r0 + crop, c0:c0 + crop].contiguous()
    bg = fr["bg"][r0:r0 + crop, c0:c0 + crop].reshape(-1, 3)
    s = O.Sampling(nc, nf, False, 0.0, False, 65536)

    def run():
        with torch.no_grad():
            return O.run_one_iter(ro, rd, pc, pf if nf > 0 else None, s, NEAR, FAR, fr["expr"], fr["latent"], bg, "validation")
    return run, crop * crop


def pick_threads(H, W, nc, nf):
    """torch's intra-op pool does not scale to a 100+-core host on these small GEMMs (oversubscription makes it
    10-30x slower), so time a 16x16 crop at a few pool sizes and keep the fastest; that count is reported."""
    cores = os.cpu_count() or 1
    best = (None, 1e30)
    for th in sorted({cores, 64, 32, 16, 8}):
        if th > cores:
            continue
        run, _ = oracle_frame_crop(0, H, W, 16, nc, nf, th)
        run()
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (th, dt)
    return best[0]


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = pick_threads(a.height, a.width, a.num_coarse, a.num_fine)
    run, rays = oracle_frame_crop(0, a.height, a.width, 64, a.num_coarse, a.num_fine, cores)
    for _ in range(max(1, min(a.warmup, 1))):
        run()
    times = []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = rays * a.steps / total
    line = {"impl": "reference", "metric": "rays/sec at 512x512 (64c+128f samples)", "value": value, "unit": "rays/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"person_1-shaped eval {a.height}x{a.width}, {a.num_coarse}c+{a.num_fine}f, 76-dim expr + 32-dim latent",
                       "sample": "64x64 centre crop of the frame per step"},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "host_cores": os.cpu_count(), "kind": "port",
                             "sample": "oracle/nerface_oracle.py (bit-exact port of the reference, torch CPU FP32) on a 64x64 crop per step"},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    if a.impl == "reference":
        return reference_arm(a)

    sys.path.insert(0, os.path.join(ROOT, "4d-facial-avatars_b200"))
    import nerface_oracle as O
    import nerf
    from nerf import _engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    H, W, nc, nf = a.height, a.width, a.num_coarse, a.num_fine
    n = H * W

    # random-init weights (CPU generator, identical on every rank), moved to the device
    mk = lambda: nerf.models.ConditionalBlendshapePaperNeRFModel(  # noqa: E731
        num_encoding_fn_xyz=10, num_encoding_fn_dir=4, include_input_xyz=True, include_input_dir=False)
    mc, mf = mk(), mk()
    mc.load_state_dict(O.random_init_params(100))
    mf.load_state_dict(O.random_init_params(101))
    mc, mf = mc.to(dev), mf.to(dev)
    eng = _engine.renderer_for(dev)
    eng.sync_weights(mc, mf)

    # synthetic frames: frame f on rank r uses generator seed 42 + (f*world + r)
    n_frames = a.steps + a.warmup
    frames = [O.synthetic_frame(f * world + rank, H, W) for f in range(min(n_frames, 4))]
    dev_frames = [dict(expr=fr["expr"].to(dev), latent=fr["latent"].to(dev), bg=fr["bg"].reshape(-1, 3).to(dev).contiguous())
                  for fr in frames]
    host_frames = [dict(expr=fr["expr"].pin_memory(), latent=fr["latent"].pin_memory(),
                        bg=fr["bg"].reshape(-1, 3).contiguous().pin_memory()) for fr in frames]
    out_buf = torch.empty((11, n), device=dev)
    out_host = torch.empty((11 * n,), dtype=torch.float32).pin_memory()
    video = torch.empty((world, n, 3), device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)  # > 126 MB L2

    def step_resident(i, ev=None):
        fr, d = frames[i % len(frames)], dev_frames[i % len(frames)]
        eng.set_frame(d["expr"], d["latent"])
        if ev:
            ev[0].record()
        v = eng.render_camera(fr["pose"], fr["intrinsics"], H, W, 0, H, NEAR, FAR, nc, nf, background=d["bg"], out=out_buf,
                              precision=a.precision)
        if ev:
            ev[1].record()
        if world > 1:
            dist.all_gather_into_tensor(video.view(-1), v["rgb_fine"].reshape(-1))
        return v

    def step_host(i):
        fr, hst = frames[i % len(frames)], host_frames[i % len(frames)]
        eng.render_frame_host(fr["pose"], fr["intrinsics"], H, W, 0, H, NEAR, FAR, hst["expr"], hst["latent"], hst["bg"],
                              nc, nf, out_host, precision=a.precision)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, kernel_events=False):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        sync_all()
        wall0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(float(i))  # L2 flush, outside the timed events
            evs[i][0].record()
            if kernel_events:
                fn(a.warmup + i, kevs[i])
            else:
                fn(a.warmup + i)
            evs[i][1].record()
        sync_all()
        wall = time.perf_counter() - wall0
        ms = sum(s.elapsed_time(e) for s, e in evs)
        kms = sum(s.elapsed_time(e) for s, e in kevs) if kernel_events else None
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), kms, wall

    for i in range(max(a.warmup, 3)):
        step_resident(i)
        step_host(i)
    sync_all()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = eng.launch_count()
    ms_total, kernel_ms, wall = timed(step_resident, a.steps, kernel_events=True)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms_e2e, _, _ = timed(lambda i: step_host(i), a.steps)

    if rank == 0:
        pk, pk_src = peaks()
        rays_total = n * world * a.steps
        value = rays_total / (ms_total * 1e-3)
        e2e_value = rays_total / (ms_e2e * 1e-3)
        evals_per_ray = 2 * nc + nf
        flop_per_launch = n * evals_per_ray * ALGO_FLOP_PER_EVAL
        k_ms = kernel_ms / a.steps
        achieved = flop_per_launch / (k_ms * 1e-3) / 1e12
        peak = pk["bf16_tflops"]
        traffic = None  # dram bytes per launch of the dominant kernel, from the committed `ncu --set full` capture
        two_tile = a.precision == "fast" and os.environ.get("NFB_KERNEL") != "v4"
        kernel_name = "nfb::v6::render2_kernel" if two_tile else "nfb::render_kernel"
        tpath = os.path.join(ROOT, "profiles", "r1b_render_kernel_ncu.json" if two_tile else "r1_render_kernel_ncu.json")
        if os.path.exists(tpath) and (H, W, nc, nf, a.precision) == (512, 512, 64, 128, "fast"):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        # in-run parity: two image rows of the last rendered frame against the oracle
        last = (a.warmup + a.steps - 1) % len(frames)
        fr, dfr = frames[last], dev_frames[last]
        eng.set_frame(dfr["expr"], dfr["latent"])  # no collective here: only rank 0 runs the parity check
        v = eng.render_camera(fr["pose"], fr["intrinsics"], H, W, 0, H, NEAR, FAR, nc, nf, background=dfr["bg"], out=out_buf,
                              precision=a.precision)
        torch.cuda.synchronize()
        ro, rd = O.ray_bundle(H, W, fr["intrinsics"], fr["pose"])
        rows = slice(H // 2, H // 2 + 2)
        s = O.Sampling(nc, nf, False, 0.0, False, 65536)
        with torch.no_grad():
            ref = O.run_one_iter(ro[rows], rd[rows], O.random_init_params(100), O.random_init_params(101), s, NEAR, FAR,
                                 fr["expr"], fr["latent"], fr["bg"][rows].reshape(-1, 3), "validation")
        sl = slice((H // 2) * W, (H // 2 + 2) * W)
        names = ["rgb_coarse", "disp_coarse", "acc_coarse", "rgb_fine", "disp_fine", "acc_fine", "w_last"]
        parity = max(float((v[k][sl].cpu().reshape(r.shape) - r).abs().max()) for k, r in zip(names, ref))
        mse = float(((v["rgb_fine"][sl].cpu().reshape(ref[3].shape) - ref[3]) ** 2).mean())
        psnr = 99.0 if mse == 0 else min(99.0, -10.0 * __import__("math").log10(mse))

        cpu = None
        if not a.no_cpu_baseline:
            cores = pick_threads(H, W, nc, nf)
            run, crays = oracle_frame_crop(0, H, W, 64, nc, nf, cores)
            run()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                run()
                ts.append(time.perf_counter() - t0)
            cpu = {"value": crays / (sorted(ts)[1]), "unit": "rays/s", "cores": cores, "kind": "port",
                   "host_cores": os.cpu_count(),
                   "sample": "oracle port of the reference (torch CPU FP32, best of 8/16/32/64/all intra-op threads), 64x64 centre crop, median of 3"}
        line = {
            "metric": "rays/sec at 512x512 (64c+128f samples)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (tcgen05)" if a.precision == "fast" else "f16 hi+lo split x3 / f32 accumulate (tcgen05)",
            "data": "synthetic",
            "config": {"workload": f"person_1-shaped eval: {H}x{W}, {nc} coarse + {nf} fine samples/ray, 76-dim expr + 32-dim latent, "
                                   f"one frame per GPU per step", "precision": a.precision, "parallelism": f"frame-per-gpu x{world}",
                       "l2": "flushed between timed steps (256 MB write outside the events)",
                       "parity_max_abs_vs_oracle": parity, "psnr_rgb_fine_db": psnr, "parity_rays": 2 * W,
                       "wall_s_timed_region": wall},
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": (76 + 32 + 3 * n) * 4,
                    "d2h_bytes_per_step": 11 * n * 4, "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "frac_of_sustained": achieved / pk.get("bf16_tflops_sustained", peak), "peak_source": pk_src,
                         "kernel": kernel_name, "kernel_ms": k_ms, "flop_per_launch": flop_per_launch, "traffic": traffic,
                         "traffic_unit": "bytes of DRAM read+write per launch (ncu --set full, profiles/" + os.path.basename(tpath).replace(".json", ".md") + ")"},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
