"""Turn the scratch artefacts of tools/profile_r2.sh (gpurun_out/) into the tracked summaries under profiles/ (round 2):
launch lists, one markdown + json per `ncu --set full` capture, SASS instruction counts of the shipped library, bench lines."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
LIB = os.path.join(ROOT, "4d-facial-avatars_b200", "lib", "libnfb.so")

METRICS = [
    ("gpu__time_duration.sum", "kernel duration under ncu"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_size", "cluster size"),
    ("launch__registers_per_thread", "registers / thread at launch (setmaxnreg re-partitions them inside render3_kernel)"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / CTA"),
    ("dram__bytes_read.sum", "DRAM read per launch"), ("dram__bytes_write.sum", "DRAM write per launch"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors (32 B) requested by the SMs (the weight stream)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor-memory pipe active"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe (HMMA sub-pipe) active"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tcgen05.ld/st issue slots"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (max of sub-metrics)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"), ("sm__cycles_elapsed.avg.per_second", "SM clock during the capture"),
]


def read_rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return None
    return {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}


def fnum(v):
    try:
        return float(str(v).replace(",", ""))
    except Exception:
        return None


def to_bytes(unit, v):
    x = fnum(v)
    if x is None:
        return None
    u = (unit or "").lower()
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def summarize(tag, rep, title, command, algo_flop=None, notes=()):
    d = read_rep(os.path.join(G, rep))
    if d is None:
        print("missing", rep)
        return
    lines = [f"# ncu --set full — {title}", "", f"Command: `{command}` (report: gpurun_out/{rep}, scratch, not tracked).", "",
             "| metric | value | note |", "|---|---|---|"]
    js = {"kernel": d.get("Kernel Name", ("", ""))[1], "report": rep}
    for key, note in METRICS:
        hit = [h for h in d if h == key]
        if not hit:
            continue
        u, v = d[hit[0]]
        lines.append(f"| `{key}` | {v} {u} | {note} |")
        js[key] = {"value": fnum(v), "unit": u}
    rd, wr = to_bytes(*d.get("dram__bytes_read.sum", (None, None))), to_bytes(*d.get("dram__bytes_write.sum", (None, None)))
    dur_u, dur_v = d.get("gpu__time_duration.sum", (None, None))
    dur_ms = fnum(dur_v) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get((dur_u or "").replace("second", "s").replace("usecond", "us").replace("msecond", "ms").replace("nsecond", "ns"), 1e-6) if fnum(dur_v) else None
    if dur_u in ("msecond", "ms"):
        dur_ms = fnum(dur_v)
    elif dur_u in ("usecond", "us"):
        dur_ms = fnum(dur_v) * 1e-3
    elif dur_u in ("nsecond", "ns"):
        dur_ms = fnum(dur_v) * 1e-6
    elif dur_u in ("second", "s"):
        dur_ms = fnum(dur_v) * 1e3
    js["dram_bytes_per_launch"] = (rd or 0) + (wr or 0) if rd is not None else None
    js["duration_ms"] = dur_ms
    js["block_size"] = int(fnum(d.get("launch__block_size", (None, "0"))[1]) or 0)
    lines += ["", f"DRAM traffic per launch = {js['dram_bytes_per_launch'] / 1e6:.2f} MB." if js["dram_bytes_per_launch"] is not None else ""]
    if algo_flop and dur_ms:
        tf = algo_flop / (dur_ms * 1e-3) / 1e12
        lines.append(f"Arithmetic: {algo_flop / 1e12:.2f} TFLOP algorithmic per launch / {dur_ms:.2f} ms (under ncu, clocks not locked) = {tf:.1f} TFLOP/s = "
                     f"{100 * tf / 1652.1:.1f}% of the measured 1652.1 TFLOP/s bf16 peak.")
        js["tflops_under_ncu"] = tf
    lines += list(notes)
    open(os.path.join(P, f"{tag}.md"), "w").write("\n".join(lines) + "\n")
    json.dump(js, open(os.path.join(P, f"{tag}.json"), "w"), indent=1)
    print("wrote", tag)


def launch_list(tag, csv_name, title, command, note):
    path = os.path.join(G, csv_name)
    if not os.path.exists(path):
        return
    allrows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = allrows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    to_ns = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}
    agg = collections.OrderedDict()
    for r in allrows[1:]:
        name = r[ki].split("(")[0]
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", "")) * to_ns.get(r[ui], 1.0)
    tot = sum(v[1] for v in agg.values()) or 1.0
    lines = [f"# ncu launch list — {title}", "", f"`{command}`", "", note, "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in agg.items():
        lines.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.2f}% |")
    open(os.path.join(P, f"{tag}.md"), "w").write("\n".join(lines) + "\n")
    print("wrote", tag)


def sass_summary():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    pats = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "USETMAXREG", "UTMALDG", "HMMA.", "SYNCS", "MUFU.SIN", "FADD2", "STL", "LDL"]
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[cur] = collections.Counter()
            continue
        if cur:
            for p_ in pats:
                if p_ in line:
                    per[cur][p_] += 1
    lines = ["# SASS evidence — shipped lib/libnfb.so (`cuobjdump -sass`, counts of instruction mnemonics per kernel)", "",
             "`UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st, `UBLKCP` = cp.async.bulk (TMA engine, linear), `UTCBAR` = tcgen05.commit,",
             "`USETMAXREG` = setmaxnreg, `SYNCS` = mbarrier ops, `STL`/`LDL` = local-memory spills.  No `HMMA` (legacy mma.sync) anywhere.", "",
             "| kernel | " + " | ".join(pats) + " |", "|---|" + "---|" * len(pats)]
    for k, c in per.items():
        if any(c[p_] for p_ in ("UTCHMMA", "LDTM", "UBLKCP")) or "kernel" in k:
            lines.append(f"| `{k}` | " + " | ".join(str(c[p_]) for p_ in pats) + " |")
    open(os.path.join(P, "r2_sass_summary.md"), "w").write("\n".join(lines) + "\n")
    print("wrote r2_sass_summary")


def r2c():
    """The training kernels as shipped at the end of round 2 (tools/profile_r2c.sh)."""
    T = "python tools/train_bench.py --steps 2 --warmup 2 --impl fused"
    summarize("r2c_dw_kernel_ncu", "prof_r2c_dw_kernel.ncu-rep",
              "`nfb::dw::dw_kernel` (weight-gradient GEMMs, BOTH networks in one launch, 8 job groups), 2048 rays 64c+64f",
              f"ncu --set full --clock-control none --import-source on -k regex:dw_kernel -s 2 -c 1 {T}")
    summarize("r2c_chain_kernel_ncu", "prof_r2c_chain_kernel.ncu-rep", "`nfb::chain::chain_kernel` (dX chain, record-saver warps), 2048 rays 64c+64f",
              f"ncu --set full ... -k regex:chain_kernel -s 2 -c 1 {T}")
    summarize("r2c_fwd_save_ncu", "prof_r2c_render_kernel.ncu-rep", "`nfb::render_kernel<fast, SAVE>` (training forward, record-saver warps), 2048 rays 64c+64f",
              f"ncu --set full ... -k regex:render_kernel -s 2 -c 1 {T}")
    launch_list("r2c_train_launches", "train_launches_r2c.csv", "training iterations as shipped (`tools/train_bench.py --impl fused`, launches 70-170 of the run)",
                "ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 100 python tools/train_bench.py --steps 4 --warmup 3 --impl fused",
                "One iteration = frame_fold, 4 torch RNG kernels (the reference's noise draws), render_kernel<SAVE>, loss_grad, composite_bwd + scale, "
                "chain, ONE dw launch (both networks), finalize + fin_dir0, adam, fold_feat + repack (+ memsets; the index kernels are the bench's batch gathers).")
    sass_summary()
    for src in ("r2c_bench_n1.json", "r2c_train_bench_n1.json"):
        if os.path.exists(os.path.join(G, src)) and os.path.getsize(os.path.join(G, src)) > 0:
            shutil.copy(os.path.join(G, src), os.path.join(P, src))
    log = os.path.join(G, "r2c_gputests.log")
    if os.path.exists(log):
        tail = open(log).read().strip().splitlines()[-3:]
        open(os.path.join(P, "r2c_gputests.txt"), "w").write("python -m pytest tests -m gpu -q -x --timeout 100   (1 x B200, tools/profile_r2c.sh)\n" + "\n".join(tail) + "\n")


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    import sys
    if "--r2c" in sys.argv:
        r2c()
        raise SystemExit(0)
    B = "python bench.py --no-extras --no-cpu-baseline --steps 1 --warmup 1"
    flop512 = 262144 * 256 * 1100032
    summarize("r2_render3_kernel_ncu", "prof_r2_render3.ncu-rep", "`nfb::v7::render3_kernel` (fast mode, shipped default), 512x512, 64c+128f",
              f"ncu --set full --clock-control none --import-source on -k regex:render3_kernel -s 3 -c 1 {B}", flop512)
    summarize("r2_render2_kernel_ncu", "prof_r2_render2.ncu-rep", "`nfb::v6::render2_kernel` (NFB_KERNEL=v6: two tiles, passes not pipelined), 512x512, 64c+128f",
              f"NFB_KERNEL=v6 ncu --set full ... -k regex:render2_kernel -s 3 -c 1 {B}", flop512)
    summarize("r2_render_kernel_exact_ncu", "prof_r2_exact.ncu-rep", "`nfb::render_kernel<exact>` (FP16 hi+lo x3), 512x512, 64c+128f",
              f"ncu --set full ... -k regex:render_kernel -s 3 -c 1 {B} --precision exact", flop512)
    summarize("r2_render3_cfg4_ncu", "prof_r2_cfg4.ncu-rep", "`nfb::v7::render3_kernel`, BASELINE config 4: 1024x1024, 128c+256f (one ray per stream)",
              f"ncu --set full ... -k regex:render3_kernel -s 3 -c 1 {B} --height 1024 --width 1024 --num-coarse 128 --num-fine 256", 1048576 * 512 * 1100032)
    T = "python tools/train_bench.py --steps 2 --warmup 2 --impl fused"
    summarize("r2_dw_kernel_ncu", "prof_r2_dw_kernel.ncu-rep", "`nfb::dw::dw_kernel` (weight-gradient GEMMs, coarse network launch), 2048 rays 64c+64f", f"ncu --set full ... -k regex:dw_kernel -s 4 -c 1 {T}")
    summarize("r2_chain_kernel_ncu", "prof_r2_chain_kernel.ncu-rep", "`nfb::chain::chain_kernel` (dX chain), 2048 rays 64c+64f", f"ncu --set full ... -k regex:chain_kernel -s 4 -c 1 {T}")
    summarize("r2_fwd_save_ncu", "prof_r2_fwd_save.ncu-rep", "`nfb::render_kernel<fast, SAVE>` (training forward), 2048 rays 64c+64f", f"ncu --set full ... -k regex:render_kernel -s 3 -c 1 {T}")
    launch_list("r2_launches", "launches_r2.csv", "evaluation (`bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline`)",
                "ncu --metrics gpu__time_duration.sum --clock-control none -c 60 python bench.py ...",
                "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.  `repack_kernel` is the one-time weight packing "
                "(one launch per network), `frame_fold_kernel` is `nfb_set_frame` (one launch per frame, both networks), `render3_kernel` the frame.")
    launch_list("r2_train_launches", "train_launches_r2.csv", "training iterations (`tools/train_bench.py --impl fused`, launches 200-320 of the run)",
                "ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 python tools/train_bench.py --steps 4 --warmup 3 --impl fused",
                "One iteration = frame_fold, 4 torch RNG kernels (the reference's noise draws), render_kernel<SAVE>, loss_grad, composite_bwd + scale, "
                "chain, 2 x dw, 2 x (finalize + fin_dir0), latent_grad, adam, repack (+ memsets).")
    sass_summary()
    for src, dst in (("r2_bench_n1.json", "r2_bench_n1.json"), ("r2_train_bench_n1.json", "r2_train_bench_n1.json")):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst))
