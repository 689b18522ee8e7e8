"""Turn the scratch artefacts in gpurun_out/ (ncu launch list + full capture, bench lines, phase timers) into the tracked
summaries under profiles/.  Usage: python tools/summarize_profiles.py r1"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
KERNEL = sys.argv[2] if len(sys.argv) > 2 else "render_kernel"   # substring of the dominant kernel's name (r1b: render2_kernel)
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)

# ---- launch list
rows = [r for r in csv.reader(open(f"{G}/launches_{tag}.csv")) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += float(r[14])
tot = sum(v[1] for v in agg.values())
rk = [float(r[14]) for r in rows if KERNEL in r[4]]
lines = [f"# ncu launch list — {tag} (`ncu --metrics gpu__time_duration.sum --clock-control none -c 80 python bench.py --steps 2 --warmup 1 --no-cpu-baseline`)", "",
         "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.  The first launches are the one-time "
         "weight packing (`nfb_load_weights`, two networks); `frame_fold_kernel` is `nfb_set_frame` (two per frame).", "",
         "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, v in agg.items():
    lines.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.2f}% |")
lines += ["", f"`{KERNEL}` launches (512x512, 64c+128f): {', '.join(f'{x / 1e6:.1f}' for x in rk)} ms — "
              f"{100 * sum(rk) / tot:.1f}% of all GPU time in the run; within a timed step (2 x frame_fold + render) it is >99.9%."]
open(f"{P}/{tag}_launches.md", "w").write("\n".join(lines) + "\n")

# ---- full capture
raw = subprocess.run(["ncu", "-i", f"{G}/prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
d = {h: (u, v) for h, u, v in zip(rr[0], rr[1], rr[2])}


def g(k):
    for h in d:
        if h == k or h.endswith(k):
            return d[h]
    return (None, None)


sel = {
    "gpu__time_duration.sum": "kernel duration under ncu",
    "launch__grid_size": "grid (persistent: one CTA per SM, clusters of 2)", "launch__block_size": "block",
    "launch__cluster_size": "cluster size", "launch__registers_per_thread": "registers / thread",
    "launch__shared_mem_per_block_dynamic": "dynamic shared memory / CTA",
    "dram__bytes_read.sum": "DRAM read per launch", "dram__bytes_write.sum": "DRAM write per launch (outputs stay in the write-back L2 during the launch)",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "DRAM throughput % of peak",
    "lts__t_sectors_srcunit_tex_op_read.sum": "L2 read sectors (32 B) requested by SMs: the weight stream (multicast: one read per SM pair; two-tile kernel: one load per tile pair)",
    "lts__t_sector_hit_rate.pct": "L2 hit rate",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor-memory pipe active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active": "tcgen05.ld/st issue slots",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "SM throughput (max of sub-metrics)",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved occupancy (10 warps of 64 by design)",
    "smsp__inst_executed.sum": "warp instructions executed", "sm__cycles_elapsed.max": "SM cycles elapsed",
}
out = [f"# ncu --set full — `{KERNEL}` (fast mode), 512x512, 64c+128f, {tag}", "",
       f"Command: `ncu --set full --clock-control none --import-source on -k regex:{KERNEL} -s 3 -c 1 python bench.py --steps 1 --warmup 1 "
       f"--no-cpu-baseline` (report: gpurun_out/prof_{tag}.ncu-rep, scratch, not tracked).", "", "| metric | value | note |", "|---|---|---|"]
for k, note in sel.items():
    u, v = g(k)
    out.append(f"| `{k}` | {v} {u} | {note} |")
ru, rv = g("dram__bytes_read.sum")
wu, wv = g("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = float(rv) * scale.get(ru, 1.0) + float(wv) * scale.get(wu, 1.0)
dur = float(g("gpu__time_duration.sum")[1])
su, sv = g("lts__t_sectors_srcunit_tex_op_read.sum")
out += ["", f"DRAM traffic per launch = {traffic / 1e6:.2f} MB (algorithmic per launch: 3.1 MB background in + 11.5 MB outputs + 3.5 MB weights/biases; "
            "the outputs had not been written back from L2 when the kernel ended).",
        f"Arithmetic: 73.82 TFLOP algorithmic per launch / {dur:.2f} ms = {73.82 / dur * 1e3:.1f} TFLOP/s = {73.82 / dur * 1e3 / 1652.1 * 100:.1f}% of the measured 1652.1 TFLOP/s bf16 peak.",
        f"L2->SM weight stream: {float(sv):.3e} sectors x 32 B = {float(sv) * 32 / 1e9:.0f} GB per launch = {float(sv) * 32 / dur / 1e9:.2f} TB/s."]
bench = json.load(open(f"{G}/bench_{tag}.json"))
ref = json.load(open(f"{G}/bench_reference_{tag}.json"))
out += ["", "## bench.py lines of the same build", "",
        f"* fast: {bench['value']:.4g} rays/s resident, e2e {bench['e2e']['value']:.4g} rays/s, roofline frac {bench['roofline']['frac']:.3f}, "
        f"parity max|d| {bench['config']['parity_max_abs_vs_oracle']:.2e}, clocks {bench['clocks']}",
        f"* cpu_baseline (oracle port, {bench['cpu_baseline']['cores']} of {bench['cpu_baseline']['host_cores']} host threads): {bench['cpu_baseline']['value']:.1f} rays/s; "
        f"`--impl reference`: {ref['value']:.1f} rays/s"]
if os.path.exists(f"{G}/bench_exact_{tag}.json"):
    ex = json.load(open(f"{G}/bench_exact_{tag}.json"))
    out.append(f"* exact (FP16 hi+lo x3): {ex['value']:.4g} rays/s, parity max|d| {ex['config']['parity_max_abs_vs_oracle']:.2e}")
for name in (f"phase_fast_{tag}.txt", f"phase_exact_{tag}.txt"):
    if os.path.exists(f"{G}/{name}"):
        out += ["", f"## phase timers ({name}; cycles per 128-row tile, one observer thread per warp role)", "", "```"] + open(f"{G}/{name}").read().strip().splitlines() + ["```"]
open(f"{P}/{tag}_render_kernel_ncu.md", "w").write("\n".join(out) + "\n")
json.dump({"kernel": KERNEL, "config": "512x512 64c+128f fast", "dram_bytes_per_launch": traffic, "duration_ms_under_ncu": dur},
          open(f"{P}/{tag}_render_kernel_ncu.json", "w"), indent=1)
for src, dst in ((f"bench_{tag}.json", f"{tag}_bench_n1.json"), (f"bench_reference_{tag}.json", f"{tag}_bench_reference_n1.json"),
                 (f"bench_exact_{tag}.json", f"{tag}_bench_exact_n1.json")):
    if os.path.exists(f"{G}/{src}"):
        shutil.copy(f"{G}/{src}", f"{P}/{dst}")
print("\n".join(out[:32]))

# ---- training iteration (optional): launch list of tools/train_bench.py + its JSON line
if os.path.exists(f"{G}/train_launches_{tag}.csv"):
    rows = [r for r in csv.reader(open(f"{G}/train_launches_{tag}.csv")) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += float(r[14])
    tot = sum(v[1] for v in agg.values())
    tl = [f"# ncu launch list of the training iteration — {tag} (`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python "
          "tools/train_bench.py --steps 2 --warmup 1`: 3 iterations of 2048 rays, 64c+64f, fwd + bwd + Adam)", "",
          "Per-launch times under ncu are serialised and cold-cache: compare shares.", "",
          "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.002:
            continue
        tl.append(f"| `{k[:90]}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.1f} | {100 * v[1] / tot:.1f}% |")
    if os.path.exists(f"{G}/train_bench_{tag}.json"):
        tb = json.load(open(f"{G}/train_bench_{tag}.json"))
        tl += ["", "## tools/train_bench.py line of the same build", "", "```", json.dumps(tb), "```"]
        shutil.copy(f"{G}/train_bench_{tag}.json", f"{P}/{tag}_train_bench_n1.json")
    for kname in ("chain", "dw"):
        rep = f"{G}/prof_{kname}_{tag}.ncu-rep"
        if os.path.exists(rep):
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rr = list(csv.reader(raw.splitlines()))
            dd = {h: (u, v) for h, u, v in zip(rr[0], rr[1], rr[2])}
            tl += ["", f"## ncu --set full — `{kname}` kernel (one launch)", "", "| metric | value |", "|---|---|"]
            for key in ("gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
                        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
                        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"):
                for h in dd:
                    if h == key or h.endswith(key):
                        tl.append(f"| `{key}` | {dd[h][1]} {dd[h][0]} |")
                        break
    open(f"{P}/{tag}_train_launches.md", "w").write("\n".join(tl) + "\n")

