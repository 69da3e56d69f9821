#!/usr/bin/env python
"""Turn gpurun_out/ ncu captures into the tracked summaries under profiles/.

usage: summarize_profiles.py <round tag, e.g. r1>

Reads   gpurun_out/launches_<tag>.csv            (ncu --metrics gpu__time_duration.sum, every launch of one bench step)
        gpurun_out/prof_<tag>_conv_tc.ncu-rep    (ncu --set full, the 14 conv blocks of one U-Net batch)
        gpurun_out/prof_<tag>_em.ncu-rep         (ncu --set full, one PR-GLS EM launch)
        gpurun_out/bench_<tag>.json              (the un-profiled bench line, for the share comparison)
Writes  profiles/<tag>_launches.md, profiles/<tag>_conv_tc.csv, profiles/<tag>_conv_tc.md, profiles/<tag>_em.md
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
TILES = int(sys.argv[2]) if len(sys.argv) > 2 else 38          # tiles of the captured U-Net batch (bench --tiles-per-batch)
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

CONV_NAMES = ["d0a 1>8", "d0b 8>16", "d1a 16>16", "d1b 16>32", "d2a 32>32", "d2b 32>64", "u2a 64>64", "u2b 64>64",
              "u1a 128>32", "u1b 32>32", "u0a 64>16", "u0b 16>16", "o_m2 32>8", "o_m1 8>8"]
GFLOP = [0.088, 1.416, 0.708, 1.416, 0.708, 1.416, 0.708, 0.708, 2.831, 0.708, 2.831, 0.708, 2.831, 0.708]  # GMAC/tile


def family(name):
    for key, fam in (("conv3_tc", "conv (tcgen05)"), ("first_conv", "conv (CUDA core, fused gather)"),
                     ("conv3_direct", "conv (CUDA core)"), ("prgls", "PR-GLS EM"),
                     ("greedy", "PR-GLS EM"), ("predict_one_rep", "PR-GLS EM"), ("trim_mean", "PR-GLS EM"),
                     ("pool_kernel", "unet aux"), ("upsample", "unet aux"), ("gather_tiles", "unet aux"),
                     ("head_scatter", "unet aux"), ("sgemm_bn", "FFN"), ("knn_features", "FFN"), ("ffn_pair", "FFN"),
                     ("box_", "LCN"), ("select_", "LCN")):
        if key in name:
            return fam
    return "other (torch fill / copies)"


def launches():
    path = os.path.join(G, f"launches_{tag}.csv")
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    per_k, per_f = collections.defaultdict(lambda: [0, 0.0]), collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)       # -> us
        name = r[ki].split("(")[0].replace("void ", "")
        per_k[name][0] += 1; per_k[name][1] += v
        f = family(r[ki]); per_f[f][0] += 1; per_f[f][1] += v
    tot = sum(v[1] for v in per_k.values())
    out = [f"# {tag}: every launch of one bench.py run (--steps 1 --warmup 3: device, serial-comparison and e2e arms, 8 frames in all), ncu gpu__time_duration.sum", "",
           "Source: `ncu --metrics gpu__time_duration.sum --clock-control none` around `python bench.py --steps 1 --warmup 3`.",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES with the bench line, not absolutes.", "",
           "| kernel family | launches | total ms | share |", "|---|---:|---:|---:|"]
    for f, (n, us) in sorted(per_f.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {f} | {n} | {us / 1e3:.3f} | {us / tot:.3f} |")
    bj = os.path.join(G, f"bench_{tag}.json")
    if os.path.isfile(bj):
        line = [l for l in open(bj) if l.startswith("{")][-1]
        b = json.loads(line)
        st = b["stage_ms_per_step"]
        out += ["", f"Un-profiled bench line of the same build: {b['ms_per_step']:.2f} ms/step; CUDA-event stage times per step: "
                + ", ".join(f"{k} {v:.2f} ms ({v / b['ms_per_step']:.3f})" for k, v in st.items()) + ".  "
                "(em / ffn run on the side streams of the frame pipeline, concurrently with conv / lcn of the next volumes, "
                "and their event times include waiting for a free SM, so the stage shares do not add up to 1; under ncu "
                "everything is serialised, which is why the EM's share of the launch list is larger than its share of "
                "the step.)"]
    out += ["", "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, (n, us) in sorted(per_k.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:70]}` | {n} | {us / 1e3:.3f} | {us / tot:.3f} |")
    open(os.path.join(P, f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")


METRICS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
           ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "tc pipe busy %"),
           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor math %"),
           ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tc smem wavefronts %"),
           ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lsu smem wavefronts %"),
           ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
           ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
           ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
           ("smsp__inst_executed.sum", "warp instr")]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def conv():
    rep = os.path.join(G, f"prof_{tag}_conv_tc.ncu-rep")
    if not os.path.isfile(rep):
        return
    h, units, rows = raw(rep)
    idx = [(h.index(m), lab, units[h.index(m)]) for m, lab in METRICS if m in h]
    ki = h.index("Kernel Name")
    with open(os.path.join(P, f"{tag}_conv_tc.csv"), "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["block", "kernel"] + [f"{lab} [{u}]" for _, lab, u in idx])
        for n, r in zip(CONV_NAMES, rows):
            wr.writerow([n, r[ki].split("(")[0][-28:]] + [r[i] for i, _, _ in idx])
    t = h.index("gpu__time_duration.sum")
    dr, dw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")

    def to_ms(v, u):
        v = float(v.replace(",", ""))
        return {"ms": v, "us": v / 1e3, "ns": v / 1e6, "s": v * 1e3}.get(u, v)

    def to_mb(v, u):
        v = float(v.replace(",", ""))
        return {"Mbyte": v, "Gbyte": v * 1e3, "Kbyte": v / 1e3, "byte": v / 1e6}.get(u, v)

    md = [f"# {tag}: convolution blocks of one U-Net batch ({TILES} tiles of 160x160x16), ncu --set full", "",
          "One row per conv block in graph order (`first_conv_kernel` = CUDA-core Cin=1 block fused with the tile gather, "
          "`conv3_tcx_kernel<Cout, BX, STAGES>` = x-stacked tcgen05 kernel, `conv3_tc_kernel<...>` = 27-tap tcgen05 kernel; "
          "see the csv for which kernel ran a block); full metric table in "
          f"`{tag}_conv_tc.csv`.  `tc pipe busy` = sm__pipe_tc_cycles_active (tensor-core pipe incl. operand fetch), "
          "`tensor math` = sm__pipe_tensor_cycles_active, `tc smem` = l1tex__data_pipe_tc_wavefronts_mem_shared "
          "(shared-memory wavefronts read by the tensor core, % of peak).", "",
          "| block | ms | TFLOP/s (algorithmic) | tc pipe busy % | tensor math % | tc smem % | DRAM MB (read+write) | algorithmic MB |",
          "|---|---:|---:|---:|---:|---:|---:|---:|"]
    cin = [1, 8, 16, 16, 32, 32, 64, 64, 128, 32, 64, 16, 32, 8]
    cout = [8, 16, 16, 32, 32, 64, 64, 64, 32, 32, 16, 16, 8, 8]
    vox = [409600, 409600, 102400, 102400, 25600, 25600, 6400, 6400, 25600, 25600, 102400, 102400, 409600, 409600]
    g = lambda m: h.index(m)
    tot_ms = 0.0
    for n, r, gm, ci, co, vx in zip(CONV_NAMES, rows, GFLOP, cin, cout, vox):
        ms = to_ms(r[t], units[t]); tot_ms += ms
        mb = to_mb(r[dr], units[dr]) + to_mb(r[dw], units[dw])
        alg = TILES * vx * 4 * ((1 if ci == 1 else ci) + co) / 1e6
        md.append(f"| {n} | {ms:.3f} | {TILES * gm * 2 / ms:.1f} | "
                  f"{float(r[g('sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed')]):.1f} | "
                  f"{float(r[g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')]):.1f} | "
                  f"{float(r[g('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed')]):.1f} | "
                  f"{mb:.0f} | {alg:.0f} |")
    md += ["", f"Sum of the 14 blocks: {tot_ms:.3f} ms per {TILES} tiles -> {TILES * 35.573 / tot_ms:.1f} TFLOP/s algorithmic "
           "(35.573 GFLOP/tile; per-launch times under ncu are cold-cache and serialised).  Reading: the tcgen05 blocks are "
           "bound by the tensor core's shared-memory operand path (`tc pipe busy` >> `tensor math`: an M=128, K=16 fp16 MMA "
           "reads a 4 KB A tile plus its B rows from shared memory at 128 B/clk, which takes longer than its math for "
           "N <= 96), not by MMA math and not by HBM; DESIGN.md 3.2 has the arithmetic."]
    open(os.path.join(P, f"{tag}_conv_tc.md"), "w").write("\n".join(md) + "\n")
    total_bytes = sum(to_mb(r[dr], units[dr]) + to_mb(r[dw], units[dw]) for r in rows[:14]) * 1e6
    json.dump({"kernel": f"14 conv blocks of one {TILES}-tile U-Net batch (first_conv + conv3_tcx + conv3_tc)", "launches": 14,
               "dram_bytes_per_launch_avg": total_bytes / 14, "tiles_per_batch": TILES,
               "source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, profiles/{tag}_conv_tc.csv"},
              open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)


def em():
    rep = os.path.join(G, f"prof_{tag}_em.ncu-rep")
    if not os.path.isfile(rep):
        return
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    keep = [l.rstrip() for l in det.splitlines() if any(k in l for k in (
        "Duration", "Registers Per Thread", "Dynamic Shared Memory", "Executed Ipc Active", "Issue Slots Busy",
        "Memory Throughput", "DRAM Throughput", "Theoretical Occupancy", "Block Size", "Grid Size", "No Eligible",
        "Warp Cycles Per Issued", "L1/TEX Hit", "Compute (SM) Throughput"))]
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep,
                            os.path.join(ROOT, "3deecelltracker_b200", "csrc", "prgls.cu"), "prgls_kernel", "22"],
                           capture_output=True, text=True).stdout
    md = [f"# {tag}: PR-GLS EM kernel (one launch: N=164 refs, M=164 targets, 19 iterations), ncu --set full", "",
          "One persistent 1024-thread CTA on ONE SM: the chip-level throughput percentages are tiny by construction; "
          "what matters is the time and where the warps wait.", "", "```"] + keep + ["```", "",
          "Stall samples attributed to source lines (scripts/ncu_lines.py: SASS offsets joined with nvdisasm line info):", "",
          "```", lines.rstrip(), "```"]
    open(os.path.join(P, f"{tag}_em.md"), "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    launches()
    conv()
    em()
    print("profiles written:", sorted(f for f in os.listdir(P) if f.startswith(tag)))
