"""Turns the ncu outputs a gpurun call brought back (gpurun_out/) into the small text summaries
kept under profiles/. Usage: python profiles/summarize.py <round-tag>
  gpurun_out/launches_<tag>.csv   from  ncu --metrics gpu__time_duration.sum --clock-control none --csv
  gpurun_out/prof_<tag>.ncu-rep   from  ncu --set full --clock-control none --import-source on
"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = ROOT / "profiles"

lst = ROOT / "gpurun_out" / f"launches_{tag}.csv"
if lst.exists():
    rows = list(csv.reader(open(lst)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        if not name.startswith("k_"):
            continue  # torch kernels of the synthetic-data generator are not the product
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v  # -> us
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out / f"{tag}_launches.md", "w") as f:
        f.write(f"# ncu launch list, round tag {tag}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none --csv python profiles/profile_target.py 10`\n"
                "(one bitplane of the bench workload: 8192x8192, 8x8 patches, 32 atoms; cold-cache, serialised launches --\n"
                "compare SHARES, not absolutes; the synthetic-data generator's torch kernels are filtered out)\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |\n")
        f.write(f"\ntotal {tot:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
    print("wrote", out / f"{tag}_launches.md")

rep = ROOT / "gpurun_out" / f"prof_{tag}.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    seen = collections.Counter()
    with open(out / f"{tag}_top_kernels.csv", "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow([w for w, _ in idx])
        wr.writerow([units[i] for _, i in idx])
        for r in data:
            name = r[hdr.index("Kernel Name")].split("(")[0]
            seen[name] += 1
            if seen[name] > 4:
                continue  # a few launches per kernel are enough
            wr.writerow([(r[i].split("(")[0] if w == "Kernel Name" else r[i]) for w, i in idx])
    print("wrote", out / f"{tag}_top_kernels.csv")
    # DRAM traffic per launch of the kernel with the largest summed duration (bench.py's roofline.traffic)
    import json
    ki = hdr.index("Kernel Name")
    di, ri, wi = hdr.index("gpu__time_duration.sum"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = collections.defaultdict(lambda: [0.0, 0.0, 0])
    for r in data:
        name = r[ki].split("(")[0].replace("void ", "").split("<")[0]
        if not name.startswith("k_"):
            continue
        tot[name][0] += float(r[di])
        tot[name][1] += float(r[ri]) * scale.get(units[ri], 1.0) + float(r[wi]) * scale.get(units[wi], 1.0)
        tot[name][2] += 1
    if tot:
        dom = max(tot.items(), key=lambda kv: kv[1][0])
        (out / "dominant_kernel_traffic.json").write_text(json.dumps({
            "kernel": dom[0], "workload": "8192x8192/8/32", "dram_bytes_per_launch": dom[1][1] / dom[1][2],
            "launches_captured": dom[1][2],
            "source": f"profiles/{tag}_top_kernels.csv (ncu --set full of profiles/profile_target.py; dram__bytes_read.sum + "
                      "dram__bytes_write.sum, mean over the captured launches of the kernel with the largest summed duration)"}, indent=1))
        print("dominant", dom[0], dom[1])
