"""Summarise gpurun_out/r02/*.ncu-rep into profiles/r02_<model>_<kernel>.{md,json} and one overview table profiles/r02_kernels.md."""
import csv, glob, io, json, os, subprocess, sys

src_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02"
tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
rows_out = []
for rep in sorted(glob.glob(os.path.join(src_dir, "*.ncu-rep"))):
    base = os.path.basename(rep)[:-8]
    out = os.path.join("profiles", f"{tag}_{base}")
    subprocess.run([sys.executable, "tools/summarize_ncu.py", rep, out], check=True)
    js = json.load(open(out + ".json"))
    g = lambda k: js.get(k, {}).get("value", "")
    def num(k):
        try:
            return float(str(g(k)).replace(",", ""))
        except ValueError:
            return float("nan")
    dur_us = num("gpu__time_duration.sum") / 1e3 if js.get("gpu__time_duration.sum", {}).get("unit") == "ns" else num("gpu__time_duration.sum")
    dram = (num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
    ur, uw = js.get("dram__bytes_read.sum", {}).get("unit", ""), js.get("dram__bytes_write.sum", {}).get("unit", "")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = num("dram__bytes_read.sum") * scale.get(ur, 1) + num("dram__bytes_write.sum") * scale.get(uw, 1)
    rows_out.append((base, dur_us, dram / 1e6, dram / 1e3 / dur_us if dur_us else 0, g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                     g("launch__registers_per_thread"), g("sm__warps_active.avg.pct_of_peak_sustained_active")))
lines = [f"# {tag}: one `ncu --set full --clock-control none` capture per kernel (warm launch inside a training step, tools/profile_round.sh)", "",
         "| capture | duration µs | DRAM MB (read+write) | DRAM GB/s | tensor pipe active % | DRAM % of peak | L2 % of peak | regs/thread | warps active % |",
         "|---|---|---|---|---|---|---|---|---|"]
for r in rows_out:
    lines.append(f"| {r[0]} | {r[1]:.1f} | {r[2]:.1f} | {r[3]:.0f} | {r[4]} | {r[5]} | {r[6]} | {r[7]} | {r[8]} |")
open(os.path.join("profiles", f"{tag}_kernels.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
