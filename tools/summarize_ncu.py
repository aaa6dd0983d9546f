"""Turn an .ncu-rep of one kernel into the markdown + json summaries kept under profiles/.

    python tools/summarize_ncu.py gpurun_out/X.ncu-rep profiles/r01_name  [--roles]      (--roles: conv_tc2 warp-role breakdown)
"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
roles = "--roles" in sys.argv
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
keys = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second",
    "smsp__inst_executed_op_tma_ld.sum", "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__sass_l1tex_m_xbar2l1tex_read_bytes_mem_global_op_ldgsts_cache_bypass.sum",
    "smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum", "smsp__inst_executed_op_ldgsts.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
]
name = rows[2][rows[0].index("Kernel Name")] if "Kernel Name" in rows[0] else "?"
lines = [f"# ncu summary: `{name}`", "", f"source report: `{rep}` (ncu --set full --clock-control none --import-source on)", "",
         "| metric | value | unit |", "|---|---|---|"]
js = {}
for k in keys:
    if k in d:
        u, v = d[k]
        lines.append(f"| {k} | {v} | {u} |")
        js[k] = {"value": v, "unit": u}
if roles:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    open("/tmp/_src.csv", "w").write(src)
    r = subprocess.run([sys.executable, "tools/ncu_roles2.py", "/tmp/_src.csv"], capture_output=True, text=True).stdout
    lines += ["", "## warp-state samples by warp role (source page)", "", "```", r.rstrip(), "```"]
open(out + ".md", "w").write("\n".join(lines) + "\n")
json.dump(js, open(out + ".json", "w"), indent=1)
print("\n".join(lines[:34]))
