"""Summarise an .ncu-rep (brought back in gpurun_out/) into profiles/<name>.md: the metrics the
roofline statement rests on, per profiled launch.  Usage: python tools/ncu_summary.py REP NAME [note]"""
import csv
import io
import subprocess
import sys

rep, name = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__inst_executed.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
]
out = [f"# ncu summary: {name}", "", note, "", f"source report: `{rep}` (ncu --set full --clock-control none)", ""]
for r in rows[2:]:
    kn = r[hdr.index("Kernel Name")]
    out.append(f"## {kn}")
    out.append("")
    out.append("| metric | value | unit |")
    out.append("|---|---|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out.append(f"| {k} | {r[i]} | {units[i]} |")
    st = [(h, r[hdr.index(h)]) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    st = sorted(st, key=lambda x: -float(x[1].replace(",", "") or 0))[:6]
    out.append("")
    out.append("top warp stall reasons (warps stalled per issue-active cycle): " +
               ", ".join(f"{h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')} {float(v):.2f}" for h, v in st))
    out.append("")
open(f"profiles/{name}.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
