"""Assemble the tracked evidence of a round under profiles/ from a tools/gpu_evidence.sh run.
Usage: python tools/make_profiles.py <tag in gpurun_out> <round prefix, e.g. r02>"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

tag, rnd = sys.argv[1], sys.argv[2]
G, P = "gpurun_out", "profiles"


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


# ---- bench lines
for w in ("c1", "c2", "c3", "c4", "c5", "reference"):
    src = f"{G}/{tag}_bench_{w}.json"
    if os.path.exists(src):
        json.dump(last_json(src), open(f"{P}/{rnd}_bench_{w}.json", "w"), indent=1)

# ---- launch list of the default bench command
src = f"{G}/{tag}_launches.csv"
if os.path.exists(src):
    shutil.copy(src, f"{P}/{rnd}_launches_bench_c5.csv")
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    # one solve = the launches between two k_intr_init_state of the default (tensor-core) sweep pass
    names = [r[4].split("(")[0].replace("void ", "") for r in rows]
    t = [float(r[-1]) / 1e3 for r in rows]
    first = next(i for i, n in enumerate(names) if n == "k_intr_init_state")
    loop = ("k_intr_sweep_ws", "k_intr_decide", "k_intr_schur", "k_intr_solve")
    second = next(i for i in range(first + 1, len(names)) if not names[i].startswith(loop))
    agg = collections.OrderedDict()
    for n, x in zip(names[first:second], t[first:second]):
        agg.setdefault(n, [0, 0.0])
        agg[n][0] += 1
        agg[n][1] += x
    tot = sum(v[1] for v in agg.values())
    out = [f"# Per-kernel share of one full LM solve at C5 (100,000 views x 400 points), round {rnd}", "",
           f"From `{rnd}_launches_bench_c5.csv`: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py "
           "--steps 2 --warmup 3` (cold-cache and serialised launches: compare shares, not absolutes).  "
           "The rows are the launches of ONE solve (`k_intr_init_state` and the LM loop that follows it).", "",
           "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
    for n, v in agg.items():
        out.append(f"| `{n}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % | {v[1] / v[0]:.1f} |")
    out.append(f"| total | {sum(v[0] for v in agg.values())} | {tot:.1f} | | |")
    open(f"{P}/{rnd}_launch_shares.md", "w").write("\n".join(out) + "\n")
    print("\n".join(out))

# ---- ncu --set full capture of the sweep
rep = f"{G}/{tag}_sweep.ncu-rep"
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, r)}
    keys = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic sm__warps_active.avg.pct_of_peak_sustained_active
dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct
smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active smsp__warps_eligible.avg.per_cycle_active
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum""".split()
    out = [f"# ncu --set full: `k_intr_sweep_ws<0, 8>` at FULL C5 size (100,000 views x 400 points, one launch = 4e7 observations), round {rnd}", "",
           f"Capture: `ncu --set full --clock-control none --import-source on -k regex:k_intr_sweep -s 2 -c 1 python tools/profile_sweep.py 100000 400` "
           f"(`tools/gpu_evidence.sh`), after the same command had exited 0 without ncu.  The .ncu-rep stays in gpurun_out/ (scratch).", "",
           "| metric | value | unit |", "|---|---|---|"]
    for k in keys:
        if k in d:
            out.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
    st = sorted(((float(d[k][0].replace(",", "") or 0), k) for k in d
                 if "issue_stalled" in k and k.endswith("per_issue_active.ratio")), reverse=True)[:10]
    out += ["", "Warp stall reasons (warps stalled per issue-active cycle):", ""]
    out += [f"* {k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')}: {v:.2f}" for v, k in st]
    reg = subprocess.run([sys.executable, "tools/ncu_regions.py", rep, "1250000"], capture_output=True, text=True).stdout
    out += ["", "## Executed instructions and stall samples per role (tools/ncu_regions.py; unit = one 32-observation chunk)", "", "```", reg.strip(), "```"]
    top = subprocess.run([sys.executable, "tools/ncu_top_stalls.py", rep], capture_output=True, text=True).stdout
    out += ["", "## Instructions with the most stall samples (tools/ncu_top_stalls.py)", "", "```"] + top.strip().splitlines()[:32] + ["```"]
    open(f"{P}/{rnd}_sweep_ncu_full.md", "w").write("\n".join(out) + "\n")

    def num(k):
        v, u = d[k]
        x = float(v.replace(",", ""))
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]

    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    json.dump({"_note": f"dram__bytes_read.sum + dram__bytes_write.sum of k_intr_sweep_ws<0, 8> from the full-size ncu --set full capture "
                        f"{rnd}_sweep_ncu_full.md (100,000 views x 400 points, one launch); read by bench.py for roofline.traffic",
               "c5": {"sweep_dram_bytes_per_launch": int(traffic), "dram_read": int(num("dram__bytes_read.sum")),
                      "dram_write": int(num("dram__bytes_write.sum"))}}, open(f"{P}/traffic.json", "w"), indent=1)
    print("traffic", traffic)
