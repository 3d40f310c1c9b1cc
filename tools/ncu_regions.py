"""Executed-instruction profile of a kernel from an .ncu-rep source page: splits the SASS at given
markers (opcode substrings) into regions and prints, per region, executed warp instructions, share of
stall samples and the opcode histogram weighted by execution count.
Usage: python tools/ncu_regions.py REP [units]   (units = work units per launch, e.g. 32-observation chunks)"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
idx = {k: i for i, k in enumerate(h)}
data = rows[2:]
tot_s = sum(int(r[idx["# Samples"]] or 0) for r in data)
tot_i = sum(int(r[idx["Instructions Executed"]] or 0) for r in data)
print(f"{rows[0][1]}: {tot_i} warp instructions executed ({tot_i / units:.1f} per unit), {tot_s} samples")


def op(src):
    t = src.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0] if t else "?"


# regions: split at USETMAXREG (role branches) - prologue | FP64 role | Jacobian role
regions, cur, name = [], [], "prologue"
for r in data:
    src = r[idx["Source"]]
    if "USETMAXREG" in src:
        regions.append((name, cur))
        cur = []
        name = "fp64 role" if "TRY_ALLOC" in src else "jacobian role"
    cur.append(r)
regions.append((name, cur))
for name, rs in regions:
    n_i = sum(int(r[idx["Instructions Executed"]] or 0) for r in rs)
    n_s = sum(int(r[idx["# Samples"]] or 0) for r in rs)
    hist = collections.Counter()
    for r in rs:
        hist[op(r[idx["Source"]])] += int(r[idx["Instructions Executed"]] or 0)
    st = collections.Counter()
    for k in h:
        if k.startswith("stall_") and "Not Issued" not in k:
            st[k] = sum(int(r[idx[k]] or 0) for r in rs)
    print(f"\n== {name}: {n_i} executed ({n_i / units:.1f} per unit), samples {n_s} ({100.0 * n_s / max(tot_s, 1):.1f} %)")
    print("   opcodes per unit: " + ", ".join(f"{k} {v / units:.1f}" for k, v in hist.most_common(28)))
    print("   stalls: " + ", ".join(f"{k[6:]} {100.0 * v / max(n_s, 1):.0f}%" for k, v in st.most_common(8)))
