"""Top stalled SASS instructions of an .ncu-rep (source page, needs -lineinfo + --import-source on).
Usage: python tools/ncu_top_stalls.py REP [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
idx = {k: i for i, k in enumerate(h)}
data = rows[2:]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print(rows[0][1], "total samples", tot)
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[idx[k]] or 0) for r in data) for k in stalls}
print("stall totals:", sorted(agg.items(), key=lambda x: -x[1])[:8])
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:n]:
    s = int(r[idx["# Samples"]])
    st = sorted(((k, int(r[idx[k]] or 0)) for k in stalls), key=lambda x: -x[1])[:2]
    print(r[idx["Address"]][-5:], f"{s:6d} {100 * s / tot:5.1f}%", r[idx["Source"]][:72].ljust(72), st)
