"""Top stall sites of one kernel in an .ncu-rep: `python tools/ncu_hot.py rep.ncu-rep <launch-index> [n]` (SASS view)."""
import csv
import subprocess
import sys

rep, kid = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
blk = rows[starts[kid]:starts[kid + 1]]
hdr = blk[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in blk[2:] if len(r) == len(hdr)]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print(blk[0][1][:100], "total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = data[i]
    s = int(r[idx["# Samples"]] or 0)
    st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {s:7d} {100.0 * s / max(tot, 1):5.1f}%  {r[idx['Source']][:90]:90s} {st}")
if len(sys.argv) > 4:
    binw = int(sys.argv[4])
    print("bins of", binw, "instructions: start, samples, %")
    for b in range(0, len(data), binw):
        s = sum(int(r[idx["# Samples"]] or 0) for r in data[b:b + binw])
        print(f"  {b:5d} {s:7d} {100.0 * s / tot:5.1f}%")
