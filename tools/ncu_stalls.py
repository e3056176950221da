"""Top warp-stall reasons and the hottest SASS lines of the first kernel in an ncu report.
    python tools/ncu_stalls.py report.ncu-rep [n_lines]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r[0].startswith("0x"):
        data.append(r)
tot = sum(int(r[idx["# Samples"]]) for r in data)
print(rows[0][1][:100], "| SASS instructions", len(data), "| samples", tot)
agg = {s: sum(int(r[idx[s]]) for r in data) for s in stalls}
print("  ".join(f"{s[6:]}={v / tot:.2f}" for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:n_lines]:
    st = {s[6:]: int(r[idx[s]]) for s in stalls if int(r[idx[s]]) > 0}
    print(f"{int(r[idx['# Samples']]):6d} {r[idx['Source']].strip()[:70]:70s}", dict(sorted(st.items(), key=lambda kv: -kv[1])[:3]))
