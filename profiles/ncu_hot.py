"""Hottest SASS instructions of an ncu report by warp-stall samples (source page): python profiles/ncu_hot.py report.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; body = rows[hi + 1:]
iS, iN, iX = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[iN] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][iN] or 0))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[j] or 0), h[j][6:]) for j in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[iN]):6d} {100*int(r[iN])/tot:5.1f}%  exec={r[iX]:>9s}  {r[iS].strip()[:70]:70s} {top}")
