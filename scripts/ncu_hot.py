"""Top sampled SASS lines of one kernel from an .ncu-rep (needs --import-source on / -lineinfo)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.015
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
end = his[1] if len(his) > 1 else len(rows)
hdr = rows[hi]
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in rows[hi + 1:end] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in data)
print("total samples", tot, "total warp inst", sum(int(r[ie] or 0) for r in data), "sass lines", len(data))
for idx, r in enumerate(data):
    if int(r[si]) > tot * frac:
        print(idx, r[1].strip()[:70], "samples", r[si], "inst", r[ie])
