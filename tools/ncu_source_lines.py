"""Stall samples and executed instructions per SOURCE LINE of one kernel, from the source page of an .ncu-rep:
    ncu -i report.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_source_lines.py src.csv [N lines, default 40]
(reports with several kernels / source files: the first section only — see profiles/README.md for the summaries)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[2]
iL = 0; iS = 1
cols = {n: i for i, n in enumerate(hdr)}
samp = hdr.index("# Samples"); inst = hdr.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tot = 0; out = []
for r in rows[3:]:
    if r[0] and r[0].isdigit():
        s = int(r[samp]) if r[samp].isdigit() else 0; tot += s
        st = {hdr[i]: int(r[i]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0}
        out.append((int(r[0]), s, int(r[inst]) if r[inst].isdigit() else 0, r[1][:90], st))
print("total samples", tot, "total inst", sum(o[2] for o in out))
for ln, s, ins, src, st in sorted(out, key=lambda o: -o[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    print(f"{ln:4d} {100*s/tot:5.1f}% inst={ins:9d} {src}\n        {top}")
