"""Top SASS instructions by stall samples from `ncu --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in rows[2:])
print("total samples", tot, "instrs", len(rows) - 2, "executed warp-instr", sum(int(r[ie] or 0) for r in rows[2:]))
top = sorted(range(2, len(rows)), key=lambda k: -int(rows[k][si] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for k in sorted(top):
    r = rows[k]
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{k-2:5d} {int(r[si]):6d} {100*int(r[si])/tot:5.1f}%  ex={r[ie]:>8}  {r[src].strip()[:70]:70s} {st}")
