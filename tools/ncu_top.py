"""Summarise an `ncu --page source --csv` dump: per kernel instance, the source lines with the most stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for n, s in enumerate(starts):
    if which is not None and n != which:
        continue
    e = starts[n + 1] if n + 1 < len(starts) else len(rows)
    hdr = rows[s + 1]
    si, ni = hdr.index("Source"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[s + 2:e] if len(r) > ni and r[ni].isdigit()]
    tot = sum(int(r[ni]) for r in body)
    print(f"== [{n}] {rows[s][1][:100]}  total samples {tot}")
    agg = {}
    for r in body:
        k = r[si].strip()[:90]
        a = agg.setdefault(k, [0, {}])
        a[0] += int(r[ni])
        for i, h in stall_cols:
            if r[i].isdigit() and int(r[i]):
                a[1][h] = a[1].get(h, 0) + int(r[i])
    for k, (c, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        ss = ", ".join(f"{h[6:]}={v}" for h, v in sorted(st.items(), key=lambda x: -x[1])[:3])
        print(f"  {100 * c / max(tot, 1):5.1f}%  {k:90s} {ss}")
