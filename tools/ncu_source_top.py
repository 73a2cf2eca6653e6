"""Developer tool: the instructions of a captured kernel with the most warp-stall samples.
    ncu -i x.ncu-rep --page source --csv --print-source sass > src_sass.csv ; python tools/ncu_source_top.py src_sass.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 45
h = rows[1]
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
data = []
for r in rows[2:]:
    if len(r) < len(h):
        continue
    try:
        s = int(r[isamp])
    except ValueError:
        continue
    data.append((s, r))
tot = sum(s for s, _ in data)
print("total samples", tot, "instructions", len(data))
for s, r in sorted(data, key=lambda x: -x[0])[:top_n]:
    st = {h[i]: int(r[i]) for i in stalls if r[i] not in ("", "0")}
    top = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"{s:6d} {100 * s / tot:5.1f}% ex={r[iex]:>10s} {r[ia][-5:]} {r[isrc][:70]:70s} {top}")
