"""Dev tool: hottest SASS instructions of an ncu source-page CSV (ncu -i rep --page source --csv --print-source sass).
usage: python tools/ncu_hot.py src.csv [top] | python tools/ncu_hot.py src.csv range 0xb110 0xe120"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
recs = rows[2:]
def f(r, k):
    try: return float(r[ix[k]].replace(",", "") or 0)
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in recs)
keys = ["stall_wait", "stall_short_sb", "stall_math", "stall_dispatch", "stall_not_selected", "stall_selected", "stall_no_inst", "stall_branch_resolving", "stall_long_sb", "stall_mio"]
if len(sys.argv) > 2 and sys.argv[2] == "range":
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    acc = 0
    for r in recs:
        a = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
        if lo <= a <= hi:
            s = f(r, "# Samples"); acc += s
            top = max(keys, key=lambda k: f(r, k))
            print(f"{a:#7x} {s:7.0f} {100*s/tot:5.2f}% ex={f(r,'Instructions Executed'):.3g} {top[6:]:12s} {r[ix['Source']][:90]}")
    print("range total", acc, f"{100*acc/tot:.1f}%")
else:
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    print("total samples", tot)
    for r in sorted(recs, key=lambda r: -f(r, "# Samples"))[:top]:
        topk = max(keys, key=lambda k: f(r, k))
        print(f"{r[ix['Address']]:>8s} {f(r,'# Samples'):7.0f} {100*f(r,'# Samples')/tot:5.2f}% {topk[6:]:12s} {r[ix['Source']][:100]}")
