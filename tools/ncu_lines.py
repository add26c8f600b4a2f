"""Dev tool: per-SOURCE-line totals of an ncu source-page CSV made with --print-source cuda (or sass,cuda):
   ncu -i rep --page source --csv --print-source cuda > src.csv ; python tools/ncu_lines.py src.csv [top]
Prints the lines with most warp samples and most executed instructions, and totals per source file."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
# find the header row (contains "# Samples" or "Samples")
h = next(i for i, r in enumerate(rows) if any("Samples" in c for c in r))
hdr = rows[h]
ix = {c: i for i, c in enumerate(hdr)}
def col(*names):
    for n in names:
        for c in hdr:
            if c.strip() == n: return ix[c]
    return None
c_s = col("# Samples", "Warp Stall Sampling (All Samples)", "Samples")
c_e = col("Instructions Executed")
c_src = col("Source")
c_line = col("#", "Line")
c_file = col("File Path", "File")
def f(r, i):
    try: return float(r[i].replace(",", "") or 0)
    except Exception: return 0.0
recs = [r for r in rows[h + 1:] if len(r) == len(hdr)]
tot_s = sum(f(r, c_s) for r in recs) or 1.0
tot_e = sum(f(r, c_e) for r in recs) or 1.0
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("columns:", hdr[:12])
print(f"total samples {tot_s:.0f}  total warp-instructions executed {tot_e:.4g}")
print("--- by samples")
for r in sorted(recs, key=lambda r: -f(r, c_s))[:top]:
    print(f"{r[c_line] if c_line is not None else '':>6s} {100*f(r,c_s)/tot_s:6.2f}% smp {100*f(r,c_e)/tot_e:6.2f}% inst  {r[c_src][:110]}")
