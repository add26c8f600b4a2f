"""Dev tool: instruction mix and stall samples per opcode from an ncu report's source page.
Usage: python tools/ncu_opmix.py report.ncu-rep [n_warp_steps]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
ex = collections.Counter(); samp = collections.Counter(); stall = collections.defaultdict(collections.Counter)
tot = 0
for r in rows:
    src = r["Source"].strip()
    if src.startswith("@"): src = src.split(None, 1)[1]
    op = src.split()[0].split(".")[0] if src else "?"
    n = int(r["Instructions Executed"] or 0)
    ex[op] += n; tot += n
    samp[op] += int(r["# Samples"] or 0)
    for k, v in r.items():
        if k.startswith("stall_") and "Not Issued" not in k and v and v != "0":
            stall[op][k[6:]] += int(v)
allsamp = sum(samp.values())
print(f"total warp instructions {tot:.4g}" + (f"  = {tot/steps:.1f} per warp-step" if steps else ""))
for op, n in ex.most_common(28):
    top = ", ".join(f"{k}:{v}" for k, v in stall[op].most_common(3))
    print(f"{op:10s} {n:14d} {100*n/tot:6.2f}%" + (f" {n/steps:7.1f}/step" if steps else "") + f"   samples {100*samp[op]/max(allsamp,1):5.1f}%  [{top}]")
