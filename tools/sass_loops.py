"""Dev tool: list the loops (backward branches) of one kernel's SASS with their instruction mix.
usage: cuobjdump -sass -fun <mangled> lib.so > f.txt; python tools/sass_loops.py f.txt [min_len]"""
import re, sys, collections
ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_idx:
            j = addr_idx[tgt]
            body = ins[j:i + 1]
            if len(body) < min_len:
                continue
            ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
            print(f"loop {tgt:#x}..{a:#x}: {len(body)} instrs  " + " ".join(f"{k}:{v}" for k, v in ops.most_common(14)))
