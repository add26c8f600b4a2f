"""Dev tool: static SASS instruction counts per source line / function region of one kernel.
   cuobjdump -xelf all obj.o ; nvdisasm --print-line-info x.cubin > dis.txt ; python tools/sass_lines.py dis.txt <kernel substring> [top]
The instruction cache of an SM holds ~2000 instructions (32 KB L1.5): code that 16 resident warps execute at different
places has to fit, so this is the number to watch besides registers."""
import re, sys, collections
txt = open(sys.argv[1]).read().split('\n')
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur_fn, cur_line, cur_file = None, None, None
per_line = collections.Counter()
per_fn = collections.Counter()
inline_stack = None
for ln in txt:
    m = re.match(r'\s*\.text\.(\S+):', ln)
    if m:
        cur_fn = m.group(1); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur_file, cur_line = m.group(1).split('/')[-1], int(m.group(2)); continue
    if cur_fn and want in cur_fn and re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+\S', ln):
        per_fn[cur_fn] += 1
        per_line[(cur_file, cur_line)] += 1
for fn, c in per_fn.items():
    print(c, fn[:120])
tot = sum(per_line.values())
print("total", tot, "instructions =", tot * 16 // 1024, "KB")
# bucket by ranges of 20 lines
buckets = collections.Counter()
for (f, l), c in per_line.items():
    buckets[(f, l // 25 * 25)] += c
print("--- by 25-line bucket")
for (f, l), c in sorted(buckets.items(), key=lambda x: -x[1])[:top]:
    print(f"{c:6d}  {f}:{l}-{l+24}")
