#!/usr/bin/env python
"""Per-source-line stall samples: joins `ncu --page source --csv` (SASS rows) with `nvdisasm -g -c` line info.
   python tools/ncu_lines.py <src.csv> <nvdisasm.txt> <kernel-substring> <source.cu> [top]"""
import collections, csv, re, sys
csvp, disp, kname, srcp = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
cur = None; ins = []; inside = False
for l in open(disp):
    if l.startswith("//----") and ".text." in l:
        inside = kname in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m:
        ins.append((cur, m.group(2)))
rows = list(csv.reader(open(csvp)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) > 2]
ix = {h: i for i, h in enumerate(hdr)}
assert len(data) == len(ins), (len(data), len(ins))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for (cur, txt), r in zip(ins, data):
    a = agg[cur]
    a[0] += int(r[ix['# Samples']]); a[1] += int(r[ix['Instructions Executed']])
    for s in stalls:
        a[2][s[6:]] += int(r[ix[s]])
tot = sum(a[0] for a in agg.values()); print('total samples', tot, 'total instr', sum(a[1] for a in agg.values()))
src = open(srcp).read().splitlines()
base = srcp.split('/')[-1]
for (f, ln), a in sorted(sorted(agg.items(), key=lambda kv: -kv[1][0])[:top], key=lambda kv: (kv[0][0], kv[0][1])):
    text = src[ln - 1].strip()[:78] if f == base and ln <= len(src) else f
    print(str(ln).rjust(4), str(a[0]).rjust(6), str(a[1]).rjust(10), str(a[2].most_common(2))[:56].ljust(56), '|', text)
