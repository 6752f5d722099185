#!/usr/bin/env python
"""DRAM bytes (read + write) per launch of every kernel in a full ncu capture -> profiles/traffic.json
    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep > profiles/traffic.json"""
import csv, io, json, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
res = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = re.sub(r"^(void )?(dagb200::)?(\w+::)?", "", d["Kernel Name"]).split("(")[0]
    name = name.replace("dagb200::", "")
    tot = sum(to_bytes(d[k], units[hdr.index(k)]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    res.setdefault(name, int(tot))
print(json.dumps(res, indent=1))
