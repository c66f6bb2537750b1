#!/usr/bin/env python
"""(round 1 helper; round 2 writes profiles/traffic.json per config, see profiles/README.md) Per-launch DRAM traffic of every kernel in an `ncu --set full` report -> profiles/traffic.json
(dram__bytes_read.sum + dram__bytes_write.sum, bytes; the last launch of each kernel name wins).
usage: python profiles/ncu_traffic.py <rep> [<rep> ...]"""
import csv
import json
import os
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
traffic = json.load(open(out_path)) if os.path.exists(out_path) else {}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    for r in rows[2:]:
        name = re.sub(r"<.*", "", r[ki].replace("void ", "").replace("<unnamed>::", "")).split("(")[0].strip()
        traffic[name] = float(r[ri]) * UNIT[units[ri]] + float(r[wi]) * UNIT[units[wi]]
json.dump(traffic, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(traffic, indent=1, sort_keys=True))
