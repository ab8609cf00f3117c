"""dram traffic per launch of the roofline kernels from `ncu --set full` reports -> profiles/r02_ncu_traffic.json
(read by bench.py).  usage: python profiles/ncu_traffic.py key=report.ncu-rep[:kernel-regex] ... [--out file]"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
res = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    if arg.startswith("--out"):
        out_path = arg.split("=", 1)[1]
        continue
    key, rest = arg.split("=", 1)
    rep, _, rx = rest.partition(":")
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], rows[1]
    picks = [dict(zip(hdr, r)) for r in rows[2:] if not rx or re.search(rx, dict(zip(hdr, r))["Kernel Name"])]
    d = picks[-1]

    def val(name):
        v, u = float(d[name].replace(",", "")), units[hdr.index(name)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    res[key] = {"dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
                "dram_bytes_read": int(val("dram__bytes_read.sum")), "dram_bytes_write": int(val("dram__bytes_write.sum")),
                "kernel": d["Kernel Name"],
                "duration_under_ncu": "%s %s" % (d["gpu__time_duration.sum"], units[hdr.index("gpu__time_duration.sum")]),
                "source": "ncu --set full --clock-control none, %s" % os.path.basename(rep)}
json.dump(res, open(out_path, "w"), indent=1)
print(json.dumps(res, indent=1))
