#!/usr/bin/env python
"""DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels captured in .ncu-rep files ->
profiles/ncu_traffic.json, which bench.py quotes as roofline.traffic.
Usage: python scripts/ncu_traffic.py OUT.json REPORT.ncu-rep:READS_PER_LAUNCH [...]"""
import csv, json, subprocess, sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
out = {}
for spec in sys.argv[2:]:
    rep, reads = spec.rsplit(":", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        total = sum(float(d[k]) * UNIT[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        name = d["Kernel Name"].split("(")[0].split("::")[-1].split("<")[0]
        out[name] = {"dram_bytes_per_launch": total, "reads_per_launch": int(reads), "duration_us_under_ncu": float(d["gpu__time_duration.sum"]) * TIME_US.get(u["gpu__time_duration.sum"], 1.0),
                     "report": rep.split("/")[-1]}
json.dump(out, open(sys.argv[1], "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
