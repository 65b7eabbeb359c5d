#!/usr/bin/env python
"""Summarise an `ncu --csv` launch list (ncu -k regex:ddn --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum): the kernels of the last bench step, per kernel ms and DRAM GB.
  python scripts/summarise_launches.py launches.csv [first-kernel-of-a-step (default build_pair_tables)] [n]
n: which segment, counted from the end (default 1 = the last; bench.py's trailing K4-alone measurement starts with
build_pair_tables too, so the timed step of a default bench capture is n = 2)."""
import collections
import csv
import io
import sys

rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
first = sys.argv[2] if len(sys.argv) > 2 else "build_pair_tables"
r = list(csv.DictReader(io.StringIO("".join(rows))))
byid = collections.OrderedDict()
for x in r:
    d = byid.setdefault(x["ID"], {"name": x["Kernel Name"].split("(")[0].replace("void ", "").replace("ddn::", "")[-44:], "grid": x["Grid Size"]})
    d[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
    d["unit_" + x["Metric Name"]] = x["Metric Unit"]
ids = list(byid)
starts = [i for i, k in enumerate(ids) if first in byid[k]["name"]]
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 1
last = starts[-nth] if len(starts) >= nth else 0
end = starts[-nth + 1] if nth > 1 and len(starts) >= nth else len(ids)
tot = 0.0
sc = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
ts = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}
for k in ids[last:end]:
    d = byid[k]
    t_ms = d["gpu__time_duration.sum"] * ts[d["unit_gpu__time_duration.sum"]]
    rd = d.get("dram__bytes_read.sum", 0) * sc.get(d.get("unit_dram__bytes_read.sum", "byte"), 1e-9)
    wr = d.get("dram__bytes_write.sum", 0) * sc.get(d.get("unit_dram__bytes_write.sum", "byte"), 1e-9)
    print(f"{t_ms:8.3f} ms  rd {rd:6.3f} GB  wr {wr:6.3f} GB  {(rd + wr) / t_ms * 1e3:6.0f} GB/s  {d['name']}  grid {d['grid']}")
    tot += t_ms
print(f"{tot:8.3f} ms total")
