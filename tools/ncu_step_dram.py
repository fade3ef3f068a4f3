"""DRAM bytes of ONE step from an ncu pass over every kernel of the step
(`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`, step run without
CUDA graphs so that every kernel is listed): per kernel (name, grid) launches, mean duration,
bytes per launch; and the step total, written to profiles/r02_step_dram_bytes.json for
bench.py's `step_hbm.step_dram_bytes`.
    python tools/ncu_step_dram.py gpurun_out/r02_step_dram.csv [--write]"""
import collections
import csv
import json
import re
import sys

path = sys.argv[1]
rows = collections.OrderedDict()   # launch id -> dict
with open(path) as f:
    lines = [ln for ln in f if ln.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
col = {h: k for k, h in enumerate(hdr)}
for row in r:
    if len(row) != len(hdr):
        continue
    d = rows.setdefault(row[col["ID"]], {"name": row[col["Kernel Name"]], "grid": row[col["Grid Size"]]})
    v = float(row[col["Metric Value"]].replace(",", ""))
    u = row[col["Metric Unit"]]
    v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1., "ns": 1e-3, "us": 1., "usecond": 1., "nsecond": 1e-3,
          "msecond": 1e3, "ms": 1e3}.get(u, 1.)
    d[row[col["Metric Name"]]] = v


def short(name):
    name = re.sub(r"\(.*$", "", name.replace("void ", ""))
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|fused::|ptail::|ctail::", "", name)
    return re.sub(r"\((bool|int)\)", "", name)


agg = collections.OrderedDict()
tot_b = tot_us = 0.
for d in rows.values():
    k = (short(d["name"]), d["grid"])
    a = agg.setdefault(k, [0, 0., 0., 0.])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.)
    a[2] += d.get("dram__bytes_read.sum", 0.)
    a[3] += d.get("dram__bytes_write.sum", 0.)
    tot_b += d.get("dram__bytes_read.sum", 0.)+d.get("dram__bytes_write.sum", 0.)
    tot_us += d.get("gpu__time_duration.sum", 0.)
print("%-58s %-14s %5s %9s %10s %10s %8s" % ("kernel", "grid", "n", "us each", "rd MB each", "wr MB each", "GB/s"))
out = []
for (name, grid), (n, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-58s %-14s %5d %9.2f %10.2f %10.2f %8.0f" % (name[:58], grid, n, us/n, rd/n/1e6, wr/n/1e6, (rd+wr)/us/1e3 if us else 0.))
    out.append({"kernel": name, "grid": grid, "launches": n, "us_each": us/n, "dram_bytes_each": (rd+wr)/n})
print("step: %d kernels, %.1f us under ncu (serialised, cold), %.3f GB of DRAM traffic" % (len(rows), tot_us, tot_b/1e9))
if "--write" in sys.argv:
    p = "profiles/r02_step_dram_bytes.json"
    json.dump({"source": path.split("/")[-1], "step_total_bytes": tot_b, "kernels_in_step": len(rows),
               "step_us_under_ncu": tot_us, "kernels": out}, open(p, "w"), indent=1)
    print("wrote", p)
