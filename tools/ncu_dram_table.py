"""Per-kernel table from `ncu --set full` raw-page CSVs (profiles/r02_ncu_full_*.csv): duration,
DRAM bytes per launch, achieved DRAM GB/s, shared-memory wavefront and issue utilisation,
occupancy -- aggregated per (kernel instantiation, grid).  Writes the DRAM bytes per launch,
keyed by the tags of the library's own accounting (f2d_prof_report), to
profiles/r02_kernel_dram_bytes.json for bench.py's `traffic`.
    python tools/ncu_dram_table.py profiles/r02_ncu_full_v1_a.csv [more.csv] [--write]"""
import collections
import csv
import json
import re
import sys

NH = 3
files = [a for a in sys.argv[1:] if not a.startswith("--")]
rows = []
for path in files:
    r = csv.reader(open(path))
    hdr = next(r)
    units = next(r)
    col = {h: k for k, h in enumerate(hdr)}
    ucol = dict(zip(hdr, units))
    for row in r:
        if len(row) == len(hdr):
            rows.append((col, ucol, row))


def num(col, ucol, row, name, default=0.):
    if name not in col:
        return default
    try:
        v = float(row[col[name]].replace(",", ""))
    except ValueError:
        return default
    u = ucol.get(name, "")
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1., "us": 1., "usecond": 1., "ns": 1e-3, "nsecond": 1e-3,
             "ms": 1e3, "msecond": 1e3}.get(u, 1.)
    return v*scale


def tag_of(name, grid):
    """library tag (bench.py kernel table) of an ncu kernel name + grid, at the sizes of a 4096^2 step"""
    g = [int(x) for x in re.findall(r"\d+", grid)]
    m = re.search(r"k_smooth2<\(bool\)(\d), \(bool\)(\d), \(int\)(\d), \(bool\)(\d)>|k_smooth2<(\d), (\d), (\d), (\d)>", name)
    if m:
        v = [x for x in m.groups() if x is not None]
        masked, stored, inp = int(v[0]), int(v[1]), int(v[2])
        mode = 1 if not masked else (0 if stored else 2)
        return "k_smooth2<mode%d,input%d> %dx%d" % (mode, inp, g[0]*64, g[1]*32)
    m = re.search(r"k_resid_restrict<\(bool\)(\d), \(bool\)(\d), \(bool\)(\d)>|k_resid_restrict<(\d), (\d), (\d)>", name)
    if m:
        v = [x for x in m.groups() if x is not None]
        masked, stored = int(v[0]), int(v[1])
        mode = 1 if not masked else (0 if stored else 2)
        return "k_resid_restrict<mode%d> %dx%d" % (mode, g[0]*64, g[1]*32)
    if "k_zsmooth_resid_restrict" in name:   # grid = coarse tiles of 32 x 16
        return "k_zsmooth_rr<mode1> %dx%d" % (g[0]*64, g[1]*32)
    if "k_resid_sumsq_tma" in name:          # tiles of 64 x 32
        return "k_resid_sumsq<mode1> %dx%d" % (g[0]*64, g[1]*32)
    if "k_resid_sumsq" in name:
        return "k_resid_sumsq<mode1> 4096x4096"
    if "k_adv<" in name:
        return "k_adv<upw1,order5,masked0> %dx%d x1" % (g[0]*64, g[1]*32)
    return None


agg = collections.OrderedDict()
for col, ucol, row in rows:
    name = re.sub(r"\(.*$", "", row[col["Kernel Name"]].replace("void ", "").replace("fused::", ""))
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    full = row[col["Kernel Name"]]
    if "k_map_vec" in full or "k_reduce" in full or "k_elementwise" in full:
        m = re.search(r"f2d_(\w+)", full)
        name = re.sub(r"<.*", "", name)+":"+(m.group(1) if m else "?")
    grid = row[col["Grid Size"]]
    key = (name, grid)
    a = agg.setdefault(key, collections.defaultdict(float))
    a["n"] += 1
    a["us"] += num(col, ucol, row, "gpu__time_duration.sum")
    a["rd"] += num(col, ucol, row, "dram__bytes_read.sum")
    a["wr"] += num(col, ucol, row, "dram__bytes_write.sum")
    a["smem"] += num(col, ucol, row, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")
    a["issue"] += num(col, ucol, row, "smsp__issue_active.avg.pct_of_peak_sustained_active",
                      num(col, ucol, row, "sm__inst_executed_realtime.avg.pct_of_peak_sustained_elapsed"))
    a["occ"] += num(col, ucol, row, "sm__warps_active.avg.pct_of_peak_sustained_active")
    a["fp64"] += num(col, ucol, row, "TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")
    a["tag"] = tag_of(full, grid)

print("%-46s %-14s %4s %9s %9s %9s %8s %6s %6s %6s %6s" % ("kernel", "grid", "n", "us", "rd MB", "wr MB", "GB/s", "smem%", "issue%", "occ%", "fp64%"))
out = {}
for (name, grid), a in agg.items():
    n = a["n"]
    us, rd, wr = a["us"]/n, a["rd"]/n, a["wr"]/n
    print("%-46s %-14s %4d %9.2f %9.2f %9.2f %8.0f %6.1f %6.1f %6.1f %6.1f" % (
        name[:46], grid, n, us, rd/1e6, wr/1e6, (rd+wr)/us/1e3 if us else 0., a["smem"]/n, a["issue"]/n, a["occ"]/n, a["fp64"]/n))
    if a["tag"]:
        out[a["tag"]] = rd+wr
if "--write" in sys.argv:
    p = "profiles/r02_kernel_dram_bytes.json"
    json.dump(out, open(p, "w"), indent=1, sort_keys=True)
    print("wrote", p, len(out), "kernels")
