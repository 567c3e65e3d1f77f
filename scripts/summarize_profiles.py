"""Turns the scratch captures under gpurun_out/ into the tracked summaries under profiles/.

  python scripts/summarize_profiles.py r1b      # suffix for the file names

Reads gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum launch list of `bench.py --steps 2 --warmup 3`),
gpurun_out/decode_full.ncu-rep (ncu --set full capture of the three decode kernels of one 1 GiB step) and the bench JSON
lines, and writes launches_<tag>.csv, launches_<tag>_summary.json, decode_full_<tag>_summary.json, roofline_traffic.json
and bench_<tag>_{ours,reference}.json."""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

# launch list
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, "launches_%s.csv" % tag))
lines = [l for l in open(os.path.join(G, "launches.csv")) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = collections.OrderedDict()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    v = float(r["Metric Value"]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r["Metric Unit"], 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
json.dump([{"kernel": k, "launches": v[0], "total_ms": round(v[1], 3), "avg_us": round(v[1] / v[0] * 1e3, 2)} for k, v in agg.items()],
          open(os.path.join(P, "launches_%s_summary.json" % tag), "w"), indent=1)

# full capture
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", os.path.join(G, "decode_full.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
out, traffic = [], {}
for row in rr[2:]:
    d = dict(zip(hdr, row)); u = dict(zip(hdr, units))
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("lzb::", "")
    e = {"kernel": name}
    for k in KEYS:
        if k in d:
            e[k] = "%s %s" % (d[k], u.get(k, ""))
    out.append(e)
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tr = sum(float(d[k]) * scale.get(u[k], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    traffic[{"k_fse_literals": "literals", "k_fse_lmds": "lmds", "k_expand": "expand"}.get(name, name)] = int(tr)
json.dump(out, open(os.path.join(P, "decode_full_%s_summary.json" % tag), "w"), indent=1)
json.dump(traffic, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
for src, dst in (("bench_ours.json", "bench_%s_ours.json" % tag), ("bench_reference.json", "bench_%s_reference.json" % tag)):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
print(json.dumps(traffic), len(out), "kernels in the full capture")
