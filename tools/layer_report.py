#!/usr/bin/env python
"""Join an ncu launch list (gpu__time_duration per launch) with the engine's GEMM plan dump (MADM_DUMP_PLAN=1 stderr lines)
and print per-layer-class time, TFLOP/s and share.  Usage: layer_report.py launches.csv plan.log"""
import collections
import csv
import re
import sys

launch_csv, plan_log = sys.argv[1], sys.argv[2]
plan = []
seen = 0
for ln in open(plan_log):
    if ln.startswith("MADM_PLAN gemm"):
        plan.append({k: float(v) for k, v in re.findall(r"(\w+)=([0-9.]+)", ln)})
rows = [l for l in open(launch_csv) if not l.startswith("==")]
gemms = []
for row in csv.DictReader(rows):
    if "gemm_tc_kernel" not in row["Kernel Name"]:
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    gemms.append(v / 1e6 if unit == "ns" else (v / 1e3 if unit == "us" else v))
n = len(gemms)
plan = plan[-n:]  # the plan is dumped once per build; keep the last n entries
assert len(plan) == n, (len(plan), n)
agg = collections.OrderedDict()
for p, ms in zip(plan, gemms):
    key = (int(p["M"]), int(p["N"]), int(p["K"]), int(p["bn"]), int(p["taps"]), int(p["res"]), int(p["f32"]), int(p["h16"]), int(p["act"]), int(p.get("tma", 0)), int(p.get("stats", 0)))
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += ms
    a[2] += p["gflop"]
tot = sum(gemms)
print(f"{n} gemm launches, {tot:.3f} ms, {sum(p['gflop'] for p in plan) / tot:.1f} TFLOP/s overall")
print("     M      N      K   bn taps res f32 h16 act tma sts    n      ms   share  TFLOP/s")
for k, (cnt, ms, gf) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:7d} {k[1]:6d} {k[2]:6d} {k[3]:4d} {k[4]:4d} {k[5]:3d} {k[6]:3d} {k[7]:3d} {k[8]:3d} {k[9]:3d} {k[10]:3d} {cnt:4d} {ms:7.3f} {100 * ms / tot:6.1f}% {gf / ms:8.1f}")
