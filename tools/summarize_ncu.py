#!/usr/bin/env python3
"""Summarises an ncu launch list (--csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum[,...]) per kernel:
launches, total time, share, DRAM MB, achieved DRAM GB/s, fraction of the measured HBM peak.  usage: summarize_ncu.py in.csv images [peak_gbs]"""
import csv, json, pathlib, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
images = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
try:
    peak = json.loads((pathlib.Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    peak = 6650.0
if len(sys.argv) > 3:
    peak = float(sys.argv[3])
hdr = rows[0]; ix = {n: i for i, n in enumerate(hdr)}
acc = collections.defaultdict(lambda: collections.defaultdict(float)); launches = collections.Counter(); seen = set()
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("olf::", "")
    metric, unit, val = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", "") or 0)
    if (r[ix["ID"]], name) not in seen:
        seen.add((r[ix["ID"]], name)); launches[name] += 1
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    acc[name][metric] += val * scale
tot = sum(a["gpu__time_duration.sum"] for a in acc.values())
print(f"# {sum(launches.values())} launches, {tot:.1f} us total GPU time (sum of kernels), {images:g} images; peak {peak} GB/s")
print("kernel,launches,total_us,share,us_per_image,dram_read_MB,dram_write_MB,dram_MB_per_image,achieved_dram_GBps,frac_of_measured_hbm_peak")
for name, a in sorted(acc.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    t = a["gpu__time_duration.sum"]; rd, wr = a.get("dram__bytes_read.sum", 0.0), a.get("dram__bytes_write.sum", 0.0)
    gbps = (rd + wr) / (t * 1e-6) / 1e9 if t else 0.0
    print(f"{name},{launches[name]},{t:.1f},{t / tot:.3f},{t / images:.1f},{rd / 1e6:.2f},{wr / 1e6:.2f},{(rd + wr) / 1e6 / images:.2f},{gbps:.1f},{gbps / peak:.4f}")
