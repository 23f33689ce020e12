#!/usr/bin/env python3
"""profiles/lsd_chain_traffic.json (what bench.py reports as roofline.traffic) from an ncu launch list of tools/profile_batch.py:
DRAM bytes, warp instructions and active threads per warp instruction of the LSD region-growing passes, per image.
usage: chain_traffic.py profiles/rNN_launches_raw.csv images > profiles/lsd_chain_traffic.json"""
import csv, json, sys, collections
src, images = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(l for l in open(src, errors="replace") if l.startswith('"')))
h = rows[0]; ix = {n: i for i, n in enumerate(h)}
acc = collections.defaultdict(lambda: collections.defaultdict(float)); n = collections.Counter(); seen = set()
for r in rows[1:]:
    if len(r) < len(h):
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("olf::", "")
    if name not in ("k_lsd_scan", "k_lsd_verify", "k_lsd_grow"):
        continue
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ix["Metric Unit"]], 1.0)
    acc[name][r[ix["Metric Name"]]] += float(r[ix["Metric Value"]].replace(",", "") or 0) * scale
    if (r[ix["ID"]], name) not in seen:
        seen.add((r[ix["ID"]], name)); n[name] += 1
per = {k: (a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]) / images for k, a in acc.items()}
out = {"source": f"{src} (ncu, tools/profile_batch.py: one batch call, {images} images, plain launches)", "images": images,
       "dram_bytes_per_image": sum(per.values()), "per_kernel_dram_bytes_per_image": per,
       "warp_instructions_per_image": {k: a["sm__inst_executed.sum"] / images for k, a in acc.items()},
       "active_threads_per_warp_instruction": {k: a["smsp__thread_inst_executed_per_inst_executed.ratio"] / n[k] for k, a in acc.items()}}
print(json.dumps(out, indent=1))
