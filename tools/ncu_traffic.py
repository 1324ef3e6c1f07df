#!/usr/bin/env python
"""profiles/ncu_summary.json from an ncu CSV holding dram__bytes_read/write.sum per launch:
average DRAM traffic per launch of the dominant kernels (bench.py puts it into roofline.traffic)."""
import collections, csv, io, json, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
by = collections.OrderedDict()
for r in rows:
    by.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.defaultdict(lambda: [0, 0.0])
for d in by.values():
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if m in d:
            tot += d[m][0] * unit.get(d[m][1], 1.0)
    key = "conv_tc" if ("conv3x3_tc_kernel" in d["name"] or "conv3x3_halo_kernel" in d["name"]) else (
        "wgrad_tc" if "wgrad_tc_kernel" in d["name"] else None)
    if key:
        agg[key][0] += 1
        agg[key][1] += tot
out = {f"{k}_dram_bytes_per_launch": v[1] / v[0] for k, v in agg.items()}
out.update({f"{k}_launches_captured": v[0] for k, v in agg.items()})
out["source"] = sys.argv[1]
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)
