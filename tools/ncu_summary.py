#!/usr/bin/env python
"""Condense ncu output into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep  > profiles/rNN_<kernel>_full.txt

`launches`: the CSV of `ncu --metrics gpu__time_duration.sum --csv --log-file ...` -> per-kernel
count / total time / share of the captured region (cold-cache and serialised: read SHARES).
`full`: a `--set full` report -> the handful of metrics the roofline discussion uses, per launch.
"""
import collections
import csv
import io
import subprocess
import sys

FULL_KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("sm__cycles_elapsed.max", "cycles elapsed"),
]


def _csv_rows(text):
    lines = [l for l in text.splitlines() if l.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def launches(path):
    rows = _csv_rows(open(path).read())
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    mu = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    mn = hdr.index("Metric Name")
    for r in rows[1:]:
        if r[mn] != "gpu__time_duration.sum":
            continue
        name = r[kn].split("(")[0].replace("void ", "")[:90]
        ns = float(r[mv].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[mu], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} us total (ncu per-launch times: cold cache, serialised)")
    print(f"{'count':>6} {'total_us':>10} {'avg_us':>8} {'share':>7}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[0]:6d} {v[1] / 1e3:10.1f} {v[1] / v[0] / 1e3:8.2f} {100 * v[1] / tot:6.1f}%  {k}")


def full(path):
    """path: a .ncu-rep, or the CSV that `ncu -i rep --page raw --csv` printed (made on the GPU box when the report itself is
    too large to bring back)."""
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = _csv_rows(out)
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: ncu --set full, per profiled launch")
    for r in rows[2:]:
        print(f"kernel: {r[col['Kernel Name']]}   grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for key, label in FULL_KEYS:
            if key in col:
                print(f"    {label:28s} {r[col[key]]} {units[col[key]]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
