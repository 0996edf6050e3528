#!/usr/bin/env python
"""Summarise ncu output for profiles/ (run in the container, ncu reads the .ncu-rep / csv brought back by gpurun).

  launch list : python tools/ncu_summary.py launches <ncu --csv log> <out.csv>
                per-kernel launch count, total time and share of all profiled launches
  full report : python tools/ncu_summary.py full <report.ncu-rep> <out.csv>
                one row per captured launch with the metrics DESIGN.md / bench.py's roofline cite
"""
import csv
import io
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
]


def short(name):
    name = name.replace("void ", "").replace("rv::", "")
    return name.split("(")[0]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.1f,%.3f\n' % (k, a[0], a[1], a[1] / tot))
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(m) for m in FULL_METRICS if m in hdr]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + ["%s [%s]" % (hdr[c], units[c]) for c in cols])
        for r in rows[2:]:
            w.writerow([short(r[ki])] + [r[c] for c in cols])
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
