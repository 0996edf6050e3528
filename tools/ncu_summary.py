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


def stalls(src, dst):
    """warp-stall breakdown (stalled warps per issue-active cycle, the ncu 'Warp State' section) of every captured launch"""
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
            and "not_issued" not in h]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [hdr[c].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "") for c in cols])
        for r in rows[2:]:
            w.writerow([short(r[ki])] + [r[c] for c in cols])
    print(open(dst).read())


def traffic(dst, tiles_per_launch, *summaries):
    """profiles/<round>_traffic.json (what bench.py quotes as roofline.traffic) from `full` summaries: the LAST launch of
    every kernel class found in them.  usage: traffic <out.json> <tiles per launch> <summary.csv> ..."""
    import json
    classes = [("gemm_qkv", "sched_kernel<5>"), ("gemm_fc1", "sched_kernel<1>"), ("attention", "siglip_attention_pp_kernel"),
               ("gemm_out", "sched_kernel<11>"), ("layernorm", "layernorm_f32_to_bf16_kernel"), ("merge_splice", "merge_splice_kernel"),
               ("preprocess", "resample_fused_kernel")]
    res = {"_source": "ncu --set full --clock-control none, one launch per kernel class of bench.py --steps 1 --warmup 3 "
                      "(tools/profile_round.sh): " + ", ".join(summaries), "tiles_per_launch": int(tiles_per_launch)}
    resid = []
    for sfile in summaries:
        rows = list(csv.reader(open(sfile)))
        hdr = rows[0]
        col = lambda name: next(i for i, h in enumerate(hdr) if h.startswith(name))
        for r in rows[1:]:
            def val(name, scale_units=True):
                c = col(name)
                v = float(r[c])
                unit = hdr[c].split("[")[-1].rstrip("]")
                if scale_units:
                    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
                return v
            e = {"kernel": r[0], "dram_read_mb": val("dram__bytes_read.sum"), "dram_write_mb": val("dram__bytes_write.sum"),
                 "tensor_pipe_active_pct": val("sm__pipe_tensor_cycles_active", False), "xu_pipe_pct": val("sm__inst_executed_pipe_xu", False),
                 "issue_active_pct": val("smsp__issue_active", False), "dram_throughput_pct": val("gpu__dram_throughput", False),
                 "duration_us_under_ncu": val("gpu__time_duration.sum")}
            if "sched_kernel<3>" in r[0]:
                resid.append(e)   # out_proj and fc2 share the kernel: the shorter launch is out_proj (K = 1152)
            for cls, pat in classes:
                if pat in r[0]:
                    res[cls] = e
    if resid:   # EPI_RESID_F32 (<3>): fc2, and out_proj too when RADVLM_B200_OUTPROJ=f32 (the shorter launch, K = 1152);
        resid.sort(key=lambda e: e["duration_us_under_ncu"])   # the default out_proj is <11> (EPI_DELTA_BF16)
        res["gemm_fc2"] = resid[-1]
        if len(resid) > 1 and "gemm_out" not in res:
            res["gemm_out"] = resid[0]
    with open(dst, "w") as f:
        json.dump(res, f, indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], *sys.argv[4:])
    else:
        {"launches": launches, "full": full, "stalls": stalls}[sys.argv[1]](sys.argv[2], sys.argv[3])
