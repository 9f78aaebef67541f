"""Turn gpurun_out/ ncu artefacts into the small tracked summaries under profiles/.
    python scripts/summarize_profiles.py <round-tag> launches gpurun_out/launches_m1.csv [name]
    python scripts/summarize_profiles.py <round-tag> raw gpurun_out/prof_tc_m1.ncu-rep [name]"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def launches(tag, path, name):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= iv:
            continue
        key = (r[ik].split("(")[0], r[ig], r[ib])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    out = os.path.join(ROOT, "profiles", f"{tag}_launches_{name}.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "block", "launches", "avg_us", "total_us", "share_of_capture"])
        for (k, g, b), v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, g, b, v[0], round(v[1] / v[0] / 1e3, 2), round(v[1] / 1e3, 1), round(v[1] / tot, 4)])
    print(out)


def raw(tag, path, name):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    ik = hdr.index("Kernel Name")
    out = os.path.join(ROOT, "profiles", f"{tag}_ncu_{name}.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [m for m, _ in cols])
        w.writerow(["(unit)"] + [rows[1][i] for _, i in cols])
        for r in rows[2:]:
            w.writerow([r[ik][:90]] + [r[i] for _, i in cols])
    print(out)


if __name__ == "__main__":
    tag, kind, path = sys.argv[1:4]
    name = sys.argv[4] if len(sys.argv) > 4 else os.path.splitext(os.path.basename(path))[0]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    (launches if kind == "launches" else raw)(tag, path, name)
