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


# names of the tensor-core launches of one eager step, in launch order (scripts/one_step.py, math mode 1, single lane)
# round 2: the first conv layer's forward has its own kernel (c1::conv1_fwd_kernel); the generic tcgen05 kernel takes the rest
STEP_ORDER = ["conv2_fwd_target", "conv3_fwd_target", "dense1_fwd_target", "conv2_fwd_online", "conv3_fwd_online",
              "dense1_fwd_online", "dense1_wgrad", "dense1_dgrad", "conv3_wgrad", "conv3_dgrad", "conv2_wgrad", "conv2_dgrad", "conv1_wgrad"]
C1_ORDER = ["conv1_fwd_target", "conv1_fwd_online"]


def traffic(tag, path, name):
    """per-launch DRAM bytes (read + write) of the first eager step's kernels -> profiles/<tag>_traffic_<name>.json (bench.py reads it)"""
    import json
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    ik, ir, iw, it = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ur, uw = unit.get(rows[1][ir], 1.0), unit.get(rows[1][iw], 1.0)
    out, tc, c1 = {}, 0, 0
    extra = [m for m in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") if m in hdr]
    for r in rows[2:]:
        b = float(r[ir]) * ur + float(r[iw]) * uw
        if "conv1_fwd_kernel" in r[ik]:
            if c1 < len(C1_ORDER):
                out[C1_ORDER[c1]] = {"dram_bytes": b, "ncu_us": float(r[it]), "kernel": r[ik][:80], **{m.split(".")[0]: float(r[hdr.index(m)]) for m in extra}}
            c1 += 1
        elif "tc_gemm_kernel" in r[ik]:
            if tc < len(STEP_ORDER):
                out[STEP_ORDER[tc]] = {"dram_bytes": b, "ncu_us": float(r[it]), "kernel": r[ik][:80], **{m.split(".")[0]: float(r[hdr.index(m)]) for m in extra}}
            tc += 1
        else:
            key = "gather_rows" if "gather_rows" in r[ik] else ("adam" if "adam" in r[ik] else r[ik].split("(")[0])
            out.setdefault(key, {"dram_bytes": b, "ncu_us": float(r[it]), "kernel": r[ik][:80]})
    try:
        out["_commit"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    except Exception:
        pass
    dst = os.path.join(ROOT, "profiles", f"{tag}_traffic_{name}.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(dst)


if __name__ == "__main__":
    tag, kind, path = sys.argv[1:4]
    name = sys.argv[4] if len(sys.argv) > 4 else os.path.splitext(os.path.basename(path))[0]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    {"launches": launches, "raw": raw, "traffic": traffic}[kind](tag, path, name)
