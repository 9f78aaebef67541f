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


# ---- one captured step (scripts/gpu_profile.sh): kernels in launch order + the engine's own ordered scope names (step_scopes.txt) ----
def _expected(scope):
    """the kernel that opens a scope; follow-on kernels of a scope (split-K reduction of an unnamed kind, ...) stay with it"""
    if scope == "sumtree_sample": return "sample_kernel"
    if scope == "sumtree_update": return "tree_update_kernel"
    if scope == "gather_rows": return "gather_rows_kernel"
    if scope == "dgrad_merge_weights": return "dgrad_merge_weights_kernel"
    if scope == "splitk_reduce": return "splitk_reduce_kernel"
    if scope.startswith("heads_fwd"): return "heads_fwd_kernel"
    if scope == "head_loss": return "head_loss"
    if scope == "heads_dgrad": return "heads_dgrad_kernel"
    if scope.endswith("_bgrad"): return "colsum_kernel"
    if scope == "adam": return "adam_kernel"
    return ("tc_gemm_kernel", "conv1_fwd_kernel", "igemm_kernel")       # a contraction


def _align(kernel_names, scopes):
    """scope name per kernel row: a row that matches the NEXT scope's opening kernel starts that scope"""
    out, si = [], -1
    for k in kernel_names:
        if si + 1 < len(scopes):
            exp = _expected(scopes[si + 1])
            exp = exp if isinstance(exp, tuple) else (exp,)
            if any(e in k for e in exp):
                si += 1
        out.append(scopes[max(si, 0)])
    return out


def _scopes(path):
    return [l.split()[0] for l in open(path) if l.strip()]


def step(tag, path, name, scopes_path=None):
    """full-set capture of ONE step -> profiles/<tag>_ncu_<name>.csv (selected metrics per launch, scope-labelled) and
    profiles/<tag>_traffic_<name>.json (per-scope DRAM bytes per launch; bench.py reads it for roofline.traffic)"""
    import json
    scopes = _scopes(scopes_path or os.path.join(ROOT, "gpurun_out", "r02_step_scopes.txt"))
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    ik = hdr.index("Kernel Name")
    data = rows[2:]
    labels = _align([r[ik] for r in data], scopes)
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    out = os.path.join(ROOT, "profiles", f"{tag}_ncu_{name}.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["scope", "kernel"] + [m for m, _ in cols])
        w.writerow(["", "(unit)"] + [rows[1][i] for _, i in cols])
        for lab, r in zip(labels, data):
            w.writerow([lab, r[ik][:80]] + [r[i] for _, i in cols])
    print(out)
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ur, uw = unit.get(rows[1][ir], 1.0), unit.get(rows[1][iw], 1.0)
    tu = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(rows[1][it], 1.0)
    extra = [m for m in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") if m in hdr]
    tr = {}
    for lab, r in zip(labels, data):
        main = any(e in r[ik] for e in (_expected(lab) if isinstance(_expected(lab), tuple) else (_expected(lab),)))
        key = lab if main else lab + "+" + r[ik].split("(")[0].split("::")[-1].split("<")[0]
        b = float(r[ir]) * ur + float(r[iw]) * uw
        tr.setdefault(key, {"dram_bytes": b, "ncu_us": float(r[it]) * tu, "kernel": r[ik][:80], **{m.split(".")[0]: float(r[hdr.index(m)]) for m in extra}})
    tr["gather_rows"] = tr.get("gather_rows", {})
    try:
        tr["_commit"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    except Exception:
        pass
    tr["_note"] = "ncu --set full, --clock-control none, one eager single-lane step (cold caches, serialised): per-launch DRAM bytes are comparable with the algorithmic bytes, the times only as shares"
    dst = os.path.join(ROOT, "profiles", f"{tag}_traffic_{name}.json")
    json.dump(tr, open(dst, "w"), indent=1)
    print(dst)


def steplist(tag, path, name, scopes_path=None):
    """launch list (gpu__time_duration only) of ONE step, scope-labelled -> profiles/<tag>_launches_<name>.csv"""
    scopes = _scopes(scopes_path or os.path.join(ROOT, "gpurun_out", "r02_step_scopes.txt"))
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > rows[hi].index("Metric Value")]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    labels = _align([r[ik] for r in data], scopes)
    tot = sum(float(r[iv].replace(",", "")) for r in data)
    out = os.path.join(ROOT, "profiles", f"{tag}_launches_{name}.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["#", "scope", "kernel", "grid", "block", "us", "share_of_step"])
        for i, (lab, r) in enumerate(zip(labels, data)):
            v = float(r[iv].replace(",", ""))
            w.writerow([i, lab, r[ik].split("(")[0][:70], r[ig], r[ib], round(v / 1e3, 2), round(v / tot, 4)])
        w.writerow(["", "TOTAL (serialised, cold caches, under ncu)", "", "", "", round(tot / 1e3, 1), 1.0])
    print(out)


if __name__ == "__main__":
    tag, kind, path = sys.argv[1:4]
    name = sys.argv[4] if len(sys.argv) > 4 else os.path.splitext(os.path.basename(path))[0]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    {"launches": launches, "raw": raw, "step": step, "steplist": steplist}[kind](tag, path, name)
