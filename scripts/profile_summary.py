"""Turns gpurun_out/*.ncu-rep / launch-list CSVs into the tracked summaries under profiles/.

    python scripts/profile_summary.py rep  gpurun_out/prof_knn2_tc_r1.ncu-rep  profiles/r1_knn2_tc
    python scripts/profile_summary.py list gpurun_out/launches_cfg3.csv       profiles/r1_launches_cfg3
"""
import collections
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out + "_ncu.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")]
            for h, u, v in zip(hdr, units, vals):
                if h in KEEP:
                    w.writerow([name[:60], h, u, v])
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    idx = {h: j for j, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) > 10]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
    with open(out + "_hotspots.txt", "w") as f:
        f.write(f"# {path}: {tot} warp-state samples; stall totals then the 25 hottest SASS instructions\n")
        agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in stalls}
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
            f.write(f"{k:28s} {v:8d} {100.0 * v / max(tot, 1):5.1f}%\n")
        f.write("\n")
        for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:25]:
            st = sorted([(h, int(r[idx[h]] or 0)) for h in stalls if int(r[idx[h]] or 0) > 0], key=lambda kv: -kv[1])[:3]
            f.write(f"{r[idx['Address']][-5:]} samples={r[idx['# Samples']]:>6s} exec={r[idx['Instructions Executed']]:>9s}  "
                    f"{r[idx['Source']][:70]:70s} {st}\n")


def launch_list(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[i_val].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[i_unit], v)
        a = agg.setdefault(r[i_name][:80], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out + ".txt", "w") as f:
        f.write(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised: compare shares)\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{a[1]:11.1f} us {a[0]:5d}x avg {a[1] / a[0]:9.1f} us {100 * a[1] / tot:5.1f}%  {k}\n")
    import shutil
    shutil.copy(path, out + ".csv")


if __name__ == "__main__":
    {"rep": rep, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
