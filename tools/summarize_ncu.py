#!/usr/bin/env python
"""Turn ncu outputs into the small, committed summaries under profiles/.

    python tools/summarize_ncu.py launches <launches.csv> <out.md>      # per-kernel time shares
    python tools/summarize_ncu.py report   <file.ncu-rep> <out.md>      # key metrics + top stall lines
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(src, dst):
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({len(rows)} launches, gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.2f} % |\n")
        ours = sum(v[1] for k, v in agg.items() if k.startswith("k_"))
        f.write(f"\nOur kernels (k_*): {100 * ours / tot:.2f} % of the device time; total {tot / 1e6:.2f} ms.\n")


def report(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of `{src.split('/')[-1]}`\n\n")
        for k in range(2, len(rows)):
            name = rows[k][hdr.index("Kernel Name")]
            f.write(f"## launch {k - 2}: `{name[:150]}`\n\n| metric | unit | value |\n|---|---|---:|\n")
            for w in KEYS:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"| {w} | {rows[1][i]} | {rows[k][i]} |\n")
            f.write("\n")
        src_csv = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src_csv.splitlines()))
        if len(srows) > 2 and "Instructions Executed" in srows[1]:
            h = srows[1]
            ia, isrc, ist = h.index("Instructions Executed"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
            data = []
            for r in srows[2:]:
                if len(r) <= ia:
                    break
                try:
                    data.append((int(r[ist]), int(r[ia]), r[isrc]))
                except ValueError:
                    continue
            tot = sum(d[0] for d in data) or 1
            f.write("## top warp-stall SASS lines (first launch)\n\n| stall samples | executed | SASS |\n|---:|---:|---|\n")
            for s_, n, t in sorted(data, reverse=True)[:25]:
                f.write(f"| {100 * s_ / tot:.1f} % | {n} | `{t.strip()[:110]}` |\n")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
