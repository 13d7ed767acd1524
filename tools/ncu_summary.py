"""Read an ncu report on the CPU box (tools only): key metrics per kernel from the raw page, and -- for kernels built in
phases separated by CTA barriers -- executed warp instructions and stall samples per phase from the source page.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    ncu -i gpurun_out/prof.ncu-rep --page source --csv --kernel-name regex:k_assemble_patch > src.csv
    python tools/ncu_summary.py --raw raw.csv --source src.csv
"""
import argparse
import csv

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
       "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def raw_page(path):
    rows = list(csv.reader(open(path)))
    header, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(header)}
    for r in rows[2:]:
        print("----", r[col["Kernel Name"]][:100])
        for m in RAW:
            if m in col and col[m] < len(r) and r[col[m]] != "":
                print("  %-82s %s %s" % (m, r[col[m]], units[col[m]]))


def source_page(path):
    blocks, cur = [], None
    for r in csv.reader(open(path)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["header"] = r
        elif r and cur is not None and "header" in cur:
            cur["rows"].append(r)
    for b in blocks:
        h = b["header"]
        i_exec, i_src, i_smp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
        total = sum(int(r[i_exec]) for r in b["rows"]) or 1
        samples = sum(int(r[i_smp]) for r in b["rows"]) or 1
        print("----", b["name"][:100], "| warp instructions", total, "| SASS lines", len(b["rows"]))
        start, phase = 0, 0
        for i, r in enumerate(b["rows"] + [None]):
            if r is None or "BAR.SYNC" in r[i_src]:
                seg = b["rows"][start:(i + 1 if r is not None else i)]
                if seg:
                    ex = sum(int(x[i_exec]) for x in seg); sm = sum(int(x[i_smp]) for x in seg)
                    print("  phase %d: SASS %4d..%4d  instructions %12d (%5.1f %%)  stall samples %8d (%5.1f %%)"
                          % (phase, start, start + len(seg) - 1, ex, 100.0 * ex / total, sm, 100.0 * sm / samples))
                start, phase = i + 1, phase + 1


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--raw")
    ap.add_argument("--source")
    a = ap.parse_args()
    if a.raw:
        raw_page(a.raw)
    if a.source:
        source_page(a.source)
