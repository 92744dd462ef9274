"""Helpers to read .ncu-rep files on the CPU box (ncu -i ... --page raw/source --csv).

    python profiles/ncu_tools.py summary gpurun_out/prof.ncu-rep
    python profiles/ncu_tools.py lines   gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def summary(rep):
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")])
        for w in RAW:
            if w in hdr:
                print(f"  {w:85s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")


def lines(rep, top=40):
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
    hdr = rows[hi]
    ie, ws = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    agg, stall, src, cur, tot, tots = {}, {}, {}, None, 0, 0
    for r in rows[hi + 1:]:
        if "Instructions Executed" in r:
            break
        if len(r) <= ie:
            continue
        if r[0] != "":
            cur = r[0]
            src[cur] = r[1]
            continue
        try:
            v, s = int(r[ie]), int(r[ws])
        except ValueError:
            continue
        agg[cur] = agg.get(cur, 0) + v
        stall[cur] = stall.get(cur, 0) + s
        tot += v
        tots += s
    print(f"total instructions {tot}, stall samples {tots}")
    for k, v in sorted(agg.items(), key=lambda kv: -stall[kv[0]])[:top]:
        print(f"{v:>11d} {100 * v / tot:5.1f}% inst {100 * stall[k] / max(tots, 1):5.1f}% stall  L{k}: {src[k].strip()[:100]}")


if __name__ == "__main__":
    cmd, rep = sys.argv[1], sys.argv[2]
    if cmd == "summary":
        summary(rep)
    else:
        lines(rep, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
