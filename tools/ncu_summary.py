#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key throughput counters + warp stall breakdown."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__t_bytes.sum", "sm__inst_executed_pipe_uniform.sum", "smsp__inst_issued.sum",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
]


def main(path, which=0):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    d = data[which]
    print("kernel:", d[hdr.index("Kernel Name")][:90], "| launches in report:", len(data))
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:76s} {d[i]:>16s} {units[i]}")
    # FLOPs per launch from the per-op thread-instruction counters (reported per elapsed cycle by this ncu version)
    try:
        cyc = float(d[hdr.index("sm__cycles_elapsed.max")].replace(",", ""))
        cnt = {}
        for op in ("ffma", "fmul", "fadd", "dfma", "dmul", "dadd"):
            cnt[op] = float(d[hdr.index(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")].replace(",", "")) * cyc
        print(f"  derived: fp32_flops_per_launch {2 * cnt['ffma'] + cnt['fmul'] + cnt['fadd']:.0f}")
        print(f"  derived: fp64_flops_per_launch {2 * cnt['dfma'] + cnt['dmul'] + cnt['dadd']:.0f}")
    except (ValueError, KeyError):
        pass
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                stalls.append((float(d[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_") : -len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    print("  warp stall reasons (warps per issue-active cycle):")
    for v, n in sorted(stalls, reverse=True)[:10]:
        print(f"    {n:28s} {v:.2f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
