#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [extra_metric_substring ...]
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__inst_executed_op_global_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    for r in data:
        print("=" * 100)
        print(r[ki][:140])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:75s} {r[i]:>18s} {units[i]}")
        st = [(float(r[i]), h[len(STALLS):].replace("_per_issue_active.ratio", ""))
              for i, h in enumerate(hdr) if h.startswith(STALLS) and r[i] not in ("", "n/a")]
        print("  stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:9]))
        for e in extra:
            for i, h in enumerate(hdr):
                if e in h:
                    print(f"  {h:75s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
