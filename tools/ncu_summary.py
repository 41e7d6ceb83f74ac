"""Reads `ncu -i <report> --page raw --csv` and prints one block per captured launch with the counters the roofline
arithmetic needs (B200_PROFILING.md): duration, DRAM bytes / %, L2, tensor pipe, issue, occupancy, registers.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv | python tools/ncu_summary.py [name-filter]
"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units, data = rows[0], rows[1], rows[2:]
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    name_i = hdr.index("Kernel Name")
    for r in data:
        if flt not in r[name_i]:
            continue
        print(f"== {r[name_i][:110]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:82s} {units[i]:>16s} {r[i]}")


if __name__ == "__main__":
    main()
