"""One ncu report -> (a) the raw page as CSV (one header row, ONE units row, one row per kernel), (b) a short text
summary of the metrics DESIGN.md quotes, value and unit side by side.
usage: ncu_summary.py report.ncu-rep out_prefix      (writes out_prefix_raw.csv and out_prefix_summary.txt)"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__cycles_elapsed.avg',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum']


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(prefix + "_raw.csv", "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    kn = h.index('Kernel Name')
    out = []
    for r in rows[2:]:
        out.append("== %s  (grid %s, block %s)" % (r[kn][:120], r[h.index('Grid Size')], r[h.index('Block Size')]))
        for k in KEYS:
            if k in h:
                i = h.index(k)
                out.append("   %-72s %18s %s" % (k, r[i], units[i]))
        st = []
        for i, name in enumerate(h):
            if 'issue_stalled' in name and name.endswith('_per_issue_active.ratio') and 'average_warps' in name:
                try:
                    st.append((float(r[i].replace(',', '')), name))
                except ValueError:
                    pass
        out.append("   warp stall reasons (warps per issue-active cycle):")
        for v, name in sorted(st, reverse=True)[:8]:
            out.append("      %6.2f  %s" % (v, name.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
    open(prefix + "_summary.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
