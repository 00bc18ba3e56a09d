"""Print the metrics that matter from an `ncu --page raw --csv` dump: python tools/ncu_summary.py raw.csv"""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed_op_shared_atom.sum', 'smsp__inst_executed_op_global_red.sum', 'smsp__inst_executed_op_shfl.sum']
STALL = 'smsp__average_warps_issue_stalled_'


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('-' * 100)
        for w in WANT:
            if w in idx:
                print(f"  {w:72s} {r[idx[w]]:>20s} {units[idx[w]]}")
        stalls = [(h, r[i]) for h, i in idx.items() if h.startswith(STALL) and h.endswith('_per_issue_active.ratio')]
        stalls = sorted(stalls, key=lambda kv: -float(kv[1].replace(',', '') or 0))[:7]
        for h, v in stalls:
            print(f"  stall {h[len(STALL):-len('_per_issue_active.ratio')]:40s} {v:>12s}")


if __name__ == '__main__':
    main(sys.argv[1])
