#!/usr/bin/env python
"""Key metrics per kernel from `ncu -i X.ncu-rep --page raw --csv`."""
import csv, sys
KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'sm__cycles_elapsed.max']
def main(path):
    rows = list(csv.reader(open(path))); h, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r)); print('##', d['Kernel Name'][:70], d.get('Block Size'), d.get('Grid Size'))
        for k in KEYS:
            if k in d: print(f'   {k:72s} {d[k]:>16s} {units[h.index(k)]}')
        st = []
        for hk in h:
            if hk.startswith('smsp__average_warps_issue_stalled') and hk.endswith('_per_issue_active.ratio') and 'not_issued' not in hk:
                st.append((float(d[hk].replace(',', '')), hk[34:-23]))
        print('   stalls/issue:', ', '.join(f'{k} {v:.2f}' for v, k in sorted(st, reverse=True)[:9]))
if __name__ == '__main__':
    main(sys.argv[1])
