#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into a few lines.
usage: python profiles/ncu_summary.py file.ncu-rep [more-metric-substrings...]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum ',
        'l1tex__data_pipe_lsu_wavefronts.sum ', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum ', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ',
        'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum ', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum ',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg ', 'smsp__issue_active.avg.pct',
        'sm__inst_executed_pipe_lsu', 'sm__inst_executed_pipe_fma', 'sm__pipe_fma_cycles_active.avg.pct', 'sm__inst_executed_pipe_alu',
        'smsp__average_warps_issue_stalled', 'lts__t_sectors_op_red.sum ', 'lts__t_sectors_op_atom.sum ', 'lts__t_sectors_srcunit_tex_op_read.sum ',
        'smsp__cycles_active.avg ', 'sm__cycles_active.avg ', 'l1tex__lsu_writeback_active', 'smsp__inst_executed_op_shared', 'lts__t_bytes.sum ',
        'l1tex__m_xbar2l1tex_read_bytes.sum ', 'l1tex__m_l1tex2xbar_write_bytes.sum ']


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
        print('==', name[:110])
        for h, u, v in zip(hdr, units, r):
            if any((h + ' ').startswith(k) or k.strip() in h for k in KEYS + extra):
                print('  %-75s %-12s %s' % (h, u, v))


if __name__ == '__main__':
    main()
