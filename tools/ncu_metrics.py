"""Print selected metrics of an .ncu-rep (raw page) — usage: python tools/ncu_metrics.py file.ncu-rep [regex...]"""
import csv
import re
import subprocess
import sys

DEFAULT = [r'^gpu__time_duration.sum$', r'^launch__grid_size$', r'^launch__cluster', r'^launch__occupancy_cluster',
           r'^launch__registers_per_thread$', r'^dram__bytes_read.sum$', r'^dram__bytes_write.sum$',
           r'^gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed$', r'^lts__throughput.avg.pct_of_peak_sustained_elapsed$',
           r'^l1tex__throughput.avg.pct_of_peak_sustained_elapsed$', r'^sm__throughput.avg.pct_of_peak_sustained_elapsed$',
           r'^sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active$', r'^sm__cycles_elapsed.avg$',
           r'^sm__cycles_active.avg$', r'^sm__inst_executed_pipe_tensor', r'^sm__pipe_tensor_subpipe.*cycles_active.avg.pct',
           r'^smsp__issue_active.avg.pct_of_peak_sustained_active$', r'^sm__warps_active.avg.pct_of_peak_sustained_active$',
           r'^smsp__inst_executed.sum$', r'^sm__ctas_launched.sum$', r'^l1tex__m_xbar2l1tex_read_bytes.sum$',
           r'^lts__t_sector_hit_rate.pct$', r'^sm__cycles_active.max$', r'^sm__cycles_active.min$']


def main():
    path = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
        print('==', name[:100])
        for i, h in enumerate(hdr):
            if any(re.search(p, h) for p in pats):
                print(f'  {h:78s} {vals[i]:>18s} {units[i]}')


if __name__ == '__main__':
    main()
