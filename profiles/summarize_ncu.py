#!/usr/bin/env python
"""One line per profiled launch from an `ncu --set full` report.
usage: ncu -i report.ncu-rep --page raw --csv | python profiles/summarize_ncu.py"""
import csv
import re
import sys

COLS = [('launch__grid_size', 'grid', 1), ('launch__registers_per_thread', 'regs/thread', 1),
        ('gpu__time_duration.sum', 'duration us', None), ('dram__bytes_read.sum', 'DRAM read MB', None),
        ('dram__bytes_write.sum', 'DRAM write MB', None),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %', 1),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %', 1),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %', 1),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe % while active', 1),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe % while active', 1),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %', 1)]


def to_unit(v, unit, want):
    v = float(v.replace(',', '')) if v else 0.0
    if want == 'us':
        return v / 1e3 if unit in ('ns', 'nsecond') else v * 1e3 if unit in ('ms', 'msecond') else v * 1e6 if unit in ('s', 'second') else v
    if want == 'MB':
        return {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(unit, 1.0) * v
    return v


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print('# columns: kernel | ' + ' | '.join(c[1] for c in COLS))
    for r in rows[2:]:
        name = re.sub(r'\(.*', '', r[idx['Kernel Name']])
        out = [name]
        for key, label, _ in COLS:
            i = idx.get(key)
            if i is None:
                out.append('-')
                continue
            want = 'us' if 'duration' in key else 'MB' if 'bytes' in key else None
            out.append(f'{to_unit(r[i], units[i], want):.1f}' if r[i] != '' else '-')
        print(' | '.join(out))


if __name__ == '__main__':
    main()
