#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read with `ncu -i X.ncu-rep --page raw --csv`) into one row per kernel launch:
duration, DRAM bytes, tensor / FMA pipe activity, shared-memory pipe, occupancy.  Usage: summarize_ncu.py raw.csv out.json"""
import csv
import json
import re
import sys

KEYS = {
    'ms': 'gpu__time_duration.sum',
    'dram_read_GB': 'dram__bytes_read.sum',
    'dram_write_GB': 'dram__bytes_write.sum',
    'tensor_pipe_active_pct': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'fma_pipe_active_pct': 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'smem_pipe_pct': 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'dram_pct_of_peak': 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'l2_hit_pct': 'lts__t_sector_hit_rate.pct',
    'l2_sectors': 'lts__t_sectors.sum',
    'issue_active_pct': 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'warps_active_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'regs': 'launch__registers_per_thread',
    'grid': 'launch__grid_size',
    'block': 'launch__block_size',
    'smem_dyn_KB': 'launch__shared_mem_per_block_dynamic',
}


# ncu picks a unit per column (ns/us/ms, byte/Kbyte/Mbyte/Gbyte): durations are brought to ms, byte counts to GB
SCALE = {('ms', 'ns'): 1e-6, ('ms', 'us'): 1e-3, ('ms', 'ms'): 1.0, ('ms', 's'): 1e3,
         ('GB', 'byte'): 1e-9, ('GB', 'Kbyte'): 1e-6, ('GB', 'Mbyte'): 1e-3, ('GB', 'Gbyte'): 1.0}


def main(raw, out):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {'kernel': re.sub(r'\(.*', '', r[hdr.index('Kernel Name')]).replace('void ', '')}
        for k, name in KEYS.items():
            if name in hdr:
                v = r[hdr.index(name)].replace(',', '')
                u = units[hdr.index(name)]
                try:
                    d[k] = float(v) * SCALE.get((k.split('_')[-1], u), 1.0)
                except ValueError:
                    d[k] = v
                d.setdefault('_units', {})[k] = u
        res.append(d)
    json.dump(res, open(out, 'w'), indent=1)
    for d in res:
        print('%-44s %8.3f ms  dram r/w %6.3f/%6.3f GB  tensor %5.1f%%  fma %5.1f%%  smem %5.1f%%  dram %5.1f%%  regs %d' % (
            d['kernel'][:44], d['ms'], d.get('dram_read_GB', 0), d.get('dram_write_GB', 0), d.get('tensor_pipe_active_pct', 0),
            d.get('fma_pipe_active_pct', 0), d.get('smem_pipe_pct', 0), d.get('dram_pct_of_peak', 0), int(d.get('regs', 0))))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
