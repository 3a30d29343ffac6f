#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of counters DESIGN.md / bench.py quote:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, io, re, subprocess, sys
KEYS = [r'^gpu__time_duration\.sum$', r'^dram__bytes_read\.sum$', r'^dram__bytes_write\.sum$', r'^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$',
        r'sm__pipe_tensor_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed$', r'^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$', r'^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$',
        r'^sm__warps_active\.avg\.pct_of_peak_sustained_active$', r'^launch__registers_per_thread$', r'^launch__grid_size$', r'^launch__block_size$',
        r'^launch__shared_mem_per_block_dynamic$', r'^lts__t_sector_hit_rate\.pct$', r'^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$',
        r'^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$', r'^smsp__inst_executed\.sum$', r'^sm__cycles_elapsed\.max$',
        r'^sm__inst_executed_pipe_tensor.*hmma.*pct', r'^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$']
for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f'# {path}  (ncu --set full --clock-control none; one launch per row)')
    for r in rows[2:]:
        name = dict(zip(hdr, r)).get('Kernel Name', '?')
        print(f'kernel: {name}')
        for h, u, v in zip(hdr, units, r):
            if any(re.search(k, h) for k in KEYS):
                print(f'    {h:90s} {v} {u}')
