#!/usr/bin/env python
"""Key metrics of `ncu --set full` captures (gpurun_out/*.ncu-rep) as text for profiles/: one block per launch plus a
machine-readable line `<kernel> dram_bytes_read=<bytes> dram_bytes_write=<bytes>` that bench.py reads for
`roofline.traffic`.   usage: python tools/ncu_summary.py rep1.ncu-rep [rep2 ...] > profiles/rNN_ncu_full_metrics.txt"""
import csv
import io
import re
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu.sum']
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}

for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print('== %s' % path.split('/')[-1])
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        short = re.sub(r'\(.*', '', name).replace('void ', '').replace('unnamed>::', '').strip()
        print('  kernel: %s' % name[:150])
        b = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('    %-82s %s %s' % (w, r[i], units[i]))
                if w.startswith('dram__bytes'):
                    b[w] = int(round(float(r[i].replace(',', '')) * SCALE.get(units[i], 1)))
        print('%s dram_bytes_read=%d dram_bytes_write=%d' % (short, b.get('dram__bytes_read.sum', 0), b.get('dram__bytes_write.sum', 0)))
        print()
