"""Summarise an `ncu --page raw --csv` export: one line per launch with the metrics the roofline argument needs.
usage: python tools/ncu_summary.py gpurun_out/prof_conv_raw.csv [more.csv ...]"""
import csv, sys
WANT = [('Kernel Name', 'kernel', 28), ('Grid Size', 'grid', 7), ('Block Size', 'block', 5), ('gpu__time_duration.sum', 'time', 10),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%', 8),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu%', 6),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu%', 6),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%', 6),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed', 'smem%', 6),
        ('sm__inst_issued.avg.pct_of_peak_sustained_active', 'issue%', 6),
        ('dram__bytes_read.sum', 'dram_rd', 10), ('dram__bytes_write.sum', 'dram_wr', 10),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 6),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%', 6),
        ('launch__registers_per_thread', 'regs', 5), ('launch__waves_per_multiprocessor', 'waves', 6),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%', 6)]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {}
    for name, short, w in WANT:
        c = [i for i, h in enumerate(hdr) if h == name]
        col[name] = c[0] if c else None
    print('#', path, '(%d launches)' % (len(rows) - 2))
    print(' '.join(('%-' + str(w) + 's') % s for _, s, w in WANT))
    print(' '.join(('%-' + str(w) + 's') % ((units[col[n]] if col[n] is not None else '-')[:w]) for n, _, w in WANT))
    for r in rows[2:]:
        out = []
        for n, s, w in WANT:
            v = r[col[n]] if col[n] is not None else '-'
            if n == 'Kernel Name':
                v = v.replace('kbnet::', '').split('(')[0][:w]
            out.append(('%-' + str(w) + 's') % v[:w])
        print(' '.join(out))
