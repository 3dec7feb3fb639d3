#!/usr/bin/env python
"""Summarise an .ncu-rep: per kernel key metrics (raw page) and the top stalled SASS lines (source page).
usage: python scripts/ncu_summary.py gpurun_out/prof_step.ncu-rep [--top 12] [--kernel regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.avg', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def run(args):
    return subprocess.run(['ncu'] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 10
    kre = re.compile(sys.argv[sys.argv.index('--kernel') + 1]) if '--kernel' in sys.argv else None
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    kn = hdr.index('Kernel Name')
    print('## raw metrics')
    for r in rows[2:]:
        if kre and not kre.search(r[kn]):
            continue
        print('\n' + r[kn][:110])
        for i, h in enumerate(hdr):
            hh = h.split('.TriageCompute.')[-1]
            if hh in KEYS:
                print('   %-80s %s %s' % (hh, r[i], units[i]))
    print('\n## top stalled instructions (source page, warp stall samples)')
    text = run(['-i', rep, '--page', 'source', '--csv'])
    blocks = re.split(r'(?m)^"Kernel Name",', text)
    seen = set()
    for b in blocks[1:]:
        lines = list(csv.reader(io.StringIO('"Kernel Name",' + b)))
        name = lines[0][1]
        if (kre and not kre.search(name)) or name in seen:
            continue
        seen.add(name)
        h = lines[1]
        if 'Source' not in h:
            continue
        si, wi = h.index('Source'), h.index('Warp Stall Sampling (All Samples)')
        stall_cols = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
        data = []
        for ln in lines[2:]:
            try:
                data.append((float(ln[wi] or 0), ln))
            except (ValueError, IndexError):
                pass
        tot = sum(d[0] for d in data) or 1.0
        data.sort(key=lambda d: -d[0])
        print('\n' + name[:110] + '   (total samples %d)' % tot)
        for w, ln in data[:top]:
            st = sorted(((float(ln[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
            print('  %5.1f%%  %-70s %s' % (100 * w / tot, ln[si][:70], ' '.join('%s=%d' % (n, v) for v, n in st if v)))


if __name__ == '__main__':
    main()
