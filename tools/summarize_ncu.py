"""Turn ncu CSV exports into the markdown summaries kept under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv  N_FORWARDS  > profiles/rNN_launches.md
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep                  > profiles/rNN_kernel.md
"""
import collections
import csv
import subprocess
import sys


def launches(path, nfwd):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        a = agg.setdefault(row['Kernel Name'][:88], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print('ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)')
    print('%d launches, %.1f us total, %d forward passes -> %.1f us per forward\n' % (
        sum(v[0] for v in agg.values()), tot, nfwd, tot / nfwd))
    print('| kernel | launches / fwd | us / fwd | avg us | share |')
    print('|---|---:|---:|---:|---:|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %.1f | %.1f | %.2f | %.1f %% |' % (k, v[0] / nfwd, v[1] / nfwd, v[1] / v[0], 100 * v[1] / tot))


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.per_cycle_active', 'sm__cycles_elapsed.max']


def full(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print('ncu --set full --clock-control none (per launch)\n')
    for r in rows[2:]:
        print('### `%s`\n' % r[idx['Kernel Name']])
        print('| metric | value | unit |\n|---|---:|---|')
        for w in WANT:
            if w in idx:
                print('| %s | %s | %s |' % (w, r[idx[w]], units[idx[w]]))
        print()


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], int(sys.argv[3]))
    else:
        full(sys.argv[2])
