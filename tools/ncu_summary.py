#!/usr/bin/env python
"""Summarise ncu outputs into small text files under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.md>     # per-kernel totals / shares from a
                                                                    # gpu__time_duration.sum launch list
  python tools/ncu_summary.py full <report.ncu-rep> <out.md>       # key counters of a --set full capture
  python tools/ncu_summary.py mix <report.ncu-rep> <out.md>        # APPEND instruction mix + top stall sites (source page)
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg',
        'smsp__inst_executed.sum']


def short(name):
    name = name.replace('void ', '')
    return name.split('(')[0][:90]


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, im = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    tot = OrderedDict()
    for r in rows[1:]:
        if r[im] != 'gpu__time_duration.sum':
            continue
        k = short(r[ik])
        t, n = tot.get(k, (0.0, 0))
        tot[k] = (t + float(r[iv].replace(',', '')), n + 1)
    total = sum(t for t, _ in tot.values())
    with open(out, 'w') as f:
        f.write('# kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: '
                'compare SHARES)\n\nsource: %s\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n' % path)
        for k, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
            f.write('| `%s` | %d | %.1f | %.1f%% |\n' % (k, n, t / 1e3, 100 * t / total))
        f.write('\ntotal %.1f us over %d launches\n' % (total / 1e3, sum(n for _, n in tot.values())))


def full(path, out):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write('# ncu --set full --clock-control none summary\n\nsource: %s\n' % path)
        for r in rows[2:]:
            f.write('\n## %s\n\n| metric | value | unit |\n|---|---|---|\n' % short(r[hdr.index('Kernel Name')]))
            for k in KEYS:
                if k in hdr:
                    f.write('| %s | %s | %s |\n' % (k, r[hdr.index(k)], units[hdr.index(k)]))


def mix(path, out):
    """append the warp-instruction mix and the top stall sites of each kernel (ncu --page source) to `out`"""
    import collections
    txt = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = dict(name=short(r[1]), rows=[])
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    with open(out, 'a') as f:
        for b in blocks:
            h = b['rows'][0]
            ia, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
            op, ops, data = collections.Counter(), collections.Counter(), []
            for r in b['rows'][1:]:
                if len(r) <= isamp:
                    continue
                toks = r[ia].split()
                o = (toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '')).split('.')[0]
                op[o] += int(r[iex] or 0)
                ops[o] += int(r[isamp] or 0)
                data.append((int(r[isamp] or 0), int(r[iex] or 0), r[ia].strip()[:90]))
            tot, ts = sum(op.values()), max(1, sum(ops.values()))
            f.write('\n## %s: warp instructions by opcode (%.3g executed)\n\n| opcode | %% of instructions | %% of stall samples |\n|---|---|---|\n' % (b['name'], tot))
            for o, c in op.most_common(14):
                f.write('| %s | %.1f | %.1f |\n' % (o, 100.0 * c / tot, 100.0 * ops[o] / ts))
            f.write('\ntop stall sites (samples, executions, SASS):\n\n')
            for d in sorted(data, reverse=True)[:10]:
                f.write('    %d  %d  %s\n' % d)


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'mix': mix}[sys.argv[1]](sys.argv[2], sys.argv[3])
