"""Short view of a bench.py JSON line (for gpurun tails)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        line = [l for l in open(path).read().splitlines() if l.startswith('{')][-1]
        d = json.loads(line)
    except Exception as e:  # noqa: BLE001
        print(path, 'unreadable:', e)
        continue
    r = d.get('roofline') or {}
    print(json.dumps({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}))
    print(' roofline', {k: r.get(k) for k in ('frac', 'whole_sweep_frac', 'per_kernel_frac', 'ms_per_launch', 'phase_ms_per_step')})
    print(' clocks', d.get('clocks'))
    p = d.get('parity_subsample')
    if p:
        print(' parity', {k: p.get(k) for k in ('dense', 'screened', 'max_scaled_error', 'ok')})
    for k in ('screened_path', 'overlap_regime', 'e2e'):
        if d.get(k):
            print(' ' + k, json.dumps(d[k])[:600])
    if d.get('other_configs'):
        for k, o in d['other_configs'].items():
            r2 = (o.get('roofline') or {})
            print(' other', k, json.dumps({kk: o.get(kk) for kk in ('ms_per_step', 'value', 'cuda_graph', 'error')}),
                  {kk: r2.get(kk) for kk in ('frac', 'whole_sweep_frac', 'phase_ms_per_step')})
