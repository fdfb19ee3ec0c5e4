"""Print the per-stage numbers of bench.py JSON lines (usage: bench_show.py file.json ...)."""
import json
import sys

for f in sys.argv[1:]:
    txt = open(f).read().strip()
    if not txt:
        print('=====', f, 'EMPTY'); continue
    d = json.loads(txt.splitlines()[-1])
    fa = d.get('fwd_adjoint')
    print('=====', f, 'ms/step', round(d['ms_per_step'], 2), 'value %.3e' % d['value'], 'fwd+adj ms',
          fa and round(fa['ms_per_step'], 2), 'reorders', d.get('storage_reorders'), 'e2e ms', d.get('e2e', {}).get('ms_per_step'))
    for k, v in (d.get('kernels') or {}).items():
        print('   F %-18s ms/launch %-8s frac %-7s share %-7s total %s' % (k, v.get('ms_per_launch'), v.get('frac'), v.get('share_of_step'), v.get('ms_total')))
    if d.get('adjoint'):
        for k, v in d['adjoint'].get('kernels', {}).items():
            print('   A %-18s ms/launch %-8s frac %-7s ms/step %s' % (k, v.get('ms_per_launch'), v.get('frac'), v.get('ms_per_step')))
    if d.get('phase_ms_per_step_rank0'):
        print('   phases', d['phase_ms_per_step_rank0'])
        if d.get('adjoint'): print('   adj phases', d['adjoint'].get('phase_ms_per_step_rank0'))
    print('   roofline', {k: d['roofline'].get(k) for k in ('kernel', 'leg', 'frac', 'step_frac')})
