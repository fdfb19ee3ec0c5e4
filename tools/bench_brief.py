"""Print the headline numbers of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'], 3), 'value', f"{d['value']:.4g}", 'e2e', f"{d['e2e']['value']:.4g}")
print({k: v.get('ms_per_launch') for k, v in d['kernels'].items()})
if d.get('adjoint'):
    print('adjoint ms/step', round(d['adjoint']['ms_per_step'], 2), d['adjoint']['stage_ms_per_step'])
