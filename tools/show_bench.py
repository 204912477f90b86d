"""Pretty-print a bench.py JSON line: python tools/show_bench.py file.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.2f} {d['unit']}  ms/step {d['ms_per_step']:.1f}  e2e {d['e2e']['value']:.2f}  launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
r = d.get('roofline') or {}
print(f"roofline: achieved {r.get('achieved', 0):.1f} / peak {r.get('peak', 0):.1f} = {r.get('frac', 0):.3f}  share {r.get('share_of_step', 0):.3f}")
steps = r.get('profile_steps', d['steps'])
for k, v in (r.get('by_kernel') or {}).items():
    print(f"  {k:60s} {v['ms'] / steps / 5:8.2f} ms/fold  n/fold {v['n'] / steps / 5:7.1f}  {v['tflops']:6.1f} TFLOP/s  frac {v['frac']:.2f}")
print('score:', {k: (round(v, 4) if isinstance(v, float) else v) for k, v in (d.get('roofline_score') or {}).items() if k in ('achieved', 'frac', 'ms')})
if d.get('cpu_baseline'):
    print('cpu:', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
