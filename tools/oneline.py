"""Print the headline fields of bench.py JSON lines: python tools/oneline.py FILE..."""
import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); r = d.get('roofline') or {}
            print(f, 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round((d.get('e2e') or {}).get('value', 0), 1),
                  'frac', r.get('frac'), 'cond', r.get('conditioning_ms_per_batch'), 'clk', (d.get('clocks') or {}).get('sm_mhz'))
