#!/bin/bash
# duration of the two additive_attn_gate launches of one decode step (warm caches), then a bench line
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off -k regex:additive_attn_gate --csv --log-file /tmp/l.csv python bench.py --profile 1 > /dev/null 2>&1
python - <<'PY'
import csv
lines=[l for l in open('/tmp/l.csv') if not l.startswith('==')]
print('gate launches (us):', ' '.join(f"{float(r['Metric Value'].replace(',',''))/1000:.1f}{r['Grid Size']}" for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum'))
PY
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['value']), round(d['ms_per_decode_step']*1000,1), round(d['e2e']['value']))"
