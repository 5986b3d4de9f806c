python -m pytest tests -m gpu -x -q -k "gate or sparse_tail or additive" 2>&1 | tail -2
for cfg in "0 1 296" "0 1 280" "0 0 296" "1 1 444" "1 1 420" "1 1 296" "1 0 296"; do set -- $cfg
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off -k regex:additive_attn_gate --csv --log-file /tmp/l.csv python bench.py --profile 1 > /dev/null 2>&1
T=$(python - <<'PY'
import csv
lines=[l for l in open('/tmp/l.csv') if not l.startswith('==')]
print(' '.join(f"{float(r['Metric Value'].replace(',',''))/1000:.1f}({r['Grid Size']})" for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum'))
PY
)
B=$(python bench.py --steps 6 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_decode_step']*1000,1))")
echo "impl $1 prop $2 slots $3 : gate us $T | bench $B"
done
