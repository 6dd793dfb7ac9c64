#!/bin/bash
# round-2 GPU call 8 (re-entry after the container was re-created): full GPU suite, default bench (mdx + htdemucs_ft), reference arm, launch list
O=gpurun_out/r02; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/r02_gpu_tests.txt 2>&1; tail -4 $O/r02_gpu_tests.txt
timeout 900 python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; tail -3 $O/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/r02_bench_n1.json').read().strip().splitlines()[-1])
print('mdx', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline'].get('by_category_ms'), round(d['roofline']['achieved'],1), d['roofline']['frac'], d['clocks'], d['parity'], d['cpu_baseline'])
a=d.get('also',{}).get('htdemucs_ft')
if a: print('htdemucs_ft', round(a['value'],1), 'e2e', round(a['e2e']['value'],1), a.get('parity'), a['roofline'], a.get('cpu_baseline'))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; cut -c1-400 $O/r02_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_final.csv python bench.py --minutes 0.5 --steps 1 --warmup 1 --no-cpu-baseline --no-parity --also none > $O/ncu_bench.log 2>&1; wc -l $O/r02_launches_final.csv
