#!/bin/bash
# round-2 GPU call 21: A/B of the sleeping waits and of 8 MDX chunks per forward on one box, then the full suite and the default bench of the final build
O=gpurun_out/r02; mkdir -p $O
for ns in 0 100; do
  B200SEP_WAIT_SLEEP_NS=$ns timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > $O/c21_sleep$ns.json 2> $O/c21_sleep$ns.err
  python - <<PY
import json
d=json.loads(open('$O/c21_sleep$ns.json').read().strip().splitlines()[-1]); a=d['also']['htdemucs_ft']
print('wait sleep $ns ns: mdx', round(d['value'],1), d['roofline']['by_category_ms'], 'htdemucs_ft', round(a['value'],1))
PY
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --also none --batch 8 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mdx batch 8:', round(d['value'],1), round(d['e2e']['value'],1))"
for ns in 0 100; do B200SEP_WAIT_SLEEP_NS=$ns timeout 300 python tests/dev/roformer_probe.py 2 2>&1 | head -1; done
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_gpu_tests.txt 2>&1; tail -4 $O/r02_gpu_tests.txt | cut -c1-300
timeout 900 python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; tail -2 $O/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/r02_bench_n1.json').read().strip().splitlines()[-1])
print('mdx', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['roofline']['achieved'],1), d['roofline']['frac'], d['clocks'], d['parity']['max_abs_diff'], d['cpu_baseline']['value'])
a=d.get('also',{}).get('htdemucs_ft')
if a: print('htdemucs_ft', round(a['value'],1), 'e2e', round(a['e2e']['value'],1), a['roofline']['achieved'], a.get('cpu_baseline',{}).get('value'))
PY
