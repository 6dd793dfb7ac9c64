#!/bin/bash
# round-2 GPU call 17: checkpoint of the final build: full GPU suite, default bench (mdx + htdemucs_ft), reference arm, launch list, mdx23c and vr bench lines
O=gpurun_out/r02; mkdir -p $O
PROFILE=1 timeout 300 python tests/dev/demucs_probe.py 4 > $O/r02_htdemucs_profile_b4_final.txt 2>&1; head -1 $O/r02_htdemucs_profile_b4_final.txt; grep -E "dconv_row|tc_attention|tc_f32_kernel" $O/r02_htdemucs_profile_b4_final.txt | cut -c1-70,150-230
PROFILE=1 timeout 300 python tests/dev/hdemucs_probe.py 4 40 > $O/r02_hdemucs_probe_b4_final.txt 2>&1; head -1 $O/r02_hdemucs_probe_b4_final.txt; grep -E "lstm_bidir" $O/r02_hdemucs_probe_b4_final.txt | cut -c1-70,150-230
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_gpu_tests.txt 2>&1; tail -4 $O/r02_gpu_tests.txt | cut -c1-300
timeout 900 python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; tail -2 $O/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/r02_bench_n1.json').read().strip().splitlines()[-1])
print('mdx', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline'].get('by_category_ms'), round(d['roofline']['achieved'],1), d['roofline']['frac'], d['clocks'], d['parity']['max_abs_diff'], d['cpu_baseline']['value'])
a=d.get('also',{}).get('htdemucs_ft')
if a: print('htdemucs_ft', round(a['value'],1), 'e2e', round(a['e2e']['value'],1), a['roofline']['achieved'], a.get('cpu_baseline',{}).get('value'))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; cut -c1-200 $O/r02_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_final.csv python bench.py --minutes 0.5 --steps 1 --warmup 1 --no-cpu-baseline --no-parity --also none > $O/ncu_bench.log 2>&1; wc -l $O/r02_launches_final.csv
timeout 600 python bench.py --workload mdx23c --minutes 1 --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_bench_mdx23c.json 2> $O/r02_bench_mdx23c.err; tail -1 $O/r02_bench_mdx23c.err; cut -c1-220 $O/r02_bench_mdx23c.json
timeout 600 python bench.py --workload vr --tracks 2 --minutes 1 --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_bench_vr.json 2> $O/r02_bench_vr.err; tail -1 $O/r02_bench_vr.err; cut -c1-220 $O/r02_bench_vr.json
