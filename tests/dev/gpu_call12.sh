#!/bin/bash
# round-2 GPU call 12: attention with independent K / V rings and a polling MMA issuer, STFT stage arithmetic; full GPU suite, probes, default bench
O=gpurun_out/r02; mkdir -p $O
timeout 300 python -m pytest tests/test_tc_f32_gpu.py -q -k fused_attention > $O/c12_attn_tests.txt 2>&1; tail -5 $O/c12_attn_tests.txt | cut -c1-300
PROFILE=1 timeout 300 python tests/dev/demucs_probe.py 4 > $O/c12_htdemucs_profile_b4.txt 2>&1; head -1 $O/c12_htdemucs_profile_b4.txt; sed -n 6,9p $O/c12_htdemucs_profile_b4.txt | cut -c1-60,150-230
timeout 300 python tests/dev/demucs_probe.py 8 2>&1 | head -1
timeout 300 python tests/dev/roformer_probe.py 2 2>&1 | head -1
timeout 1500 python -m pytest tests -m gpu -q > $O/c12_gpu_tests.txt 2>&1; tail -6 $O/c12_gpu_tests.txt | cut -c1-300
timeout 900 python bench.py > $O/c12_bench_n1.json 2> $O/c12_bench_n1.err; tail -3 $O/c12_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/c12_bench_n1.json').read().strip().splitlines()[-1])
print('mdx', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline'].get('by_category_ms'), round(d['roofline']['achieved'],1), d['clocks'], d['parity']['max_abs_diff'])
a=d.get('also',{}).get('htdemucs_ft')
if a: print('htdemucs_ft', round(a['value'],1), 'e2e', round(a['e2e']['value'],1), a['roofline']['achieved'], a.get('cpu_baseline',{}).get('value'))
PY
