#!/bin/bash
# round-2 GPU call 3: weight multicast over CTA clusters (correctness + cluster-size sweep), fused DConv, new bench + sharded world-1 tests
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_demucs_gpu.py -x -q > $O/c3_umma_demucs.txt 2>&1; tail -15 $O/c3_umma_demucs.txt
timeout 900 python -m pytest tests/test_mdx_gpu.py tests/test_mdxc_gpu.py tests/test_sharded_gpu.py -x -q > $O/c3_mdx.txt 2>&1; tail -15 $O/c3_mdx.txt
for cs in 4 2 1; do
  B200SEP_CLUSTER=$cs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --also none > $O/c3_bench_cs$cs.json 2> $O/c3_bench_cs$cs.err; tail -2 $O/c3_bench_cs$cs.err
  python - <<PY
import json
d=json.loads(open('$O/c3_bench_cs$cs.json').read().strip().splitlines()[-1])
print('cluster $cs', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['by_category_ms'], round(d['roofline']['achieved'],1), d['clocks'], d['parity'])
PY
done
PROFILE=1 timeout 600 python tests/dev/demucs_probe.py 4 > $O/c3_demucs_probe_b4.txt 2>&1; grep -E "^batch|b200sep" $O/c3_demucs_probe_b4.txt | cut -c1-60,100-200 | head -30
timeout 600 python tests/dev/demucs_probe.py 8 > $O/c3_demucs_probe_b8.txt 2>&1; tail -1 $O/c3_demucs_probe_b8.txt
timeout 900 python bench.py --steps 2 --warmup 3 > $O/c3_bench_full.json 2> $O/c3_bench_full.err; tail -3 $O/c3_bench_full.err; cut -c1-1500 $O/c3_bench_full.json
