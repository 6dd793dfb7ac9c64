#!/bin/bash
# round-2 GPU call 7: resident weights at scale 0, conv channel-step choice for wide tiles (MDX23C), full GPU suite
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/c7_gpu_tests.txt 2>&1; tail -6 $O/c7_gpu_tests.txt
for wr in 1 0; do
  B200SEP_WRES=$wr timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --also none > $O/c7_bench_wres$wr.json 2> $O/c7_bench_wres$wr.err; tail -2 $O/c7_bench_wres$wr.err
  python - <<PY
import json
d=json.loads(open('$O/c7_bench_wres$wr.json').read().strip().splitlines()[-1])
print('resident weights $wr', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['by_category_ms'], round(d['roofline']['achieved'],1), d['clocks'], d['parity'])
PY
done
timeout 600 python bench.py --workload mdx23c --minutes 1 --steps 2 --warmup 3 --no-cpu-baseline > $O/c7_bench_mdx23c.json 2> $O/c7_bench_mdx23c.err; tail -3 $O/c7_bench_mdx23c.err; cut -c1-250 $O/c7_bench_mdx23c.json
timeout 600 python bench.py --workload vr --tracks 2 --minutes 1 --steps 2 --warmup 3 --no-cpu-baseline > $O/c7_bench_vr.json 2> $O/c7_bench_vr.err; tail -3 $O/c7_bench_vr.err; cut -c1-250 $O/c7_bench_vr.json
