#!/bin/bash
# round-2 GPU call 2: new conv3x3 kernel (TMA-store epilogue) -- unit tests, MDX / MDXC / Demucs parity, bench
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_umma_gpu.py -x -q > $O/c2_umma.txt 2>&1; tail -5 $O/c2_umma.txt
timeout 900 python -m pytest tests/test_mdx_gpu.py tests/test_mdxc_gpu.py -x -q > $O/c2_mdx.txt 2>&1; tail -5 $O/c2_mdx.txt
timeout 900 python -m pytest tests/test_demucs_gpu.py -x -q > $O/c2_demucs.txt 2>&1; tail -5 $O/c2_demucs.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/c2_bench_n1.json 2> $O/c2_bench_n1.err; cut -c1-300 $O/c2_bench_n1.json; tail -3 $O/c2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/c2_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['by_category_ms'], d['roofline']['achieved'])
PY
