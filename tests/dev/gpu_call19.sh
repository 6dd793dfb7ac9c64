#!/bin/bash
# round-2 GPU call 19 (2 GPUs): world-2 sharded tests (NCCL halo + gather) and the N=2 bench of the default run (mdx + htdemucs_ft) with its parity field
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_sharded_gpu.py -q > $O/r02_sharded_tests_n2.txt 2>&1; tail -4 $O/r02_sharded_tests_n2.txt | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err; tail -2 $O/r02_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02/r02_bench_n2.json').read().strip().splitlines()[-1])
    print('N=2 mdx', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['parity'])
    a=d['also']['htdemucs_ft']; print('N=2 htdemucs_ft', round(a['value'],1), 'e2e', round(a['e2e']['value'],1), a['parity'])
except Exception as e: print('parse failed', e)
PY
