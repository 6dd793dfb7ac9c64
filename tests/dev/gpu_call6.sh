#!/bin/bash
# round-2 GPU call 6 (2 GPUs): MMA-vs-N probe, world-2 sharded tests, N=2 bench of every workload
O=gpurun_out/r02; mkdir -p $O
( cd tests/dev && timeout 120 ./mma_n_probe ) > $O/mma_n_probe.txt 2>&1; cat $O/mma_n_probe.txt
timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q > $O/c6_sharded_tests.txt 2>&1; tail -8 $O/c6_sharded_tests.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > $O/c6_bench_n2.json 2> $O/c6_bench_n2.err; tail -3 $O/c6_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02/c6_bench_n2.json').read().strip().splitlines()[-1])
    print('N=2 mdx', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], d['parity'])
    a=d['also']['htdemucs_ft']; print('N=2 htdemucs_ft', round(a['value'],1), 'e2e', round(a['e2e']['value'],1), a['parity'], a['roofline']['achieved'])
except Exception as e: print('parse failed', e)
PY
timeout 900 $TR bench.py --gpus 2 --workload mdx23c --minutes 2 --steps 2 --warmup 3 --no-cpu-baseline > $O/c6_bench_mdx23c_n2.json 2> $O/c6_bench_mdx23c_n2.err; tail -3 $O/c6_bench_mdx23c_n2.err; cut -c1-200 $O/c6_bench_mdx23c_n2.json
timeout 900 $TR bench.py --gpus 2 --workload vr --tracks 4 --minutes 1 --steps 2 --warmup 3 --no-cpu-baseline > $O/c6_bench_vr_n2.json 2> $O/c6_bench_vr_n2.err; tail -3 $O/c6_bench_vr_n2.err; cut -c1-200 $O/c6_bench_vr_n2.json
python - <<'PY'
import json
for f in ('c6_bench_mdx23c_n2','c6_bench_vr_n2'):
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['parity'], round(d['roofline']['achieved'],1))
    except Exception as e: print(f, 'parse failed', e)
PY
