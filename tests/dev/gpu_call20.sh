#!/bin/bash
# round-2 GPU call 20 (2 GPUs): find the N=2 failure of the second (htdemucs_ft) workload of the default bench
O=gpurun_out/r02; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --workload htdemucs_ft --minutes 1 --steps 1 --warmup 3 --no-cpu-baseline > $O/c20_direct.json 2> $O/c20_direct.err; grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/c20_direct.err | grep -B2 -A25 "Traceback" | head -60 | cut -c1-220; cut -c1-300 $O/c20_direct.json | tail -2
timeout 600 $TR bench.py --gpus 2 --minutes 1 --steps 1 --warmup 3 --no-cpu-baseline > $O/c20_default.json 2> $O/c20_default.err; grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/c20_default.err | grep -B2 -A25 "Traceback" | head -60 | cut -c1-220
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02/c20_default.json').read().strip().splitlines() if l.startswith('{')][-1]); print(d.get('also'))
except Exception as e: print('parse failed', e)
PY
